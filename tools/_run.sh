python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
run() { python bench.py --no-cpu-baseline --steps 20 --warmup 5 "${@:2}" 2>>gpurun_out/b60.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['whole_step']['frac'] if 'whole_step' in d['roofline'] else '')"; }
UNETK_SPLIT_WIDE=0 run nested_nosplit --model NestedUNet
UNETK_SPLIT_WIDE=32 run nested_split32 --model NestedUNet
run nested_split64 --model NestedUNet
UNETK_SPLIT_WIDE=0 run nested_nosplit --model NestedUNet
run nested_split64 --model NestedUNet
python tools/profile_step.py --model NestedUNet > gpurun_out/step_profile60_nested.txt 2>&1
grep -E "^conv3x3_dgrad" gpurun_out/step_profile60_nested.txt | sort -k7 -n -r | head -8
