python -m pytest tests/test_gpu_conv_kernels.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
run() { python bench.py --no-cpu-baseline --steps 20 --warmup 5 "${@:2}" 2>>gpurun_out/b64.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"; }
UNETK_SCATTER_DGRAD=0 run nested_adds --model NestedUNet
run nested_scatter --model NestedUNet
UNETK_SCATTER_DGRAD=0 run nested_adds --model NestedUNet
run nested_scatter --model NestedUNet
python tools/profile_step.py --model NestedUNet > gpurun_out/step_profile64_nested.txt 2>&1
grep -A12 "^total" gpurun_out/step_profile64_nested.txt
tail -3 gpurun_out/b64.err
