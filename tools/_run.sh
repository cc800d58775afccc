python -m pytest tests/test_gpu_variants.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
UNETK_BILINEAR_BLOCK=0 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -k upsample 2>&1 | grep -E "^E  |passed|failed|Error" | head -5
run() { python bench.py --no-cpu-baseline --steps 20 --warmup 5 "${@:2}" 2>>gpurun_out/b61.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['whole_step']['frac'] if 'whole_step' in d['roofline'] else '')"; }
UNETK_BILINEAR_BLOCK=0 run nested_pixel --model NestedUNet
run nested_block --model NestedUNet
UNETK_BILINEAR_BLOCK=0 run nested_pixel --model NestedUNet
run nested_block --model NestedUNet
python tools/profile_step.py --model NestedUNet > gpurun_out/step_profile61_nested.txt 2>&1
grep -E "^upsample" gpurun_out/step_profile61_nested.txt | sort -k8 -n -r | head -8; grep -A10 "^total" gpurun_out/step_profile61_nested.txt | grep -E "total|upsample"
