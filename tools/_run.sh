python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_variants.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
run() { python bench.py --no-cpu-baseline --steps 20 --warmup 5 "${@:2}" 2>>gpurun_out/b63.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"; }
UNETK_FUSE_COPIES=0 run nested_copy_kernels --model NestedUNet
run nested_fused_copies --model NestedUNet
UNETK_FUSE_COPIES=0 run nested_copy_kernels --model NestedUNet
run nested_fused_copies --model NestedUNet
UNETK_LIB=jcfszxc_unet_b200/ab/libunetk_A.so run unet_A
run unet_B
