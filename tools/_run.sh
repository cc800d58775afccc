python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
run() { python bench.py --no-cpu-baseline --steps 20 --warmup 5 "${@:2}" 2>>gpurun_out/b59.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['roofline']['whole_step']['frac'] if 'whole_step' in d['roofline'] else '')"; }
for i in 1 2; do
UNETK_LIB=jcfszxc_unet_b200/ab/libunetk_A.so run A
run B
done
