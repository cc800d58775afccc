python -m pytest tests/test_gpu_unet.py -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20
run() { python bench.py --no-cpu-baseline --steps 30 --warmup 5 "${@:2}" 2>>gpurun_out/b67.err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"; }
for i in 1 2; do
UNETK_SINGLE_GRAPH=0 run four_graphs
run one_graph
done
