python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench66_n8.json 2> gpurun_out/bench66_n8.err
echo "rc=$?"; python -c "
import json
for ln in open('gpurun_out/bench66_n8.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d.get('loss'), d['clocks'])"
tail -3 gpurun_out/bench66_n8.err
