python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench58_n2.json 2> gpurun_out/bench58_n2.err
echo "rc=$?"; wc -c gpurun_out/bench58_n2.json; tail -15 gpurun_out/bench58_n2.err; nvidia-smi -L
