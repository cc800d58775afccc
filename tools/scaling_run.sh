#!/bin/bash
# One call on an 8-GPU box (gpurun --gpus 8): weak scaling of the headline config and the strong-scaling curve of
# BASELINE.json configs[2] (AttentionUNet, global batch 16 -> 16/8/4/2 images per GPU), plus the per-kernel profile of
# the per-GPU shape at N = 8 (batch 2), which names the kernels that limit it.  Results: gpurun_out/r02_scale_*.json
mkdir -p gpurun_out
run() {  # n, tag, extra args...
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-variants "$@" \
      > gpurun_out/r02_scale_${tag}_n1.json 2> gpurun_out/r02_scale_${tag}_n1.err
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $n --steps 20 --warmup 5 "$@" > gpurun_out/r02_scale_${tag}_n$n.json 2> gpurun_out/r02_scale_${tag}_n$n.err
  fi
  echo "$tag n=$n rc=$?"
}
run 1 weak
run 8 weak
for n in 1 2 4 8; do run $n strong --model AttentionUNet --scaling strong --global-batch 16; done
timeout 200 python tools/profile_step.py --model AttentionUNet --batch 2 > gpurun_out/r02_step_profile_AttentionUNet_b2.txt 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_scale_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms", d["config"]["per_gpu_batch"], "per GPU", d.get("replicas_in_sync"),
              (d.get("grad_buckets") or {}).get("mode"))
    except Exception as e:
        print(f, "ERR", e)
PY
