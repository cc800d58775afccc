#!/bin/bash
# A/B of the SMs left to NCCL while gradient buckets are in flight (Trainer.sm_reserve / NCCL_MAX_CTAS) on one box:
#   bash tools/ab_dp_reserve.sh N "r0:c0 r1:c1 ..."      (c = - : NCCL's default CTA count)
N=${1:-2}; CFGS=${2:-"0:- 4:4 8:8"}
mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-baseline --no-variants > gpurun_out/r02_ab${N}_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02_ab${N}_n1.json')); print('n1:', round(d['value'],1), round(d['ms_per_step'],3))"
for cfg in $CFGS; do r=${cfg%%:*}; c=${cfg##*:}
  if [ "$c" = "-" ]; then export NCCL_MAX_CTAS=64; else export NCCL_MAX_CTAS=$c; fi
  UNETK_DP_SM_RESERVE=$r timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02_ab${N}_r${r}_c${c}.json 2>gpurun_out/r02_ab${N}_r${r}_c${c}.err
  python -c "
import json; d=json.load(open('gpurun_out/r02_ab${N}_r${r}_c${c}.json')); print('N=$N reserve $r ctas $c:', round(d['value'],1), round(d['ms_per_step'],3), d['replicas_in_sync'])" || tail -3 gpurun_out/r02_ab${N}_r${r}_c${c}.err
done
