#!/bin/bash
# Same-box A/B of two builds of libunetk.so (box-to-box variance is ~5%, larger than most kernel changes):
#   tools/ab.sh jcfszxc_unet_b200/ab/libunetk_A.so jcfszxc_unet_b200/ab/libunetk_B.so [rounds]
A=$1; B=$2; R=${3:-2}
mkdir -p gpurun_out
for i in $(seq 1 $R); do
  for v in A B; do
    lib=$A; [ $v = B ] && lib=$B
    UNETK_LIB=$lib timeout 300 python tools/profile_step.py > gpurun_out/ab_${v}_$i.txt 2>&1
    echo "== $v round $i: $(grep -E '^total' gpurun_out/ab_${v}_$i.txt)"
    grep -E "^  unetk_(conv3x3_fwd_bnstats|conv3x3_dgrad|conv3x3_dgrad_colsum|conv3x3_wgrad|convT2x2_fwd|bn_bwd_reduce|bn_bwd_apply|bn_apply) " gpurun_out/ab_${v}_$i.txt
  done
done
