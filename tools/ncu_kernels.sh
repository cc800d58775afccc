#!/bin/bash
# Per-launch ncu counters of single library calls (tools/one_kernel.py): time, DRAM bytes, L2 -> SM bytes, tensor-pipe activity.
#   tools/ncu_kernels.sh out.csv "conv3x3_wgrad --cin 128 --cout 64" "upconv_wgrad --s 256 --cin 128 --cout 64" ...
out=$1; shift
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
: > $out
for spec in "$@"; do
  echo "## $spec ${UNETK_ENV}" >> $out
  timeout 300 ncu --metrics $M --clock-control none --csv -k regex:'wgrad|conv_gemm|conv3x3|halo|rows' python tools/one_kernel.py $spec --reps 1 2>/dev/null | grep -v "^==" >> $out
done
