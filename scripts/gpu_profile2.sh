#!/bin/bash
# targeted --set full captures of the first (= full-resolution) launches of each heavy kernel
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-profile-calls"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
for K in conv_tc_kernel wgrad_tc_kernel attn_fwd_tc attn_bwd_tc gn_bwd_apply gn_bwd_reduce gn_apply_kernel bias_grad_vec; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 3 -o gpurun_out/prof_$K -f $CMD > gpurun_out/ncu_$K.log 2>&1
  echo "$K exit=$?"
done
