#!/bin/bash
# One `ncu --set full` capture of the heavy kernels of the bench command (B200_PROFILING.md recipe): plain run first,
# then a single ncu invocation.  SKIP / COUNT select a window of matching launches (forward end + backward start by
# default, i.e. the full-resolution layers of both passes).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
SKIP=${SKIP:-150}
COUNT=${COUNT:-130}
CMD="python bench.py --steps 1 --warmup 0 --batch 64 --no-cpu-baseline --no-profile-calls"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on \
    -k regex:'conv_tc_kernel|wgrad_tc2_kernel|wgrad_tc_kernel|attn_fwd_tc2|attn_bwd_tc2|gn_bwd_apply|gn_bwd_reduce|gn_apply_kernel|gn_stats_kernel' \
    -s $SKIP -c $COUNT -f -o gpurun_out/prof_r1_full $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
ls -la gpurun_out | tail -n 6
