#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): plain run first, then the launch list, then one
# --set full capture of the dominant kernels.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-profile-calls"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel|wgrad_tc_kernel|attn_fwd_tc|attn_bwd_tc|gn_bwd_apply|gn_bwd_reduce|gn_apply' -s 60 -c 24 -o gpurun_out/prof_r1 $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
ls -la gpurun_out | tail -n 12
