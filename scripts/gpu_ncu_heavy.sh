#!/bin/bash
# `ncu --set full` of scripts/heavy_kernels.py (each heavy kernel at its headline shape): plain run first, then one ncu
# invocation; the raw page is exported to CSV on the box so that only small files travel back.
set -u
mkdir -p gpurun_out
CMD="python scripts/heavy_kernels.py"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none \
    -k regex:'conv_tc_kernel|wgrad_tc|attn_fwd_tc|attn_bwd_tc|gn_bwd_apply|gn_bwd_reduce|gn_apply_kernel|gn_stats_kernel|fcomb_members' \
    -f -o /tmp/prof_r1_heavy $CMD > gpurun_out/ncu_heavy.log 2>&1
echo "capture exit=$?"
ncu -i /tmp/prof_r1_heavy.ncu-rep --page raw --csv > gpurun_out/r1_heavy_raw.csv 2>gpurun_out/ncu_export.log
ls -la /tmp/prof_r1_heavy.ncu-rep gpurun_out | tail -n 8
SZ=$(stat -c %s /tmp/prof_r1_heavy.ncu-rep)
if [ "$SZ" -lt 40000000 ]; then cp /tmp/prof_r1_heavy.ncu-rep gpurun_out/; fi
