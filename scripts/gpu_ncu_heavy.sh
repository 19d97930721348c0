#!/bin/bash
# `ncu --set full` of scripts/heavy_kernels.py (each heavy kernel at its headline shape): plain run first, then one ncu
# invocation; the raw page is exported to CSV on the box so that only small files travel back.
set -u
mkdir -p gpurun_out
CMD="python scripts/heavy_kernels.py"
$CMD > gpurun_out/plain_heavy.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_heavy.log; exit 1; }
ncu --set full --clock-control none \
    -k regex:'conv_tc_kernel|wgrad_tc|attn_fwd_tc|attn_bwd_tc|attn_dq_convert|gn_bwd_apply|gn_bwd_reduce|gn_apply_kernel|gn_stats_kernel|fcomb_members|pack_weights_multi|adamw_multi' \
    -f -o /tmp/prof_r2_heavy $CMD > gpurun_out/ncu_heavy.log 2>&1
echo "capture exit=$?"
ncu -i /tmp/prof_r2_heavy.ncu-rep --page raw --csv > gpurun_out/r2_heavy_raw.csv 2>gpurun_out/ncu_export.log
python scripts/ncu_summary.py gpurun_out/r2_heavy_raw.csv > gpurun_out/r2_ncu_heavy_kernels.tsv
cat gpurun_out/r2_ncu_heavy_kernels.tsv | cut -c1-200
