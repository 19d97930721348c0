#!/bin/bash
# `ncu --set full` of the attention forward kernel only (scripts/heavy_kernels.py shapes); rows for profiles/r2_ncu_heavy_kernels.tsv
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'attn_fwd_tc3' -f -o /tmp/prof_attn_fwd python scripts/heavy_kernels.py > gpurun_out/ncu_attn_fwd.log 2>&1
echo "capture exit=$?"
ncu -i /tmp/prof_attn_fwd.ncu-rep --page raw --csv > gpurun_out/r2_attn_fwd_raw.csv 2>gpurun_out/ncu_export2.log
python scripts/ncu_summary.py gpurun_out/r2_attn_fwd_raw.csv > gpurun_out/r2_ncu_attn_fwd.tsv
cat gpurun_out/r2_ncu_attn_fwd.tsv | cut -c1-200
