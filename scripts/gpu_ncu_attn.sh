#!/bin/bash
# `ncu --set full --import-source on` of the attention backward / forward kernels at T=4096, heads=4, batch 64; exports the
# raw page and the per-instruction source page (stall reasons) as CSV so that only small files travel back.
set -u
mkdir -p gpurun_out
CMD="python scripts/one_attn.py 4 4096 64"
$CMD > gpurun_out/plain_attn.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_attn.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'attn_bwd_tc2|attn_fwd_tc2' -s 2 -c 2 \
    -f -o /tmp/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "capture exit=$?"
ncu -i /tmp/prof_attn.ncu-rep --page raw --csv > gpurun_out/attn_raw.csv 2>gpurun_out/ncu_attn_export.log
ncu -i /tmp/prof_attn.ncu-rep --page source --csv --print-source sass > gpurun_out/attn_source_sass.csv 2>>gpurun_out/ncu_attn_export.log
ls -la /tmp/prof_attn.ncu-rep gpurun_out | tail -n 8
