#!/bin/bash
# ncu launch list of the bench command (B200_PROFILING.md recipe): plain run first, then one pass collecting
# gpu__time_duration.sum for every launch of the STEADY-STATE timed steps only (bench.py brackets its timed region with
# cudaProfilerStart/Stop when PU_NCU_RANGE=1; warm-up, optimizer-state initialisation and the end-to-end region stay
# outside).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --headline-only --no-profile-calls"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
PU_NCU_RANGE=1 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 20000 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
python scripts/launch_summary.py gpurun_out/launches.csv > gpurun_out/r2_launch_list.tsv
head -30 gpurun_out/r2_launch_list.tsv
