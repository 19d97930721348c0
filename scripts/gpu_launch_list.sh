#!/bin/bash
# ncu launch list of the bench command (B200_PROFILING.md recipe): plain run first, then one pass collecting
# gpu__time_duration.sum for every launch.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --batch 64 --no-cpu-baseline --no-profile-calls --ensemble-members 0"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit=$?"
ls -la gpurun_out | tail -n 6
