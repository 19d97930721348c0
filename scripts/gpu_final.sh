#!/bin/bash
# final validation of a build: every GPU test, the full bench line, the steady-state launch list
mkdir -p gpurun_out
timeout 1000 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/gputests.log 2>&1; echo "tests exit=$?"; tail -n 3 gpurun_out/gputests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit=$?"
bash scripts/gpu_launch_list.sh 2>&1 | tail -2
