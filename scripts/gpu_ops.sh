#!/bin/bash
# runs the per-kernel GPU tests in separate processes (a trapped kernel poisons its CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -rA -s --timeout 300 "$@" > gpurun_out/ops_$name.log 2>&1; echo "$name exit=$?"; tail -n 4 gpurun_out/ops_$name.log; }
run base -k "not tc"
run tcfwd -k "conv_tc_fwd"
run tcdgrad -k "dgrad and tc"
run tcwgrad -k "wgrad and tc"
