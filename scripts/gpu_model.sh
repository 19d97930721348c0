#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x --timeout 300 -k "not tc" > gpurun_out/ops_base.log 2>&1; echo "ops base exit=$?"; tail -n 3 gpurun_out/ops_base.log
timeout 1500 python -m pytest tests/test_model_gpu.py -m gpu -q -rA -s --timeout 600 "$@" > gpurun_out/model.log 2>&1; echo "model exit=$?"
grep -E "^(PASSED|FAILED|ERROR)|rel err|total |worst|step [0-9]" gpurun_out/model.log | tail -n 60
