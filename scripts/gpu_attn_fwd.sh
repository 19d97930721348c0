timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "tmem_accumulator" --timeout 300 2>&1 | tail -15
PU_ATTN_FWD=3 python scripts/bench_attn.py 2>&1 | tail -3
python scripts/bench_attn.py 2>&1 | tail -3
