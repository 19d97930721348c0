timeout 1000 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/gputests.log 2>&1; echo "tests exit=$?"; tail -n 3 gpurun_out/gputests.log
python bench.py --steps 6 --warmup 3 --headline-only > gpurun_out/bench8.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/bench8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"]["sm_mhz"], {k:v for k,v in d["kernel_ms_per_step"].items() if v>0.4})
PY
