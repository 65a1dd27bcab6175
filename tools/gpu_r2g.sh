# round 2, run G (one GPU): the slab-layout tests (one context per rank on one device), then the whole GPU suite and the C3 line
set -x
timeout 900 python -m pytest tests/test_gpu_slab.py -q --timeout=300 --durations=5 2>&1 | tail -40 | tee gpurun_out/pytest_r2g_slab.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 --deselect tests/test_gpu_slab.py 2>&1 | tail -8 | tee gpurun_out/pytest_r2g.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; tail -3 gpurun_out/bench_r2g.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2g.json")); print(round(d["value"],2), d["ms_per_step"], d.get("e2e"), {k:round(v,3) for k,v in d["stage_ms"].items()})
except Exception as e: print("no line", e)
PY
