# round 2, run J (one GPU): the whole GPU suite (slab layout and plan memory included), then the C3 line
set -x
timeout 900 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -40 | cut -c1-600 | tee gpurun_out/pytest_r2j.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; tail -3 gpurun_out/bench_r2j.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2j.json")); print(round(d["value"],2), d["ms_per_step"], d.get("e2e"), d.get("e2e_pageable"), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"])
except Exception as e: print("no line", e)
PY
