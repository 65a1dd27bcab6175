# round 2, run D (final single-GPU evidence): the whole GPU suite, smoke, the C3 bench line, the ncu launch list of one bench
# command and a full ncu capture of the two sweeps
set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 --durations=6 2>&1 | tail -14 | tee gpurun_out/pytest_r2d.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; tail -3 gpurun_out/bench_r2d.err
timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2d_c2.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2d.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu_r2d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mme_flat_kernel|nn_rows_kernel' -c 3 -o gpurun_out/prof_r2d python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/b_ncu2_r2d.log 2>&1
python - <<'PY'
import json
for f in ("bench_r2d","bench_r2d_c2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), d["ms_per_step"], d.get("e2e"), d.get("e2e_pageable"), {k:round(v,3) for k,v in d["stage_ms"].items()}, d.get("gpu_launches"), d.get("roofline",{}).get("frac"), d.get("roofline_binding"), d.get("clocks"), d.get("cpu_baseline"))
    except Exception as e: print(f, "no line", e)
PY
