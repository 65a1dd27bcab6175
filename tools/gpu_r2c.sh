# round 2, run C: GPU suite (dense tables) with a per-test timeout, the suite again with every lattice on the sparse cell
# table, A/B of the sweep / build variants, bench lines (C3, S1 site-scale sparse)
set -x
timeout 900 python -m pytest tests -m gpu -q -x --timeout=240 --durations=8 2>&1 | tail -40 > gpurun_out/pytest_r2c.log; tail -18 gpurun_out/pytest_r2c.log
ME_FORCE_SPARSE=1 timeout 900 python -m pytest tests -m gpu -q --timeout=240 --deselect tests/test_gpu_fullsize.py::test_c3_full_size_properties 2>&1 | tail -50 > gpurun_out/pytest_r2c_sparse.log; tail -25 gpurun_out/pytest_r2c_sparse.log
timeout 420 python tools/ab_kernels.py C3 "" "ME_NN_KERNEL=rows" "ME_NN_KERNEL=rows4" "ME_NN_KERNEL=rows8" "ME_NN_KERNEL=tma" "ME_NN_KERNEL=tma4" "ME_BUILD=bucket" "ME_FORCE_SPARSE=1" 2>&1 | tee gpurun_out/ab_r2c.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -3 gpurun_out/bench_r2c.err
timeout 600 python bench.py --config S1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_S1_n1.json 2> gpurun_out/bench_S1_n1.err; tail -3 gpurun_out/bench_S1_n1.err
python - <<'PY'
import json
for f in ("bench_r2c","bench_S1_n1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), d["ms_per_step"], d.get("e2e"), d.get("e2e_pageable"), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"])
    except Exception as e: print(f, "no line", e)
PY
