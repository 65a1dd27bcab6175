# round 2, run A: full-size C3 oracle fixture (CPU, on the box's host cores), the whole GPU suite incl. the new scene
# parity tests with the round-1 MME kernel and with the new rows kernel, A/B of the MME variants, a bench line
set -x
nproc; free -g | head -2
python tests/golden/make_c3_oracle.py --out gpurun_out/c3_oracle.json 2>&1 | tail -2
cp gpurun_out/c3_oracle.json tests/golden/c3_oracle.json
ME_MME_KERNEL=flat timeout 1500 python -m pytest tests -m gpu -q --durations=12 2>&1 | tail -60 > gpurun_out/pytest_r2a_flat.log; tail -30 gpurun_out/pytest_r2a_flat.log
timeout 1200 python -m pytest tests -m gpu -q -k "mme or c3 or c4 or c5 or scene or host or golden" 2>&1 | tail -40 > gpurun_out/pytest_r2a_rows.log; tail -25 gpurun_out/pytest_r2a_rows.log
timeout 600 python tools/ab_kernels.py C3 "ME_MME_KERNEL=flat" "" "ME_MME_ROWS=8,2" "ME_MME_ROWS=16,2" "ME_MME_ROWS=12,4" "ME_MME_ROWS=16,4" 2>&1 | tee gpurun_out/ab_r2a.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -3 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2a.json")); print(round(d["value"],2), d["e2e"], d.get("e2e_pageable"), {k:round(v,3) for k,v in d["stage_ms"].items()}, d.get("roofline_binding"), d["check"])
PY
