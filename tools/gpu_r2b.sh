# round 2, run B: full GPU suite on the current kernels (dense tables), again with every lattice forced onto the sparse
# cell table, A/B of the NN row walk and the MME row-walk variants, bench lines (C3, S1 site-scale sparse, C4/C5 at N=1)
set -x
timeout 1500 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -50 > gpurun_out/pytest_r2b.log; tail -22 gpurun_out/pytest_r2b.log
ME_FORCE_SPARSE=1 timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fullsize.py::test_c3_full_size_properties 2>&1 | tail -60 > gpurun_out/pytest_r2b_sparse.log; tail -30 gpurun_out/pytest_r2b_sparse.log
timeout 600 python tools/ab_kernels.py C3 "" "ME_NN_KERNEL=rows" "ME_NN_KERNEL=rows1" "ME_NN_KERNEL=rows4" "ME_NN_KERNEL=rows8" "ME_NN_KERNEL=tma" "ME_NN_KERNEL=tma4" "ME_MME_KERNEL=rows;ME_MME_ROWS=16,2" "ME_MME_KERNEL=rows;ME_MME_ROWS=12,4" "ME_BUILD=bucket" "ME_FORCE_SPARSE=1" 2>&1 | tee gpurun_out/ab_r2b.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -3 gpurun_out/bench_r2b.err
timeout 900 python bench.py --config S1 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_S1_n1.json 2> gpurun_out/bench_S1_n1.err; tail -3 gpurun_out/bench_S1_n1.err
timeout 900 python bench.py --config C4 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench_C4_n1.json 2> gpurun_out/bench_C4_n1.err; tail -3 gpurun_out/bench_C4_n1.err
timeout 900 python bench.py --config C5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C5_n1.json 2> gpurun_out/bench_C5_n1.err; tail -3 gpurun_out/bench_C5_n1.err
python - <<'PY'
import json
for f in ("bench_r2b","bench_S1_n1","bench_C4_n1","bench_C5_n1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), d["ms_per_step"], d.get("e2e"), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"])
    except Exception as e: print(f, "no line", e)
PY
