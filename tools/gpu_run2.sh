# flat kernels: parity tests, bench, A/B against the previous kernels
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat.json 2> gpurun_out/bench_flat.err
ME_NN_TILE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nntile.json 2>/dev/null
ME_MME_WALK=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mmewalk.json 2>/dev/null
python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_flat_c2.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench_flat","bench_nntile","bench_mmewalk","bench_flat_c2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"]["mme"], d["check"]["n_far"])
    except Exception as e: print(f, "ERR", e)
PY
