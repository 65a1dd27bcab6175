# round-1 v3 kernels: tests, bench (both arms), ncu launch list, ncu full of the two sweeps
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v3_c2.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mme_flat_kernel|nn_flat_kernel' -c 3 -o gpurun_out/prof_v3 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
python - <<'PY'
import json
for f in ("bench_v3","bench_v3_c2"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["gpu_launches"], d["roofline"]["frac"], d["clocks"])
PY
