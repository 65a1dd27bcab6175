# round 2, run E: the large BASELINE configs on ONE B200 (C4: 50 M vs 20 M, C5: 200 M vs 200 M; clouds generated on the device)
set -x
timeout 900 python bench.py --config C4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_C4_n1.json 2> gpurun_out/bench_C4_n1.err; tail -3 gpurun_out/bench_C4_n1.err
timeout 1200 python bench.py --config C5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_C5_n1.json 2> gpurun_out/bench_C5_n1.err; tail -3 gpurun_out/bench_C5_n1.err
python - <<'PY'
import json
for f in ("bench_C4_n1","bench_C5_n1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), d["ms_per_step"], d.get("e2e"), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"])
    except Exception as e: print(f, "no line", e)
PY
