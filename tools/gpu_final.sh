# final round-1 evidence: tests, both bench arms, ncu launch list, ncu full of the two sweeps
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --config C2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final_c2.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mme_flat_kernel|nn_flat_kernel' -c 3 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import json
for f in ("bench_final","bench_final_c2","bench_ref"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, round(d["value"],2), round(d["e2e"]["value"],2), d.get("stage_ms") and {k:round(v,3) for k,v in d["stage_ms"].items()}, d.get("gpu_launches"), d.get("roofline",{}).get("frac"), d.get("clocks"), d.get("cpu_baseline"))
PY
