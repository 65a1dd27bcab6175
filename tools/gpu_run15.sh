python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --config C4 --scale 0.2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_C4.json 2> gpurun_out/bench_C4.err || tail -5 gpurun_out/bench_C4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_C4.json")); print("C4", round(d["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"]["mme"])
PY
