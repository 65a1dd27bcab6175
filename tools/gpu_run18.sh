python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_v4.json")); print(round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"]["n_far"], d["check"]["full_cd"])
PY
