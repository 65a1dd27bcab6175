python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
for cfg in "C4 0.2" "C3 1.0"; do set -- $cfg
  timeout 300 python bench.py --config $1 --scale $2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err || tail -5 gpurun_out/bench_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json")); print(sys.argv[1], round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"]["mme"], d["ms_per_step"])
PY
done
