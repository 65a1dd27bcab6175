for cfg in "C5 0.05" "C4 0.2"; do set -- $cfg
  timeout 300 python bench.py --config $1 --scale $2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err || tail -5 gpurun_out/bench_$1.err
  python - "$1" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json")); print(sys.argv[1], d["config"]["n_est"], d["config"]["n_gt"], round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"])
PY
done
