python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
for v in 0 1; do
  if [ $v = 1 ]; then export ME_NO_PREFETCH=1; fi
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pf$v.json 2> gpurun_out/bench_pf$v.err
  python - "$v" <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_pf{sys.argv[1]}.json")); print("no_prefetch" if sys.argv[1]=="1" else "prefetch", round(d["value"],1), round(d["e2e"]["value"],1), round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()})
PY
done
