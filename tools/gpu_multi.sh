# usage: bash tools/gpu_multi.sh N:CONFIG:STEPS[:LAYOUT] ... — one bench line per spec on N GPUs of this box (torchrun + NCCL;
# LAYOUT slab (default) or replicated).  A failed or hung line is retried once with the replicated layout.
set -x
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -9
PORT=29517
for spec in "$@"; do
  IFS=: read N CFG STEPS LAYOUT <<< "$spec"
  LAYOUT=${LAYOUT:-slab}
  tag=bench_${CFG}_n${N}_${LAYOUT}
  for attempt in 1 2; do
    PORT=$((PORT+1))
    if [ "$N" = "1" ]; then
      timeout 150 python bench.py --gpus 1 --config $CFG --steps $STEPS --warmup 3 --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
    else
      timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
        bench.py --gpus $N --config $CFG --steps $STEPS --warmup 3 --no-cpu-baseline --layout $LAYOUT > gpurun_out/$tag.json 2> gpurun_out/$tag.err
    fi
    rc=$?
    if [ $rc -eq 0 ] && python -c "import json,sys; d=json.load(open('gpurun_out/$tag.json')); sys.exit(0 if d.get('value') else 1)"; then break; fi
    echo "rc=$rc"; tail -8 gpurun_out/$tag.err; cp gpurun_out/$tag.err gpurun_out/$tag.first.err
    [ "$LAYOUT" = "replicated" ] && break
    LAYOUT=replicated; tag=bench_${CFG}_n${N}_${LAYOUT}
  done
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag.json"))
    print("$tag", round(d["value"],2), d["ms_per_step"], (d.get("e2e") or {}).get("value"), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"], d["config"]["parallelism"])
except Exception as e:
    print("$tag: no bench line:", e)
PY
done
