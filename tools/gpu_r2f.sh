# round 2, run F (ONE 8-GPU box): C4 (50 M vs 20 M) on 4 GPUs and C5 (200 M vs 200 M) on 8 GPUs, full size, torchrun + NCCL.
# A failed line is retried once with the replicated-lattice layout (ME_NO_SLAB=1) so the box time is not wasted.
set -x
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -9
run() {  # N CFG STEPS PORT
  local N=$1 CFG=$2 STEPS=$3 PORT=$4 tag=bench_$2_n$1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --config $CFG --steps $STEPS --warmup 2 --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
  local rc=$?
  if [ $rc -ne 0 ] || ! python -c "import json,sys; d=json.load(open('gpurun_out/$tag.json')); sys.exit(0 if d.get('value') else 1)"; then
    echo "rc=$rc: retry with ME_NO_SLAB=1"; tail -5 gpurun_out/$tag.err; cp gpurun_out/$tag.err gpurun_out/$tag.first.err
    ME_NO_SLAB=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
      bench.py --gpus $N --config $CFG --steps $STEPS --warmup 2 --no-cpu-baseline > gpurun_out/$tag.json 2> gpurun_out/$tag.err
    echo "rc=$?"
  fi
  tail -3 gpurun_out/$tag.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$tag.json"))
    print("$tag", round(d["value"],2), d["ms_per_step"], d["e2e"], {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"], d["config"])
except Exception as e:
    print("$tag: no bench line:", e)
PY
}
run 4 C4 5 29517
run 8 C5 3 29527
