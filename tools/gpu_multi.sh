# usage: bash tools/gpu_multi.sh <N> <config> [steps] — one bench line of <config> on N GPUs of this box (torchrun, NCCL)
N=$1; CFG=$2; STEPS=${3:-5}
set -x
nvidia-smi --query-gpu=index,name,memory.total --format=csv | head -9
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 --config $CFG --steps $STEPS --warmup 2 --no-cpu-baseline > gpurun_out/bench_${CFG}_n1.json 2> gpurun_out/bench_${CFG}_n1.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --config $CFG --steps $STEPS --warmup 2 --no-cpu-baseline > gpurun_out/bench_${CFG}_n${N}.json 2> gpurun_out/bench_${CFG}_n${N}.err
fi
echo "rc=$?"; tail -5 gpurun_out/bench_${CFG}_n${N}.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${CFG}_n${N}.json"))
    print(round(d["value"],2), d["ms_per_step"], d["e2e"], {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"], d["config"])
except Exception as e:
    print("no bench line:", e)
PY
