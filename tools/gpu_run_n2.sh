set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n2.json")); print(d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"])
PY
