set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("bench_n1","bench_n2"):
    s=open(f"gpurun_out/{f}.json").read(); print(f, "lines:", s.count("\n"))
    d=json.loads(s); print(d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"])
PY
