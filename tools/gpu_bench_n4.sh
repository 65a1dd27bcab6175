python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
python -c "
import json
d=json.load(open('gpurun_out/bench_n4.json')); print(d['n_gpus'], round(d['value'],1), round(d['e2e']['value'],1), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['stage_ms'].items()}, d['check']['n_inlier'][-1], d['check']['full_cd'])"
