python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v3c.json 2> gpurun_out/bench_v3c.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_v3c.json")); print(round(d["value"],1), round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["check"]["scs"], d["gpu_launches"])
PY
python - <<'PY'
# timing of the down-sampling pre-step on 10M points
import time, numpy as np, torch
from cloud_map_evaluation_b200 import _abi as A, api, synth
est = synth.uniform_box(10_000_000, synth.box_side_for_density(10_000_000), 7, noise_sigma=0.01)
h = torch.from_numpy(est).pin_memory()
with api.MapEvalB200() as ctx:
    for s in (0.01, 0.05):
        for it in range(3):
            ctx.set_cloud_ptr(A.ME_CLOUD_EST, h.data_ptr(), len(est), keepalive=h); ctx.synchronize()
            t0 = time.perf_counter(); n = ctx.voxel_downsample(A.ME_CLOUD_EST, s); ctx.synchronize(); t1 = time.perf_counter()
        print(f"voxel_downsample 10M points s={s}: {n} voxels, {1e3*(t1-t0):.2f} ms")
PY
