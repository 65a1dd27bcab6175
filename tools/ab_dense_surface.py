#!/usr/bin/env python
"""MME on dense surfaces (C5 regime: ~1 cm spacing, r = 0.1 m -> ~314 neighbours): shared lattice vs the sweep's own."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cloud_map_evaluation_b200 import _abi as A, api, synth
n = 8_000_000
est = synth.indoor_scene(n, synth.EST_SEED, synth.EST_NOISE_SIGMA, rooms=2)      # 4 rooms, ~10^4 pts/m^2
d = torch.from_numpy(est).cuda()
for shared in (True, False):
    if shared: os.environ["ME_MME_SHARED_LATTICE"] = "1"
    else: os.environ.pop("ME_MME_SHARED_LATTICE", None)
    with api.MapEvalB200() as ctx:
        for it in range(3):
            ctx.set_cloud_device(A.ME_CLOUD_EST, d.data_ptr(), n, keepalive=d)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            m = ctx.eval_mme_accum(A.ME_CLOUD_EST, 0.1, 10)
            torch.cuda.synchronize(); t1 = time.perf_counter()
        st = ctx.stage_times_ms()
    print(f"shared_lattice={shared}: total {1e3*(t1-t0):.2f} ms (grid {st['grid_est']:.2f}, mme {st['mme_est']:.2f}) mme {m.sum_entropy/m.n_valid:.10f} n_valid {m.n_valid}", flush=True)
