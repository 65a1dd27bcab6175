#!/usr/bin/env python
"""A/B of kernel variants on one config (device-resident clouds): stage times per environment setting.
    python tools/ab_kernels.py C3 "ME_MME_VARIANT=81" "ME_MME_VARIANT=162,ME_MME_CARVEOUT=50" ...
"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cloud_map_evaluation_b200 import _abi as A, api, synth

cfg_name = sys.argv[1]
settings = sys.argv[2:] or [""]
est, gt, cfg = synth.make_pair(cfg_name)
dev = torch.device("cuda", 0)
d_est = torch.from_numpy(est).to(dev); d_gt = torch.from_numpy(gt).to(dev)
p = A.make_nn_params(cfg["tau"], 1.0)
TUNE = [k for k in os.environ if k.startswith("ME_")]
for st in settings:
    for k in list(os.environ):
        if k.startswith("ME_"): del os.environ[k]
    cell = 0.0
    for kv in filter(None, st.split(";")):
        k, v = kv.split("=", 1)
        if k == "CELL": cell = float(v)
        else: os.environ[k] = v
    ctx = api.MapEvalB200(device=0, vmd_voxel_size=cfg["vmd_voxel_size"], nn_cell_size=cell)
    acc = {}
    reps = 6
    for it in range(reps + 2):
        ctx.set_cloud_device(A.ME_CLOUD_EST, d_est.data_ptr(), len(est), keepalive=d_est)
        ctx.set_cloud_device(A.ME_CLOUD_GT, d_gt.data_ptr(), len(gt), keepalive=d_gt)
        m = ctx.eval_mme_accum(A.ME_CLOUD_EST, cfg["nn_radius"], 10)
        e, g = ctx.eval_nn_accum(p)
        if it >= 2:
            for k, v in ctx.stage_times_ms().items(): acc[k] = acc.get(k, 0) + v / reps
    res = ctx.nn_finalize(p, e, g)
    print(f"{st or 'default':45s} mme {acc['mme_est']:.3f} nn {acc['nn_est_to_gt']:.3f} {acc['nn_gt_to_est']:.3f} grid {acc['grid_est']:.3f} {acc['grid_gt']:.3f}"
          f" | chk mme {m.sum_entropy / max(1, m.n_valid):.12f} nv {m.n_valid} ninl {list(res.est_to_gt.n_inlier)[-1]} cd {res.full_cd:.12f} far {e.n_far},{g.n_far}", flush=True)
    ctx.close()
