# compute-sanitizer on a small full pass (all kernels of the path incl. down-sampling, far queries, tile and plane fallbacks)
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from cloud_map_evaluation_b200 import _abi as A, api, synth
est, gt, cfg = synth.make_pair("C2", scale=0.02)
far = np.array([[30.0, 30.0, 30.0], [-12.0, 3.0, 1.0]])
est = np.ascontiguousarray(np.concatenate([est, far]))
p = A.make_nn_params(cfg["tau"], 1.0)
for env in ({}, {"ME_NN_TILE": "1", "ME_MME_SHARED_LATTICE": "1"}):
    for k in ("ME_NN_TILE", "ME_MME_SHARED_LATTICE"): os.environ.pop(k, None)
    os.environ.update(env)
    for cell in (0.0, 0.02):
        with api.MapEvalB200(vmd_voxel_size=cfg["vmd_voxel_size"], nn_cell_size=cell, rank=1, world=3) as ctx:
            ctx.set_cloud(A.ME_CLOUD_EST, est); ctx.set_cloud(A.ME_CLOUD_GT, gt)
            ctx.voxel_downsample(A.ME_CLOUD_EST, 0.01)
            m = ctx.eval_mme_accum(A.ME_CLOUD_EST, 0.1, 10)
            e, g = ctx.eval_nn_accum(p)
            ctx.get_nn(A.ME_CLOUD_EST); ctx.get_entropies(A.ME_CLOUD_EST)
            a = ctx.calculateVMD(cfg["vmd_voxel_size"], 100, 5)
            print(env, cell, m.n_valid, e.n_corr, g.n_corr, e.n_far, a.n_pairs, flush=True)
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitize_$tool.log | head -10; tail -3 gpurun_out/sanitize_$tool.log
done
