"""Regenerates tests/golden/voxel_fixture.npz from the reference's shipped sample outputs.

Run in the authoring container (needs /root/reference; the GPU box does not have it):
    python tests/golden/make_golden.py

Sources (data fixtures of the reference, not source code):
  map_eval/scripts/voxel_errors.txt         27 columns written by MapEval::calculateVMD (map_eval.cpp:292-302):
      vmin[3] vmax[3] mu_est[3] W n_gt n_est sigma_est[6: 00 01 02 11 12 22] mu_gt[3] sigma_gt[6]
  map_eval/scripts/voxel_wasserstein_cdf.txt  sorted W + (i+1)/n (map_eval.cpp:337-340)
Known answers of the same run, from the README run log (README.md:170, image-20250214100110872.png):
  VMD 0.35303, SCS 0.78121 (printed with setprecision(5), map_eval.cpp:462-463).
"""
import json
import os
import numpy as np

REF = "/root/reference/map_eval/scripts"
HERE = os.path.dirname(os.path.abspath(__file__))

rows = np.loadtxt(os.path.join(REF, "voxel_errors.txt"), dtype=np.float64)
cdf = np.loadtxt(os.path.join(REF, "voxel_wasserstein_cdf.txt"), dtype=np.float64)
assert rows.shape == (7129, 27) and cdf.shape == (7129, 2)
np.savez_compressed(os.path.join(HERE, "voxel_fixture.npz"), rows=rows, cdf=cdf)
with open(os.path.join(HERE, "readme_run_log.json"), "w") as f:
    json.dump({"source": "README.md:170 (image-20250214100110872.png), scene redbird_02",
               "voxel_size": 3.0, "VMD": 0.35303, "SCS": 0.78121, "print_precision": 5,
               "n_est": 12795056, "n_gt": 132012045}, f, indent=1)
print("rows", rows.shape, "cdf", cdf.shape, os.path.getsize(os.path.join(HERE, "voxel_fixture.npz")), "bytes")
