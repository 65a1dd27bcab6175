#!/usr/bin/env python
"""Generates tests/golden/c3_oracle.json: the CPU oracle (oracle/, the restatement of the reference's metric path) run
ONCE on the full headline workload C3 (10 M vs 10 M points) — every scalar and every integer count of
the pass, plus order-sensitive checksums of the two nearest-neighbour index arrays.  The pass takes about a minute on
the 128 host cores of the B200 box (several minutes on a small machine); tests/test_gpu_fullsize.py compares the CUDA
path with this fixture, and bench.py's `check` block must agree with it.

    python tests/golden/make_c3_oracle.py [--out tests/golden/c3_oracle.json] [--config C3] [--scale 1.0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cloud_map_evaluation_b200 import _abi as A  # noqa: E402
from cloud_map_evaluation_b200 import synth  # noqa: E402


def index_checksum(idx):
    """order-sensitive checksum of an int32 index array, mod 2^64"""
    i = np.arange(idx.shape[0], dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(np.sum((idx.astype(np.int64) + 1).astype(np.uint64) * (i * np.uint64(2654435761) + np.uint64(1)),
                          dtype=np.uint64))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "c3_oracle.json"))
    ap.add_argument("--config", default="C3")
    ap.add_argument("--scale", type=float, default=1.0)
    args = ap.parse_args()
    from oracle import oracle as O
    est, gt, cfg = synth.make_pair(args.config, scale=args.scale)
    p = A.make_nn_params(cfg["tau"], 1.0)              # path A as written + full CD, as bench.py runs it
    t = {}
    t0 = time.perf_counter()
    nn, ie, ig = O.eval_nn(est, gt, p, want_indices=True)
    t["nn_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    mme, ent = O.eval_mme(est, cfg["nn_radius"], 10, want_entropies=True)
    t["mme_s"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    awd = O.eval_awd(est, gt, cfg["vmd_voxel_size"], 100, 5)
    t["awd_s"] = time.perf_counter() - t0
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count()
    out = {
        "config": args.config, "scale": args.scale, "n_est": int(len(est)), "n_gt": int(len(gt)),
        "tau": cfg["tau"], "icp_max_distance": 1.0, "nn_radius": cfg["nn_radius"], "vmd_voxel_size": cfg["vmd_voxel_size"],
        "generator": "cloud_map_evaluation_b200/synth.py make_pair (seeds 20250001 / 20250002)",
        "made_by": "tests/golden/make_c3_oracle.py", "host_threads": threads, "oracle_seconds": t,
        "nn": A.struct_to_dict(nn),
        "nn_index_checksum": {"est_to_gt": index_checksum(ie), "gt_to_est": index_checksum(ig)},
        "mme": A.struct_to_dict(mme),
        "mme_entropy_sum": float(np.sum(ent)), "mme_entropy_abs_sum": float(np.sum(np.abs(ent))),
        "mme_nonzero": int(np.count_nonzero(ent)),
        "awd": A.struct_to_dict(awd),
    }
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {args.out}: nn {t['nn_s']:.1f} s, mme {t['mme_s']:.1f} s, awd {t['awd_s']:.1f} s on {threads} threads")


if __name__ == "__main__":
    main()
