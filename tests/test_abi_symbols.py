"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly the symbols the header
declares, the ctypes mirrors match the C struct sizes, and the host-only finalisers agree with the oracle's
restatement of the reference formulas (no GPU compute is called here)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from cloud_map_evaluation_b200 import _abi as A
from cloud_map_evaluation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mapeval_b200.h")).read()
    return sorted(set(re.findall(r"^ME_API\s+[\w\s\*]+?\b(me_[a-z_0-9]+)\s*\(", text, flags=re.M)))


def test_header_and_library_export_the_same_symbols(L):
    declared = _header_symbols()
    assert declared == sorted(_lib.SYMBOLS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = sorted(s for s in re.findall(r" T (me_[a-z_0-9]+)", out))
    assert exported == declared
    for s in declared:
        assert hasattr(L, s)


def test_struct_sizes_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    names = ["me_options", "me_nn_params", "me_nn_accum", "me_dir_result", "me_nn_result", "me_mme_accum",
             "me_mme_result", "me_awd_result"]
    src.write_text('#include <stdio.h>\n#include "mapeval_b200.h"\nint main(){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in names) + "return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes == [C.sizeof(getattr(A, n)) for n in names]
    assert C.sizeof(A.me_nn_accum) == 8 * (A.ME_NN_ACCUM_I64 + A.ME_NN_ACCUM_F64)


def test_no_cpu_fallback_without_a_device(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    opt = A.me_options()
    opt.abi_version, opt.world = A.ME_ABI_VERSION, 1
    ctx = C.c_void_p()
    assert L.me_create(C.byref(opt), C.byref(ctx)) == A.ME_ERR_NO_DEVICE
    assert b"no CPU path" in L.me_last_error(None)
    opt.abi_version = 99
    assert L.me_create(C.byref(opt), C.byref(ctx)) == A.ME_ERR_INVALID


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cloud_map_evaluation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("no cpu", ""), f"{f} mentions the oracle"


def test_finalizers_match_the_reference_formulas(L):
    """me_nn_finalize / me_mme_finalize (host only) against accumulators computed with numpy."""
    from oracle import oracle as O
    from cloud_map_evaluation_b200 import synth
    est, gt, cfg = synth.make_pair("C1", scale=0.05)
    tau = cfg["tau"]
    p = A.make_nn_params(tau, 1.0, pairing=A.ME_PAIRING_GEOMETRIC)
    exp = O.eval_nn(est, gt, p)
    accs = []
    for q, r in ((est, gt), (gt, est)):
        idx, d2 = O.knn1(q, r)
        keep = d2 <= 1.0
        d = q[keep] - r[idx[keep]]
        sq = d[:, 0] ** 2 + (d[:, 1] ** 2 + d[:, 2] ** 2)
        nd = np.sqrt(sq)
        a = A.me_nn_accum()
        a.n_query, a.n_corr = len(q), int(keep.sum())
        for k, t in enumerate(tau):
            m = nd <= t
            a.n_inlier[k], a.sum_d[k], a.sum_d2[k] = int(m.sum()), nd[m].sum(), sq[m].sum()
        a.sum_d_all, a.sum_d2_all, a.sum_nn_dist = nd.sum(), sq.sum(), np.sqrt(d2).sum()
        accs.append(a)
    out = A.me_nn_result()
    assert L.me_nn_finalize(C.byref(p), C.byref(accs[0]), C.byref(accs[1]), len(est), len(gt), C.byref(out)) == 0
    for d in ("est_to_gt", "gt_to_est"):
        g, e = getattr(out, d), getattr(exp, d)
        assert list(g.n_inlier) == list(e.n_inlier) and g.n_corr == e.n_corr
        for k in ("mean", "rmse", "fitness", "sigma"):
            np.testing.assert_allclose(list(getattr(g, k)), list(getattr(e, k)), rtol=1e-9, err_msg=k)
    for k in ("cd", "f1", "iou"):
        np.testing.assert_allclose(list(getattr(out, k)), list(getattr(exp, k)), rtol=1e-9)
    np.testing.assert_allclose(out.full_cd, exp.full_cd, rtol=1e-12)

    m = A.me_mme_accum()
    m.n_query, m.n_valid, m.sum_entropy, m.min_entropy, m.max_entropy = 10, 4, -32.0, -9.5, -6.25
    r = A.me_mme_result()
    assert L.me_mme_finalize(C.byref(m), 10, C.byref(r)) == 0
    assert (r.mme, r.n_valid, r.n_total, r.min_abs_entropy, r.max_abs_entropy) == (-8.0, 4, 10, 6.25, 9.5)
    m.n_valid, m.sum_entropy, m.min_entropy, m.max_entropy = 0, 0.0, np.inf, -np.inf
    assert L.me_mme_finalize(C.byref(m), 10, C.byref(r)) == 0
    assert r.mme == 0.0 and np.isnan(r.min_abs_entropy)
