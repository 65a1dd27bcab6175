"""GPU parity on the generators of the large BASELINE configs, at >= 1 M points and at the configs' own densities and
radii: C4 (outdoor surfaces, 1000 / 400 pts/m^2, MME radius 1.0 m on a lattice of its own, 3.0 m voxels), C5 (dense
indoor, ~9000 pts/m^2, r = 0.1 m, 2.0 m voxels), and ground-truth-style surfaces (2 mm and 0.2 mm noise, k >= 5) where
the smallest covariance eigenvalue amplifies every error of the accumulated offsets.

The full scenes are generated on the device (synth_torch, the same generators) and cropped, so that the oracle finishes
in seconds; both sides see the same arrays.  Bar: integer counts bit-exact, floats within 1e-5 (asserted tighter)."""
import numpy as np
import pytest

from cloud_map_evaluation_b200 import _abi as A
from cloud_map_evaluation_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-9
RTOL_MME = 1e-7
RTOL_ENT = 1e-6


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def api():
    from cloud_map_evaluation_b200 import api as _api
    return _api


def _cmp_dir(got, exp):
    assert got.n_source == exp.n_source and got.n_corr == exp.n_corr and got.n_ub == exp.n_ub
    assert list(got.n_inlier) == list(exp.n_inlier)
    for k in ("mean", "rmse", "fitness", "sigma"):
        np.testing.assert_allclose(list(getattr(got, k)), list(getattr(exp, k)), rtol=RTOL, atol=1e-300, err_msg=k)
    np.testing.assert_allclose(got.sum_nn_dist, exp.sum_nn_dist, rtol=RTOL)


def _full_parity(api, O, est, gt, cfg, gt_mme=False):
    p = A.make_nn_params(cfg["tau"], 1.0)
    with api.MapEvalB200(vmd_voxel_size=cfg["vmd_voxel_size"]) as ctx:
        ctx.set_cloud(A.ME_CLOUD_EST, est)
        ctx.set_cloud(A.ME_CLOUD_GT, gt)
        mme, ent = ctx.computeMME(A.ME_CLOUD_EST, cfg["nn_radius"], 10, want_entropies=True)
        mme_gt = ctx.computeMME(A.ME_CLOUD_GT, cfg["nn_radius"], 5, want_entropies=True) if gt_mme else None
        nn = ctx.calculateMetricsWithInitialMatrix(p)
        idx_e, d2_e = ctx.get_nn(A.ME_CLOUD_EST)
        idx_g, d2_g = ctx.get_nn(A.ME_CLOUD_GT)
        awd, rows = ctx.calculateVMD(cfg["vmd_voxel_size"], 100, 5, want_rows=True)
    onn, oie, oig = O.eval_nn(est, gt, p, want_indices=True)
    _cmp_dir(nn.est_to_gt, onn.est_to_gt)
    _cmp_dir(nn.gt_to_est, onn.gt_to_est)
    for k in ("cd", "f1", "iou"):
        np.testing.assert_allclose(list(getattr(nn, k)), list(getattr(onn, k)), rtol=RTOL, equal_nan=True, err_msg=k)
    np.testing.assert_allclose(nn.full_cd, onn.full_cd, rtol=RTOL)
    oi, od2 = O.knn1(est, gt)
    np.testing.assert_array_equal(d2_e, od2)          # bit-identical squared distances
    # surface scans hold exact ties (equal distances to two references): the index may differ only where d2 ties
    diff = np.nonzero(idx_e != oi)[0]
    for i in diff[:50]:
        d = est[i] - gt[[idx_e[i], oi[i]]]
        dd = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        assert dd[0] == dd[1]
    assert len(diff) == 0 or len(diff) < 1e-4 * len(est)
    omme, oent = O.eval_mme(est, cfg["nn_radius"], 10, want_entropies=True)
    assert mme.n_valid == omme.n_valid and mme.n_valid > 0.5 * len(est)
    np.testing.assert_array_equal(ent != 0, oent != 0)
    np.testing.assert_allclose(mme.mme, omme.mme, rtol=RTOL_MME)
    np.testing.assert_allclose(ent, oent, rtol=RTOL_ENT, atol=1e-12)
    np.testing.assert_allclose([mme.min_abs_entropy, mme.max_abs_entropy], [omme.min_abs_entropy, omme.max_abs_entropy], rtol=RTOL_ENT)
    if gt_mme:
        og, ogent = O.eval_mme(gt, cfg["nn_radius"], 5, want_entropies=True)
        assert mme_gt[0].n_valid == og.n_valid
        np.testing.assert_allclose(mme_gt[0].mme, og.mme, rtol=RTOL_MME)
        np.testing.assert_allclose(mme_gt[1], ogent, rtol=RTOL_ENT, atol=1e-12)
    oawd, orows = O.eval_awd(est, gt, cfg["vmd_voxel_size"], 100, 5, want_rows=True)
    for k in ("n_pairs", "n_scs", "n_voxels_est", "n_voxels_gt", "n_active", "n_old", "n_new"):
        assert getattr(awd, k) == getattr(oawd, k), k
    assert awd.n_pairs > 20
    np.testing.assert_allclose([awd.awd, awd.scs], [oawd.awd, oawd.scs], rtol=1e-8)
    v = cfg["vmd_voxel_size"]

    def srt(r):
        keys = np.rint(r[:, 0:3] / v).astype(np.int64)
        return r[np.lexsort((keys[:, 2], keys[:, 1], keys[:, 0]))]
    a, b = srt(rows), srt(orows)
    np.testing.assert_array_equal(a[:, [0, 1, 2, 3, 4, 5, 10, 11]], b[:, [0, 1, 2, 3, 4, 5, 10, 11]])
    np.testing.assert_allclose(a[:, 6:9], b[:, 6:9], rtol=1e-11)             # mu_est
    np.testing.assert_allclose(a[:, 18:21], b[:, 18:21], rtol=1e-11)         # mu_gt
    np.testing.assert_allclose(a[:, 9], b[:, 9], rtol=1e-7)                  # W
    np.testing.assert_allclose(a[:, 12:18], b[:, 12:18], rtol=1e-6, atol=1e-14)
    np.testing.assert_allclose(a[:, 21:27], b[:, 21:27], rtol=1e-6, atol=1e-14)
    return mme, nn, awd


def _crop(t, lo, hi):
    m = (t[:, 0] >= lo[0]) & (t[:, 0] < hi[0]) & (t[:, 1] >= lo[1]) & (t[:, 1] < hi[1])
    return t[m].cpu().numpy()


def test_c4_outdoor_generator_at_config_density(api, O):
    """C4: the full 50 M / 20 M outdoor scene generated on the device, a 36 x 36 m window of it evaluated:
    ~1.6 M est / ~0.65 M gt points, r = 1.0 m (~3000 neighbours, MME on a lattice of its own), v = 3.0 m."""
    import torch
    from cloud_map_evaluation_b200 import synth_torch
    est_t, gt_t, cfg = synth_torch.make_pair("C4", device="cuda")
    lo, hi = (92.0, 92.0), (128.0, 128.0)
    est, gt = _crop(est_t, lo, hi), _crop(gt_t, lo, hi)
    del est_t, gt_t
    torch.cuda.empty_cache()
    assert len(est) > 1_000_000 and len(gt) > 400_000
    assert cfg["nn_radius"] == 1.0 and cfg["vmd_voxel_size"] == 3.0
    mme, nn, awd = _full_parity(api, O, est, gt, cfg)
    assert mme.n_valid > 0.99 * len(est)


def test_c4_outdoor_whole_scene_sparse(api, O):
    """The whole 200 x 200 m scene at 1/25 of the point count (2 M / 0.8 M points): large extent, few points per cell."""
    est, gt, cfg = synth.make_pair("C4", scale=0.04)
    _full_parity(api, O, est, gt, cfg)


def test_c5_indoor_generator_at_config_density(api, O):
    """C5's generator at its own density: one 8 x 8 x 3 m room with 2 M points (8900 pts/m^2, as 200 M points over the
    10 x 10 rooms), r = 0.1 m (~280 neighbours), v = 2.0 m; GT-style MME (k >= 5) on the 2 mm-noise cloud as well."""
    cfg = dict(synth.CONFIGS["C5"])
    est = synth.indoor_scene(2_000_000, synth.EST_SEED, synth.EST_NOISE_SIGMA, rooms=1)
    gt = synth.indoor_scene(2_000_000, synth.GT_SEED, synth.GT_SURFACE_NOISE_SIGMA, rooms=1)
    _full_parity(api, O, est, gt, cfg, gt_mme=True)


@pytest.mark.parametrize("sigma", [2e-3, 2e-4])
def test_mme_on_nearly_flat_surfaces(api, O, sigma):
    """Ground-truth-style surfaces: a tilted plane at 1 cm spacing with 2 mm / 0.2 mm noise, far from the origin, k >= 5.
    lambda_1 / lambda_3 reaches ~6e4 at 0.2 mm, so errors of the accumulated offsets are amplified accordingly."""
    n = 400_000
    idx = np.arange(n, dtype=np.uint64)
    u = synth.uniform24(77, idx, 0).astype(np.float64) * 6.3
    v = synth.uniform24(77, idx, 1).astype(np.float64) * 6.3
    w = sigma * synth.gaussian(77, idx, 8)
    a, b = np.array([0.8, 0.36, 0.48]), np.array([-0.6, 0.48, 0.64])
    nrm = np.cross(a, b)
    pts = u[:, None] * a + v[:, None] * b + w[:, None] * nrm + np.array([731.25, -1204.5, 88.0])
    pts = pts.astype(np.float32).astype(np.float64)
    # fp32 storage at ~1000 m quantises coordinates to 6e-5 m: comparable to the 0.2 mm noise, as real GT scans are
    with api.MapEvalB200() as ctx:
        ctx.set_cloud(A.ME_CLOUD_EST, pts)
        ctx.set_cloud(A.ME_CLOUD_GT, pts[:1000])
        for k in (5, 10):
            got, ent = ctx.computeMME(A.ME_CLOUD_EST, 0.1, k, want_entropies=True)
            exp, oent = O.eval_mme(pts, 0.1, k, want_entropies=True)
            assert got.n_valid == exp.n_valid and got.n_valid > 0.9 * n
            np.testing.assert_array_equal(ent != 0, oent != 0)
            np.testing.assert_allclose(got.mme, exp.mme, rtol=RTOL_MME)
            np.testing.assert_allclose(ent, oent, rtol=1e-5, atol=1e-12)       # the stated bar, per point
            assert np.max(np.abs(ent - oent) / np.maximum(np.abs(oent), 1e-300)) < 2e-6
