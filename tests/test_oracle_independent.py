"""Cross-checks the oracle's third-party stand-ins (KD-tree, 3x3 algebra) against independent implementations:
scipy.spatial.cKDTree, brute force and numpy.linalg.  These are the [ext] boundaries the reference delegates to
Open3D/nanoflann/Eigen and ships no fixtures for."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from cloud_map_evaluation_b200 import _abi as A
from cloud_map_evaluation_b200 import synth
from oracle import oracle as O


def _pair(n=20000, seed=0):
    side = synth.box_side_for_density(n)
    gt = synth.uniform_box(n, side, synth.GT_SEED + seed)
    est = synth.uniform_box(n + 137, side, synth.EST_SEED + seed, noise_sigma=0.01)
    return est, gt


def test_knn1_matches_bruteforce_and_ckdtree():
    est, gt = _pair(4000)
    idx, d2 = O.knn1(est, gt)
    diff = est[:, None, :] - gt[None, :, :]
    bf = ((diff[..., 0] ** 2 + diff[..., 1] ** 2) + diff[..., 2] ** 2)   # nanoflann accumulation order
    np.testing.assert_array_equal(idx, bf.argmin(axis=1).astype(np.int32))
    np.testing.assert_array_equal(d2, bf.min(axis=1))
    dk, ik = cKDTree(gt).query(est, k=1)
    np.testing.assert_array_equal(idx, ik.astype(np.int32))
    np.testing.assert_allclose(np.sqrt(d2), dk, rtol=1e-14)


def test_knn1_far_queries_and_duplicates():
    rng = np.random.RandomState(1)
    gt = rng.rand(3000, 3)
    gt[10] = gt[5]                      # duplicate reference point: lower index wins
    q = np.concatenate([gt[:50], rng.rand(200, 3) * 40 - 20, gt[5:6]])
    idx, d2 = O.knn1(q, gt)
    dk, _ = cKDTree(gt).query(q, k=1)
    np.testing.assert_allclose(np.sqrt(d2), dk, rtol=1e-14, atol=0)
    assert idx[-1] == 5 and d2[-1] == 0.0


def _numpy_dir(src, tgt, nn_idx, nn_d2, keep, tau):
    sel = np.nonzero(keep)[0]
    d = src[sel] - tgt[nn_idx[sel]]
    sq = d[:, 0] ** 2 + (d[:, 1] ** 2 + d[:, 2] ** 2)
    nd = np.sqrt(sq)
    nc = len(sel)
    out = dict(n_corr=nc, n_inlier=[], mean=[], rmse=[], fitness=[], sigma=[])
    for t in tau:
        m = nd <= t
        mean = nd[m].sum() / nc
        out["n_inlier"].append(int(m.sum()))
        out["mean"].append(mean)
        out["rmse"].append(np.sqrt(sq[m].sum() / nc))
        out["fitness"].append(m.sum() / len(src))
        out["sigma"].append(np.sqrt(((nd - mean) ** 2).sum() / nc))
    return out


@pytest.mark.parametrize("cutoff", [A.ME_CUTOFF_SQDIST_LE_R, A.ME_CUTOFF_DIST_LT_R])
def test_eval_nn_against_numpy(cutoff):
    est, gt = _pair(15000)
    tau = [0.2, 0.1, 0.08, 0.05, 0.01]
    R = 0.05  # small so that the cut-off actually drops pairs (sqrt(0.05) = 0.22 m vs 0.05 m)
    p = A.make_nn_params(tau, R, cutoff_mode=cutoff, pairing=A.ME_PAIRING_GEOMETRIC)
    res = O.eval_nn(est, gt, p)
    te, tg = cKDTree(gt), cKDTree(est)
    d_e, i_e = te.query(est)
    d_g, i_g = tg.query(gt)
    keep = (lambda d: d * d <= R) if cutoff == A.ME_CUTOFF_SQDIST_LE_R else (lambda d: d < R)
    for got, exp in ((res.est_to_gt, _numpy_dir(est, gt, i_e, d_e ** 2, keep(d_e), tau)),
                     (res.gt_to_est, _numpy_dir(gt, est, i_g, d_g ** 2, keep(d_g), tau))):
        assert got.n_corr == exp["n_corr"]
        assert list(got.n_inlier) == exp["n_inlier"]
        for k in ("mean", "rmse", "fitness", "sigma"):
            np.testing.assert_allclose(list(getattr(got, k)), exp[k], rtol=1e-12)
    np.testing.assert_allclose(res.full_cd, d_e.mean() + d_g.mean(), rtol=1e-12)
    np.testing.assert_allclose(list(res.cd), np.array(list(res.est_to_gt.rmse)) + np.array(list(res.gt_to_est.rmse)))


def test_eval_nn_as_written_pairing_swaps_lookup():
    """map_eval.cpp:1233 stores (nn_est, i_gt) but :1241 passes source=gt, target=est."""
    est, gt = _pair(5000)
    tau = [0.5, 0.3, 0.2, 0.1, 0.05]
    p = A.make_nn_params(tau, 1.0, pairing=A.ME_PAIRING_AS_WRITTEN)
    res, _, nn_g = O.eval_nn(est, gt, p, want_indices=True)
    n_gt, n_est = len(gt), len(est)
    a, b = nn_g.astype(np.int64), np.arange(n_gt)
    ok = (a < n_gt) & (b < n_est)
    d = gt[a[ok]] - est[b[ok]]
    nd = np.sqrt(d[:, 0] ** 2 + (d[:, 1] ** 2 + d[:, 2] ** 2))
    assert res.gt_to_est.n_ub == int((~ok).sum())
    assert res.gt_to_est.n_corr == int(ok.sum())
    assert list(res.gt_to_est.n_inlier) == [int((nd <= t).sum()) for t in tau]


def _numpy_mme(xyz, r, kmin):
    tree = cKDTree(xyz)
    ent = np.zeros(len(xyz))
    for i, nb in enumerate(tree.query_ball_point(xyz, r)):
        nb = np.array(nb)
        d = xyz[nb] - xyz[i]
        d2 = (d[:, 0] ** 2 + d[:, 1] ** 2) + d[:, 2] ** 2
        nb = nb[d2 < r * r]
        nb = nb[nb != i]
        if len(nb) < kmin:
            continue
        c = np.cov(xyz[nb].T)
        with np.errstate(all="ignore"):
            e = 0.5 * np.log(2 * np.pi * np.e * np.linalg.det(c))
        if np.isfinite(e):
            ent[i] = e
    return ent


@pytest.mark.parametrize("kmin", [10, 5])
def test_eval_mme_against_numpy(kmin):
    est, _ = _pair(6000)
    r = 0.1
    res, ent = O.eval_mme(est, r, kmin, want_entropies=True)
    exp = _numpy_mme(est, r, kmin)
    assert res.n_valid == int((exp != 0).sum())
    np.testing.assert_array_equal(ent != 0, exp != 0)
    np.testing.assert_allclose(ent, exp, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(res.mme, exp[exp != 0].mean(), rtol=1e-10)
    np.testing.assert_allclose(res.max_abs_entropy, abs(exp[exp != 0].min()), rtol=1e-10)
    np.testing.assert_allclose(res.min_abs_entropy, abs(exp[exp != 0].max()), rtol=1e-10)


def _numpy_w(mu1, sig1, n1, mu2, sig2, n2):
    def clamp(s, n):
        s = s.reshape(3, 3) / (n - 1)
        s = (s + s.T) / 2
        w, v = np.linalg.eigh(s)
        return v @ np.diag(np.maximum(w, 1e-6)) @ v.T
    s1, s2 = clamp(sig1, n1), clamp(sig2, n2)
    l1 = np.linalg.cholesky(s1)
    t = np.linalg.cholesky(l1 @ s2 @ l1.T)
    d = (mu1 - mu2) @ (mu1 - mu2) + np.trace(s1 + s2) - 2 * np.trace(t)
    return np.sqrt(max(0.0, d))


def test_voxel_gaussians_and_awd_against_numpy():
    est, gt = _pair(60000)
    v = 0.5
    keys, counts, mu, sigma = O.voxel_map(gt, v)
    kk = np.floor(gt / v).astype(np.int64)
    uk, inv = np.unique(kk, axis=0, return_inverse=True)
    assert len(uk) == len(keys)
    lut = {tuple(k): i for i, k in enumerate(uk)}
    for j in range(0, len(keys), 7):
        pts = gt[inv.reshape(-1) == lut[tuple(keys[j])]]
        n = len(pts)
        assert n == counts[j]
        np.testing.assert_allclose(mu[j], pts.mean(axis=0), rtol=1e-12)
        if n > 10:   # stored sigma = M2/(n-1)^2 (voxel_calculator.cpp:48 and :102)
            m2 = (pts - pts.mean(0)).T @ (pts - pts.mean(0))
            np.testing.assert_allclose(sigma[j].reshape(3, 3), m2 / (n - 1) ** 2, rtol=1e-9, atol=1e-18)
    res, rows = O.eval_awd(est, gt, v, min_points=100, scs_radius=5, want_rows=True)
    assert res.n_pairs == len(rows) and res.n_pairs > 20
    ws = []
    for r in rows:
        s_e = np.array([r[12], r[13], r[14], r[13], r[15], r[16], r[14], r[16], r[17]])
        s_g = np.array([r[21], r[22], r[23], r[22], r[24], r[25], r[23], r[25], r[26]])
        w = _numpy_w(r[18:21], s_g, int(r[10]), r[6:9], s_e, int(r[11]))
        np.testing.assert_allclose(r[9], w, rtol=1e-9)
        ws.append(w)
    np.testing.assert_allclose(res.awd, np.mean(ws), rtol=1e-12)
    keys_p = np.rint(rows[:, 0:3] / v).astype(np.int32)
    scs, cnt = O.scs(keys_p, rows[:, 9])
    np.testing.assert_allclose(res.scs, scs, rtol=1e-12)
    assert res.n_scs == cnt
    ek = np.unique(np.floor(est / v).astype(np.int64), axis=0)
    es, gs = {tuple(k) for k in ek}, {tuple(k) for k in uk}
    assert (res.n_active, res.n_old, res.n_new) == (len(es & gs), len(gs - es), len(es - gs))


def test_voxel_downsample_against_numpy():
    """oracle_voxel_downsample vs an independent numpy statement of Open3D's VoxelDownSample."""
    rng = np.random.RandomState(11)
    p = rng.rand(30000, 3) * np.array([4.0, 2.0, 1.0]) - np.array([100.0, 0.5, -7.0])
    for s in (0.02, 0.11, 0.9):
        o = O.voxel_downsample(p, s)
        org = p.min(0) - 0.5 * s
        k = np.floor((p - org) / s).astype(np.int64)
        u, inv = np.unique(k, axis=0, return_inverse=True)
        sums = np.zeros((len(u), 3))
        np.add.at(sums, inv.ravel(), p)          # np.add.at accumulates in input order, like AccumulatedPoint::AddPoint
        m = sums / np.bincount(inv.ravel())[:, None]
        assert o.shape == m.shape
        np.testing.assert_array_equal(o, m)      # np.unique sorts rows lexicographically = increasing (ix, iy, iz)


def test_icp_point_to_point_against_numpy_scipy():
    """oracle_icp_point_to_point (restating Open3D RegistrationICP + Eigen::umeyama) vs an independent numpy / cKDTree
    statement of the same loop."""
    from scipy.spatial import cKDTree
    from cloud_map_evaluation_b200 import synth
    est, gt, cfg = synth.make_pair("C1", scale=0.1)
    th = np.deg2rad(2.0)
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    c = gt.mean(0)
    src = (est - c) @ Rz.T + c + np.array([0.03, -0.02, 0.01])
    R = 0.2
    tree = cKDTree(gt)

    def step(pcd):
        d, i = tree.query(pcd, k=1)
        keep = d * d < R * R
        p, q = pcd[keep], gt[i[keep]]
        ms, md = p.mean(0), q.mean(0)
        U, s, Vt = np.linalg.svd(((q - md).T @ (p - ms)) / len(p))
        S = np.eye(3)
        if np.linalg.det(U) * np.linalg.det(Vt) < 0:
            S[2, 2] = -1
        T = np.eye(4)
        T[:3, :3] = U @ S @ Vt
        T[:3, 3] = md - T[:3, :3] @ ms
        return T, int(keep.sum()), float(np.sqrt((d[keep] ** 2).sum() / keep.sum()))

    T, pcd = np.eye(4), src.copy()
    for _ in range(4):
        upd, _, _ = step(pcd)
        T = upd @ T
        pcd = pcd @ upd[:3, :3].T + upd[:3, 3]
    _, nc, rm = step(pcd)
    To, fit, rmo, nco, ito = O.icp_point_to_point(src, gt, R, max_iter=4, rel_fitness=0.0, rel_rmse=0.0)
    assert ito == 4 and nco == nc
    np.testing.assert_allclose(To, T, atol=1e-13)
    np.testing.assert_allclose([fit, rmo], [nc / len(src), rm], rtol=1e-12)
