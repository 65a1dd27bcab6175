"""The oracle's restatement of Open3D's generalized ICP and point-to-plane ICP (MapEval::performICPRegistration cases 2
and 1, map_eval.cpp:1375-1386) against an independent numpy / scipy statement of the same equations: neighbours from
scipy.spatial.cKDTree, normals from numpy.linalg.eigh of the centred covariance, the normal equations in the M^-1 form
(the oracle builds W = M^-1/2 as Open3D does), numpy.linalg.solve.  The reference holds no registration fixtures, so
this boundary is unpinned; what is checked here is that the two statements of the published algorithm agree."""
import numpy as np
import pytest
from scipy.spatial import cKDTree

from oracle import oracle as O


def _rot(rx, ry, rz):
    ca, sa, cb, sb, cg, sg = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rz = np.array([[cg, -sg, 0], [sg, cg, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def _scene(n, seed, noise):
    """a bumpy floor and a wall whose normal is close to +y: no normal comes near -x (the e1 quirk has its own test)"""
    rs = np.random.RandomState(seed)
    m = n // 3
    x, y = rs.rand(n - m) * 6, rs.rand(n - m) * 6
    floor = np.stack([x, y, 0.15 * np.sin(1.3 * x) * np.cos(0.9 * y)], 1)
    u, v = rs.rand(m) * 6, rs.rand(m) * 2.5
    wall = np.stack([u, 6.0 + 0.1 * np.sin(u), v], 1)
    p = np.concatenate([floor, wall]) + noise * rs.randn(n, 3)
    return p.astype(np.float32).astype(np.float64)


def _np_normals(xyz, k=20):
    _, idx = cKDTree(xyz).query(xyz, k=k)
    nb = xyz[idx]                                   # (n, k, 3)
    c = nb - nb.mean(1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / k
    w, v = np.linalg.eigh(cov)
    return v[:, :, 0]                               # eigenvector of the smallest eigenvalue (sign arbitrary)


def _np_cov_from_normals(nr, eps=1e-3):
    return np.eye(3)[None] + (eps - 1.0) * nr[:, :, None] * nr[:, None, :]      # = Rx diag(eps,1,1) Rx^T away from the quirk


def _skew(v):
    z = np.zeros(len(v))
    return np.stack([np.stack([z, -v[:, 2], v[:, 1]], 1), np.stack([v[:, 2], z, -v[:, 0]], 1), np.stack([-v[:, 1], v[:, 0], z], 1)], 1)


def _np_icp(est, gt, R, T0, method, gt_normals=None, max_iter=30):
    tree = cKDTree(gt)
    T = T0.copy()
    pcd = est @ T[:3, :3].T + T[:3, 3]
    if method == 2:
        Cs = _np_cov_from_normals(_np_normals(est))
        Ct = _np_cov_from_normals(_np_normals(gt))
        Cs = T[:3, :3] @ Cs @ T[:3, :3].T

    def evaluate(p):
        d, j = tree.query(p, k=1)
        keep = d * d < R * R
        return keep, j, keep.mean(), (np.sqrt(np.mean(d[keep] ** 2)) if keep.any() else 0.0)
    keep, j, fit, rmse = evaluate(pcd)
    it = 0
    while it < max_iter:
        vs, vt = pcd[keep], gt[j[keep]]
        d = vs - vt
        if method == 2:
            Minv = np.linalg.inv(Ct[j[keep]] + Cs[keep])
            A = np.concatenate([-_skew(vs), np.repeat(np.eye(3)[None], len(vs), 0)], 2)      # (m, 3, 6)
            JTJ = np.einsum("mia,mij,mjb->ab", A, Minv, A)
            JTr = np.einsum("mia,mij,mj->a", A, Minv, d)
        else:
            nt = gt_normals[j[keep]]
            J = np.concatenate([np.cross(vs, nt), nt], 1)
            r = np.einsum("mi,mi->m", d, nt)
            JTJ, JTr = J.T @ J, J.T @ r
        x = np.linalg.solve(JTJ, -JTr)
        upd = np.eye(4)
        upd[:3, :3] = _rot(x[0], x[1], x[2])
        upd[:3, 3] = x[3:]
        T = upd @ T
        pcd = pcd @ upd[:3, :3].T + upd[:3, 3]
        if method == 2:
            Cs = upd[:3, :3] @ Cs @ upd[:3, :3].T
        f0, r0 = fit, rmse
        keep, j, fit, rmse = evaluate(pcd)
        it += 1
        if abs(f0 - fit) < 1e-6 and abs(r0 - rmse) < 1e-6:
            break
    return T, fit, rmse, int(keep.sum()), it


def test_knn_normals_match_numpy():
    xyz = _scene(6000, 1, 0.004)
    got = O.estimate_normals_knn(xyz, 20)
    exp = _np_normals(xyz, 20)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, rtol=1e-12)
    dots = np.abs(np.einsum("ni,ni->n", got, exp))
    assert np.all(dots > 1 - 1e-9)          # same line (the sign is the solver's business)


def test_gicp_covariance_and_the_e1_quirk():
    rs = np.random.RandomState(3)
    for _ in range(50):
        n = rs.randn(3)
        n /= np.linalg.norm(n)
        if n[0] < -0.99:
            continue
        np.testing.assert_allclose(O.gicp_covariance(n), np.eye(3) + (1e-3 - 1) * np.outer(n, n), atol=1e-14)
    n = np.array([-0.995, 0.0, np.sqrt(1 - 0.995 ** 2)])           # c < -0.99: Rx = I -> diag(eps, 1, 1) whatever n is
    np.testing.assert_array_equal(O.gicp_covariance(n), np.diag([1e-3, 1.0, 1.0]))
    np.testing.assert_allclose(O.gicp_covariance(-n), np.eye(3) + (1e-3 - 1) * np.outer(n, n), atol=1e-14)


@pytest.mark.parametrize("method", [2, 1])
def test_registration_matches_numpy(method):
    gt = _scene(9000, 11, 0.002)
    T_true = np.eye(4)
    T_true[:3, :3] = _rot(0.012, -0.008, 0.015)
    T_true[:3, 3] = [0.03, -0.02, 0.015]
    est = _scene(7000, 12, 0.006)
    est = (est - T_true[:3, 3]) @ T_true[:3, :3]           # est = T_true^-1 (points): the registration should recover T_true
    est = est.astype(np.float32).astype(np.float64)
    T0 = np.eye(4)
    T0[:3, 3] = [0.004, 0.0, -0.003]
    if method == 2:
        T, fit, rmse, nc, it = O.icp_generalized(est, gt, 0.5, T0)
        Tn, fn, rn, ncn, itn = _np_icp(est, gt, 0.5, T0, 2)
    else:
        nrm = _np_normals(gt)
        T, fit, rmse, nc, it = O.icp_point_to_plane(est, gt, nrm, 0.5, T0)
        Tn, fn, rn, ncn, itn = _np_icp(est, gt, 0.5, T0, 1, gt_normals=nrm)
    assert it == itn and nc == ncn and 2 <= it <= 30
    np.testing.assert_allclose(T, Tn, rtol=0, atol=1e-9)
    np.testing.assert_allclose([fit, rmse], [fn, rn], rtol=1e-9)
    np.testing.assert_allclose(T[:3, 3], T_true[:3, 3], atol=5e-3)            # and it actually registers
    np.testing.assert_allclose(T[:3, :3], T_true[:3, :3], atol=2e-3)
