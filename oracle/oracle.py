"""ctypes binding of oracle/liboracle_mapeval.so — the CPU restatement of the reference.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under cloud_map_evaluation_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cloud_map_evaluation_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_mapeval.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "mapeval_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle_mapeval.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_transform.argtypes = [dp, C.c_int64, dp]
        L.oracle_knn1.argtypes = [dp, C.c_int64, dp, C.c_int64, ip, dp, C.c_int]
        L.oracle_eval_nn.argtypes = [dp, C.c_int64, dp, C.c_int64, C.POINTER(A.me_nn_params),
                                     C.POINTER(A.me_nn_result), ip, ip, C.c_int]
        L.oracle_eval_mme.argtypes = [dp, C.c_int64, C.c_double, C.c_int32, C.POINTER(A.me_mme_result), dp, C.c_int]
        L.oracle_eval_awd.argtypes = [dp, C.c_int64, dp, C.c_int64, C.c_double, C.c_int32, C.c_int32,
                                      C.POINTER(A.me_awd_result), C.POINTER(C.c_int64), C.POINTER(dp)]
        L.oracle_free.argtypes = [C.c_void_p]
        L.oracle_set_voxel_hash.argtypes = [C.c_int]
        L.oracle_set_voxel_hash.restype = None
        L.oracle_wasserstein.restype = C.c_double
        L.oracle_wasserstein.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int]
        L.oracle_scs.argtypes = [ip, dp, C.c_int64, C.c_int, dp, C.POINTER(C.c_int64)]
        L.oracle_voxel_map.argtypes = [dp, C.c_int64, C.c_double, C.POINTER(C.c_int64), C.POINTER(ip),
                                       C.POINTER(ip), C.POINTER(dp), C.POINTER(dp)]
        L.oracle_icp_point_to_point.argtypes = [dp, C.c_int64, dp, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double, dp, dp,
                                                dp, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.oracle_voxel_downsample.argtypes = [dp, C.c_int64, C.c_double, C.POINTER(C.c_int64), C.POINTER(dp)]
        L.oracle_estimate_normals_knn.argtypes = [dp, C.c_int64, C.c_int, dp]
        L.oracle_gicp_covariance.argtypes = [dp, C.c_double, dp]
        L.oracle_icp_generalized.argtypes = [dp, C.c_int64, dp, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double, dp, dp,
                                             dp, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.oracle_icp_point_to_plane.argtypes = [dp, C.c_int64, dp, C.c_int64, dp, C.c_double, C.c_int, C.c_double, C.c_double, dp,
                                                dp, dp, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2 and a.shape[1] == 3
    return a


def num_threads():
    return lib().oracle_num_threads()


def transform(xyz, T):
    out = _cloud(xyz).copy()
    T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
    lib().oracle_transform(_dptr(out), out.shape[0], _dptr(T))
    return out


def knn1(query, ref, threads=0):
    q, r = _cloud(query), _cloud(ref)
    idx = np.empty(q.shape[0], np.int32)
    d2 = np.empty(q.shape[0], np.float64)
    lib().oracle_knn1(_dptr(q), q.shape[0], _dptr(r), r.shape[0], _iptr(idx), _dptr(d2), threads)
    return idx, d2


def eval_nn(est, gt, params, threads=0, want_indices=False):
    e, g = _cloud(est), _cloud(gt)
    res = A.me_nn_result()
    ie = np.empty(e.shape[0], np.int32) if want_indices else None
    ig = np.empty(g.shape[0], np.int32) if want_indices else None
    rc = lib().oracle_eval_nn(_dptr(e), e.shape[0], _dptr(g), g.shape[0], C.byref(params), C.byref(res),
                              _iptr(ie) if want_indices else None, _iptr(ig) if want_indices else None, threads)
    if rc != 0:
        raise RuntimeError(f"oracle_eval_nn failed: {rc}")
    return (res, ie, ig) if want_indices else res


def eval_mme(xyz, radius, min_neighbors, threads=0, want_entropies=False):
    p = _cloud(xyz)
    res = A.me_mme_result()
    ent = np.zeros(p.shape[0], np.float64)
    rc = lib().oracle_eval_mme(_dptr(p), p.shape[0], float(radius), int(min_neighbors), C.byref(res), _dptr(ent),
                               threads)
    if rc != 0:
        raise RuntimeError(f"oracle_eval_mme failed: {rc}")
    return (res, ent) if want_entropies else res


def eval_awd(est, gt, voxel_size, min_points=100, scs_radius=5, want_rows=False):
    e, g = _cloud(est), _cloud(gt)
    res = A.me_awd_result()
    n_rows = C.c_int64(0)
    rows_p = C.POINTER(C.c_double)()
    rc = lib().oracle_eval_awd(_dptr(e), e.shape[0], _dptr(g), g.shape[0], float(voxel_size), int(min_points),
                               int(scs_radius), C.byref(res), C.byref(n_rows),
                               C.byref(rows_p) if want_rows else None)
    if rc != 0:
        raise RuntimeError(f"oracle_eval_awd failed: {rc}")
    if not want_rows:
        return res
    rows = np.ctypeslib.as_array(rows_p, shape=(max(n_rows.value, 0), 27)).copy() if n_rows.value > 0 \
        else np.zeros((0, 27))
    lib().oracle_free(rows_p)
    return res, rows


def set_voxel_hash(fast):
    """False (default): the reference's XOR voxel hash; True: a mixing hash — same voxel Gaussians, another iteration
    order, ~100x faster on large clouds.  Used by bench.py's timed CPU baseline only."""
    lib().oracle_set_voxel_hash(1 if fast else 0)


def wasserstein(mu1, sigma1, n1, mu2, sigma2, n2):
    a = [np.ascontiguousarray(x, dtype=np.float64).reshape(-1) for x in (mu1, sigma1, mu2, sigma2)]
    return lib().oracle_wasserstein(_dptr(a[0]), _dptr(a[1]), int(n1), _dptr(a[2]), _dptr(a[3]), int(n2))


def scs(keys3, w, radius=5):
    k = np.ascontiguousarray(keys3, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = C.c_double(0)
    cnt = C.c_int64(0)
    lib().oracle_scs(_iptr(k), _dptr(w), k.shape[0], int(radius), C.byref(out), C.byref(cnt))
    return out.value, cnt.value


def voxel_map(xyz, voxel_size):
    p = _cloud(xyz)
    nv = C.c_int64(0)
    kp, cp = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
    mp, sp = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
    lib().oracle_voxel_map(_dptr(p), p.shape[0], float(voxel_size), C.byref(nv), C.byref(kp), C.byref(cp),
                           C.byref(mp), C.byref(sp))
    n = nv.value
    keys = np.ctypeslib.as_array(kp, shape=(n, 3)).copy()
    counts = np.ctypeslib.as_array(cp, shape=(n,)).copy()
    mu = np.ctypeslib.as_array(mp, shape=(n, 3)).copy()
    sigma = np.ctypeslib.as_array(sp, shape=(n, 9)).copy()
    for ptr in (kp, cp, mp, sp):
        lib().oracle_free(ptr)
    return keys, counts, mu, sigma


def voxel_downsample(xyz, voxel_size):
    """open3d PointCloud::VoxelDownSample (map_eval.cpp:38-39); output voxels in increasing (ix, iy, iz)."""
    a = _cloud(xyz)
    n_out = C.c_int64(0)
    out = C.POINTER(C.c_double)()
    rc = lib().oracle_voxel_downsample(_dptr(a), a.shape[0], float(voxel_size), C.byref(n_out), C.byref(out))
    if rc != 0:
        raise ValueError("oracle_voxel_downsample failed")
    res = np.ctypeslib.as_array(out, shape=(n_out.value, 3)).copy()
    lib().oracle_free(out)
    return res


def icp_point_to_point(est, gt, max_dist, T_init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """open3d RegistrationICP + TransformationEstimationPointToPoint (map_eval.cpp:1370-1374).
    Returns (T 4x4, fitness, inlier_rmse, n_corr, iterations)."""
    e, g = _cloud(est), _cloud(gt)
    Ti = np.ascontiguousarray(np.eye(4) if T_init is None else T_init, dtype=np.float64).reshape(16)
    To = np.empty(16, np.float64)
    fit, rm = C.c_double(0), C.c_double(0)
    nc, it = C.c_int64(0), C.c_int32(0)
    lib().oracle_icp_point_to_point(_dptr(e), e.shape[0], _dptr(g), g.shape[0], float(max_dist), int(max_iter), float(rel_fitness),
                                    float(rel_rmse), _dptr(Ti), _dptr(To), C.byref(fit), C.byref(rm), C.byref(nc), C.byref(it))
    return To.reshape(4, 4), fit.value, rm.value, nc.value, it.value


def estimate_normals_knn(xyz, knn=20):
    """open3d PointCloud::EstimateNormals(KDTreeSearchParamKNN(knn)) on a cloud without normals [ext-gicp]."""
    a = _cloud(xyz)
    out = np.empty_like(a)
    lib().oracle_estimate_normals_knn(_dptr(a), a.shape[0], int(knn), _dptr(out))
    return out


def gicp_covariance(normal, eps=1e-3):
    n = np.ascontiguousarray(normal, dtype=np.float64).reshape(3)
    out = np.empty(9, np.float64)
    lib().oracle_gicp_covariance(_dptr(n), float(eps), _dptr(out))
    return out.reshape(3, 3)


def _icp_call(fn, e, g, extra, max_dist, T_init, max_iter, rel_fitness, rel_rmse):
    Ti = np.ascontiguousarray(np.eye(4) if T_init is None else T_init, dtype=np.float64).reshape(16)
    To = np.empty(16, np.float64)
    fit, rm = C.c_double(0), C.c_double(0)
    nc, it = C.c_int64(0), C.c_int32(0)
    rc = fn(_dptr(e), e.shape[0], _dptr(g), g.shape[0], *extra, float(max_dist), int(max_iter), float(rel_fitness),
            float(rel_rmse), _dptr(Ti), _dptr(To), C.byref(fit), C.byref(rm), C.byref(nc), C.byref(it))
    if rc != 0:
        raise RuntimeError(f"oracle ICP failed: {rc}")
    return To.reshape(4, 4), fit.value, rm.value, nc.value, it.value


def icp_generalized(est, gt, max_dist, T_init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """open3d RegistrationGeneralizedICP + TransformationEstimationForGeneralizedICP() (map_eval.cpp:1381-1385).
    Returns (T 4x4, fitness, inlier_rmse, n_corr, iterations)."""
    return _icp_call(lib().oracle_icp_generalized, _cloud(est), _cloud(gt), (), max_dist, T_init, max_iter, rel_fitness, rel_rmse)


def icp_point_to_plane(est, gt, gt_normals, max_dist, T_init=None, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """open3d RegistrationICP + TransformationEstimationPointToPlane (map_eval.cpp:1375-1379); the target needs normals."""
    nr = _cloud(gt_normals)
    return _icp_call(lib().oracle_icp_point_to_plane, _cloud(est), _cloud(gt), (_dptr(nr),), max_dist, T_init, max_iter,
                     rel_fitness, rel_rmse)
