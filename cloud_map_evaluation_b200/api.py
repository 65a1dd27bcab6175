"""Python host mirror of the C-ABI: a thin context object used by the tests, bench.py and the multi-GPU plumbing.

Names follow the reference's members (`MapEval::computeMME`, `calculateMetricsWithInitialMatrix`, `calculateVMD`,
map_eval/src/map_eval.h:193-260) so the parity tests read like calls into the reference.  All computation happens
in libmapeval_b200.so on the GPU; this file only marshals arguments.
"""
import ctypes as C

import numpy as np

from . import _abi as A
from . import _lib


class MapEvalError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libmapeval_b200 error {code}: {msg}")
        self.code = code


class MapEvalB200:
    """One context = one GPU.  rank/world select this context's contiguous query range (multi-GPU sharding)."""

    def __init__(self, device=0, rank=0, world=1, stream=None, nn_cell_size=0.0, max_grid_cells=0,
                 vmd_voxel_size=0.0):
        self._L = _lib.load()
        opt = A.me_options()
        opt.abi_version = A.ME_ABI_VERSION
        opt.device, opt.rank, opt.world = int(device), int(rank), int(world)
        opt.stream = C.c_void_p(int(stream)) if stream else None
        opt.nn_cell_size = float(nn_cell_size)
        opt.max_grid_cells = int(max_grid_cells)
        opt.vmd_voxel_size = float(vmd_voxel_size)
        self._ctx = C.c_void_p()
        rc = self._L.me_create(C.byref(opt), C.byref(self._ctx))
        if rc != A.ME_OK:
            msg = self._L.me_last_error(None)
            self._ctx = None
            raise MapEvalError(rc, msg.decode() if msg else "me_create failed")
        self.n = [0, 0]
        self._keep = [None, None]   # keeps host/device buffers alive while the library may read them
        self.rank, self.world = int(rank), int(world)

    # -- lifecycle -------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._L.me_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != A.ME_OK:
            msg = self._L.me_last_error(self._ctx)
            raise MapEvalError(rc, msg.decode() if msg else "")

    def set_stream(self, stream_ptr):
        self._check(self._L.me_set_stream(self._ctx, C.c_void_p(int(stream_ptr)) if stream_ptr else None))

    def set_shard(self, rank, world):
        self._check(self._L.me_set_shard(self._ctx, rank, world))
        self.rank, self.world = int(rank), int(world)

    def synchronize(self):
        self._check(self._L.me_synchronize(self._ctx))

    # -- clouds ----------------------------------------------------------------------------------------------
    def set_cloud(self, which, xyz):
        """xyz: (N,3) float64 numpy array (host) or a pinned torch CPU tensor's data pointer via .numpy()."""
        a = np.ascontiguousarray(xyz, dtype=np.float64)
        if a.ndim != 2 or a.shape[1] != 3:
            raise ValueError("cloud must be (N, 3)")
        self._keep[which] = a
        self.n[which] = a.shape[0]
        self._check(self._L.me_set_cloud(self._ctx, which, a.ctypes.data_as(C.c_void_p), a.shape[0]))

    def set_cloud_ptr(self, which, host_ptr, n, keepalive=None):
        self._keep[which] = keepalive
        self.n[which] = int(n)
        self._check(self._L.me_set_cloud(self._ctx, which, C.c_void_p(int(host_ptr)), int(n)))

    def set_cloud_device(self, which, device_ptr, n, keepalive=None):
        """Borrow an fp64 (N,3) device buffer (e.g. a torch.cuda tensor's data_ptr())."""
        self._keep[which] = keepalive
        self.n[which] = int(n)
        self._check(self._L.me_set_cloud_device(self._ctx, which, C.c_void_p(int(device_ptr)), int(n)))

    def transform(self, which, T):
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(16)
        self._check(self._L.me_transform(self._ctx, which, T.ctypes.data_as(C.POINTER(C.c_double))))

    def performICPRegistration(self, max_correspondence_distance, T_init=None, max_iteration=30, relative_fitness=1e-6,
                               relative_rmse=1e-6, method=A.ME_ICP_POINT_TO_POINT):
        """map_eval.cpp:1366-1394: RegistrationICP with the point-to-point (method 0) or point-to-plane (1) estimation, or
        RegistrationGeneralizedICP (2, `registration_methods` of the config); the estimated cloud held by the context is
        transformed by the result.  Returns (T 4x4, me_icp_result)."""
        T = np.ascontiguousarray(np.eye(4) if T_init is None else T_init, dtype=np.float64).reshape(16)
        out = A.me_icp_result()
        self._check(self._L.me_icp(self._ctx, int(method), float(max_correspondence_distance), int(max_iteration),
                                   float(relative_fitness), float(relative_rmse),
                                   T.ctypes.data_as(C.POINTER(C.c_double)), C.byref(out)))
        self._keep[A.ME_CLOUD_EST] = None
        return np.array(list(out.transformation)).reshape(4, 4), out

    def set_normals(self, which, normals):
        a = np.ascontiguousarray(normals, dtype=np.float64)
        self._check(self._L.me_set_normals(self._ctx, which, a.ctypes.data_as(C.POINTER(C.c_double)), a.shape[0]))

    def estimate_normals(self, which, knn=20):
        """PointCloud::EstimateNormals(KDTreeSearchParamKNN(knn)); returns the (N, 3) normals in caller order."""
        self._check(self._L.me_estimate_normals(self._ctx, which, int(knn)))
        out = np.empty((self.n[which], 3), np.float64)
        self._check(self._L.me_get_normals(self._ctx, which, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def voxel_downsample(self, which, voxel_size):
        """PointCloud::VoxelDownSample on the held cloud (map_eval.cpp:38-39); returns the new point count."""
        n = C.c_int64(0)
        self._check(self._L.me_voxel_downsample(self._ctx, which, float(voxel_size), C.byref(n)))
        self.n[which] = n.value
        self._keep[which] = None
        return n.value

    def get_cloud(self, which):
        """The cloud currently held by the context as an (N, 3) float64 array."""
        n = C.c_int64(0)
        self._check(self._L.me_get_cloud(self._ctx, which, None, 0, C.byref(n)))
        out = np.empty((n.value, 3), np.float64)
        self._check(self._L.me_get_cloud(self._ctx, which, out.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)))
        return out

    def build_grid(self, which):
        self._check(self._L.me_build_grid(self._ctx, which))

    # -- a1/a2/a4 ----------------------------------------------------------------------------------------------
    def eval_nn_accum(self, params):
        e, g = A.me_nn_accum(), A.me_nn_accum()
        self._check(self._L.me_eval_nn_accum(self._ctx, C.byref(params), C.byref(e), C.byref(g)))
        return e, g

    def nn_finalize(self, params, e, g):
        out = A.me_nn_result()
        rc = self._L.me_nn_finalize(C.byref(params), C.byref(e), C.byref(g), self.n[0], self.n[1], C.byref(out))
        if rc != A.ME_OK:
            raise MapEvalError(rc, "me_nn_finalize")
        return out

    def calculateMetricsWithInitialMatrix(self, params):
        """map_eval.cpp:1204-1260 (single GPU)."""
        out = A.me_nn_result()
        self._check(self._L.me_eval_nn(self._ctx, C.byref(params), C.byref(out)))
        return out

    def get_nn(self, which_query):
        n = self.n[which_query]
        idx = np.empty(n, np.int32)
        d2 = np.empty(n, np.float64)
        self._check(self._L.me_get_nn(self._ctx, which_query, idx.ctypes.data_as(C.c_void_p),
                                      d2.ctypes.data_as(C.c_void_p)))
        return idx, d2

    # -- a5-a9 -------------------------------------------------------------------------------------------------
    def eval_mme_accum(self, which, radius, min_neighbors):
        acc = A.me_mme_accum()
        self._check(self._L.me_eval_mme_accum(self._ctx, which, float(radius), int(min_neighbors), C.byref(acc)))
        return acc

    def mme_finalize(self, acc, which):
        out = A.me_mme_result()
        rc = self._L.me_mme_finalize(C.byref(acc), self.n[which], C.byref(out))
        if rc != A.ME_OK:
            raise MapEvalError(rc, "me_mme_finalize")
        return out

    def computeMME(self, which, radius, min_neighbors, want_entropies=False):
        """map_eval.cpp:1608-1737 (min_neighbors=10) / :1438-1535 (min_neighbors=5), single GPU."""
        out = A.me_mme_result()
        ent = np.zeros(self.n[which], np.float64) if want_entropies else None
        self._check(self._L.me_eval_mme(self._ctx, which, float(radius), int(min_neighbors), C.byref(out),
                                        ent.ctypes.data_as(C.c_void_p) if want_entropies else None))
        return (out, ent) if want_entropies else out

    def get_entropies(self, which):
        ent = np.zeros(self.n[which], np.float64)
        self._check(self._L.me_get_entropies(self._ctx, which, ent.ctypes.data_as(C.c_void_p)))
        return ent

    # -- a10-a16 -----------------------------------------------------------------------------------------------
    def calculateVMD(self, voxel_size, min_points=100, scs_radius=5, want_rows=False):
        """map_eval.cpp:240-390."""
        out = A.me_awd_result()
        n_rows = C.c_int64(0)
        rows_p = C.POINTER(C.c_double)()
        self._check(self._L.me_eval_awd(self._ctx, float(voxel_size), int(min_points), int(scs_radius), C.byref(out),
                                        C.byref(n_rows), C.byref(rows_p) if want_rows else None))
        if not want_rows:
            return out
        rows = np.ctypeslib.as_array(rows_p, shape=(n_rows.value, 27)).copy() if n_rows.value > 0 \
            else np.zeros((0, 27))
        self._L.me_free(rows_p)
        return out, rows

    def awd_from_rows(self, rows27, voxel_size, scs_radius=5):
        """W per row + AWD / SCS from given voxel Gaussians (27-column voxel_errors.txt rows)."""
        rows = np.ascontiguousarray(rows27, dtype=np.float64)
        assert rows.ndim == 2 and rows.shape[1] == 27
        out = A.me_awd_result()
        w = np.empty(rows.shape[0], np.float64)
        dp = C.POINTER(C.c_double)
        self._check(self._L.me_awd_from_rows(self._ctx, rows.ctypes.data_as(dp), rows.shape[0], float(voxel_size), int(scs_radius),
                                             w.ctypes.data_as(dp), C.byref(out)))
        return out, w

    # -- device-resident accumulators (multi-GPU passes without host round trips) ---------------------------------
    def accum_reset(self):
        self._check(self._L.me_accum_reset(self._ctx))

    def eval_nn_accum_device(self, params):
        self._check(self._L.me_eval_nn_accum_device(self._ctx, C.byref(params)))

    def eval_mme_accum_device(self, which, radius, min_neighbors):
        self._check(self._L.me_eval_mme_accum_device(self._ctx, which, float(radius), int(min_neighbors)))

    def accum_block(self):
        """(device pointer, n_sum, n_max) of the fp64 accumulator block"""
        ptr = C.POINTER(C.c_double)()
        ns, nm = C.c_int32(0), C.c_int32(0)
        self._check(self._L.me_accum_block(self._ctx, C.byref(ptr), C.byref(ns), C.byref(nm)))
        return C.cast(ptr, C.c_void_p).value, ns.value, nm.value

    def accum_fetch(self, want_mme=(True, False)):
        e, g = A.me_nn_accum(), A.me_nn_accum()
        mm = [A.me_mme_accum() if w else None for w in want_mme]
        self._check(self._L.me_accum_fetch(self._ctx, C.byref(e), C.byref(g), C.byref(mm[0]) if mm[0] else None,
                                           C.byref(mm[1]) if mm[1] else None))
        return e, g, [m for m in mm if m is not None]

    # -- layouts of a multi-GPU job; the voxel stage in two halves ------------------------------------------------
    def set_layout(self, layout):
        """A.ME_LAYOUT_REPLICATED (default) or A.ME_LAYOUT_SLAB: every rank lays out only the voxel layers it owns."""
        self._check(self._L.me_set_layout(self._ctx, int(layout)))

    def layout_active(self):
        """{'layout', 'axis', 'n_laid_out': [est, gt], 'n_owned': [est, gt]} once the lattices are built"""
        lay, ax = C.c_int32(0), C.c_int32(0)
        nl, no = (C.c_int64 * 2)(), (C.c_int64 * 2)()
        self._check(self._L.me_layout_active(self._ctx, C.byref(lay), C.byref(ax), nl, no))
        return {"layout": lay.value, "axis": ax.value, "n_laid_out": list(nl), "n_owned": list(no)}

    def voxel_begin(self, voxel_size, min_points=100):
        self._check(self._L.me_voxel_begin(self._ctx, float(voxel_size), int(min_points)))

    def voxel_w_table(self):
        """(device pointer, n) of the W table over the estimated cloud's voxels (-1 = no pair)"""
        ptr = C.POINTER(C.c_double)()
        n = C.c_int64(0)
        self._check(self._L.me_voxel_w_table(self._ctx, C.byref(ptr), C.byref(n)))
        return C.cast(ptr, C.c_void_p).value, n.value

    def voxel_finish_accum_device(self, scs_radius=5):
        self._check(self._L.me_voxel_finish_accum_device(self._ctx, int(scs_radius)))

    def accum_fetch_awd(self):
        out = A.me_awd_result()
        self._check(self._L.me_accum_fetch_awd(self._ctx, C.byref(out)))
        return out

    # -- introspection -----------------------------------------------------------------------------------------
    def stage_times_ms(self):
        ms = (C.c_double * A.ME_N_STAGE_TIMES)()
        self._check(self._L.me_get_stage_times(self._ctx, ms))
        return dict(zip(A.STAGE_NAMES, list(ms)))

    def launch_count(self):
        return int(self._L.me_launch_count(self._ctx))
