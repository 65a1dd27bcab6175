"""ctypes mirror of include/mapeval_b200.h (struct layouts and constants only — no library is loaded here)."""
import ctypes as C

ME_ABI_VERSION = 1

ME_OK = 0
ME_ERR_INVALID = -1
ME_ERR_NO_DEVICE = -2
ME_ERR_CUDA = -3
ME_ERR_NOMEM = -4
ME_ERR_EMPTY = -5
ME_ERR_RANGE = -6

ME_CLOUD_EST = 0
ME_CLOUD_GT = 1

ME_CUTOFF_SQDIST_LE_R = 0
ME_CUTOFF_DIST_LT_R = 1

ME_PAIRING_AS_WRITTEN = 0
ME_PAIRING_GEOMETRIC = 1

class me_lattice_plan(C.Structure):
    _fields_ = [("v", C.c_double), ("h", C.c_double), ("m", C.c_int32), ("sparse", C.c_int32), ("nvox", C.c_int32 * 3),
                ("dims", C.c_int32 * 3), ("ncells", C.c_int64)]


ME_LAYOUT_REPLICATED = 0
ME_LAYOUT_SLAB = 1
ME_ICP_POINT_TO_POINT = 0
ME_ICP_POINT_TO_PLANE = 1
ME_ICP_GENERALIZED = 2

ME_N_STAGE_TIMES = 9
STAGE_NAMES = ("grid_est", "grid_gt", "nn_est_to_gt", "nn_gt_to_est", "mme_est", "mme_gt",
               "voxel_moments", "awd", "scs")


class me_options(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("stream", C.c_void_p), ("nn_cell_size", C.c_double), ("max_grid_cells", C.c_int64),
                ("vmd_voxel_size", C.c_double)]


class me_nn_params(C.Structure):
    _fields_ = [("tau", C.c_double * 5), ("icp_max_distance", C.c_double), ("cutoff_mode", C.c_int32),
                ("pairing", C.c_int32), ("want_full_cd", C.c_int32), ("directions", C.c_int32)]


ME_NN_ACCUM_I64 = 9
ME_NN_ACCUM_F64 = 13


class me_nn_accum(C.Structure):
    _fields_ = [("n_query", C.c_int64), ("n_corr", C.c_int64), ("n_inlier", C.c_int64 * 5), ("n_ub", C.c_int64),
                ("n_far", C.c_int64),
                ("sum_d", C.c_double * 5), ("sum_d2", C.c_double * 5), ("sum_d_all", C.c_double),
                ("sum_d2_all", C.c_double), ("sum_nn_dist", C.c_double)]


class me_dir_result(C.Structure):
    _fields_ = [("n_source", C.c_int64), ("n_corr", C.c_int64), ("n_inlier", C.c_int64 * 5), ("n_ub", C.c_int64),
                ("mean", C.c_double * 5), ("rmse", C.c_double * 5), ("fitness", C.c_double * 5),
                ("sigma", C.c_double * 5), ("sum_nn_dist", C.c_double)]


class me_nn_result(C.Structure):
    _fields_ = [("est_to_gt", me_dir_result), ("gt_to_est", me_dir_result), ("cd", C.c_double * 5),
                ("f1", C.c_double * 5), ("iou", C.c_double * 5), ("full_cd", C.c_double)]


class me_mme_accum(C.Structure):
    _fields_ = [("n_query", C.c_int64), ("n_valid", C.c_int64), ("sum_entropy", C.c_double),
                ("min_entropy", C.c_double), ("max_entropy", C.c_double)]


class me_mme_result(C.Structure):
    _fields_ = [("mme", C.c_double), ("n_valid", C.c_int64), ("n_total", C.c_int64),
                ("min_abs_entropy", C.c_double), ("max_abs_entropy", C.c_double)]


class me_awd_result(C.Structure):
    _fields_ = [("awd", C.c_double), ("scs", C.c_double), ("n_pairs", C.c_int64), ("n_scs", C.c_int64),
                ("n_voxels_est", C.c_int64), ("n_voxels_gt", C.c_int64), ("n_active", C.c_int64),
                ("n_old", C.c_int64), ("n_new", C.c_int64)]


class me_icp_result(C.Structure):
    _fields_ = [("transformation", C.c_double * 16), ("fitness", C.c_double), ("inlier_rmse", C.c_double),
                ("n_corr", C.c_int64), ("iterations", C.c_int32), ("converged", C.c_int32)]


def make_nn_params(tau, icp_max_distance=1.0, cutoff_mode=ME_CUTOFF_SQDIST_LE_R, pairing=ME_PAIRING_AS_WRITTEN,
                   want_full_cd=True, directions=0):
    p = me_nn_params()
    if len(tau) < 5:
        raise ValueError("accuracy_level needs 5 thresholds (map_eval_main.cpp:133-137)")
    for i in range(5):
        p.tau[i] = float(tau[i])
    p.icp_max_distance = float(icp_max_distance)
    p.cutoff_mode = int(cutoff_mode)
    p.pairing = int(pairing)
    p.want_full_cd = 1 if want_full_cd else 0
    p.directions = int(directions)
    return p


def struct_to_dict(s):
    """Recursively convert a ctypes struct into plain python (lists for arrays)."""
    out = {}
    for name, _ in s._fields_:
        v = getattr(s, name)
        if isinstance(v, C.Structure):
            out[name] = struct_to_dict(v)
        elif isinstance(v, C.Array):
            out[name] = list(v)
        else:
            out[name] = v
    return out
