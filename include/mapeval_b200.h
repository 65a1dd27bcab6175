/*
 * mapeval_b200.h — C-ABI of libmapeval_b200.so
 *
 * B200-native (sm_100a) replacement for the metric hot path of MapEval
 * (JokerJohn/Cloud_Map_Evaluation).  The reference has no FFI of its own: the
 * path sits behind C++ member functions of `class MapEval` that talk through
 * member state (map_eval/src/map_eval.h:121-362).  Each entry point below names
 * the reference member function(s) it replaces (paths relative to the reference
 * root, `map_eval/src/`).  Plain C: pointers, sizes, PODs.  No C++/torch types.
 *
 * Conventions
 *   - every function returns ME_OK (0) or a negative me_status; after an error
 *     me_last_error(ctx) holds a message (the reference's convention is the
 *     0 / -1 return of MapEval::process(), map_eval.cpp:16,28,34,101).
 *   - clouds are N x 3 fp64, array-of-structs — the layout of Open3D's
 *     `std::vector<Eigen::Vector3d> points_` that the reference reads.
 *   - one me_ctx = one GPU = one caller thread at a time.  Multi-GPU = one
 *     context per process/GPU with {rank, world} set; the *_accum entry points
 *     return sum-reducible partial accumulators for this rank's query range,
 *     the caller all-reduces them (NCCL) and calls the host-side *_finalize.
 *   - there is NO CPU fallback in this library: without a usable CUDA device
 *     me_create fails with ME_ERR_NO_DEVICE.
 */
#ifndef MAPEVAL_B200_H_
#define MAPEVAL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ME_ABI_VERSION 1

#if defined(__GNUC__)
#define ME_API __attribute__((visibility("default")))
#else
#define ME_API
#endif

typedef enum {
  ME_OK = 0,
  ME_ERR_INVALID = -1,     /* bad argument / call order                       */
  ME_ERR_NO_DEVICE = -2,   /* no CUDA device / wrong architecture             */
  ME_ERR_CUDA = -3,        /* a CUDA runtime call failed (see me_last_error)  */
  ME_ERR_NOMEM = -4,       /* device or host allocation failed                */
  ME_ERR_EMPTY = -5,       /* a required cloud is empty (map_eval.cpp:32-35)  */
  ME_ERR_RANGE = -6        /* coordinates / voxel indices outside what the grids can index */
} me_status;

enum { ME_CLOUD_EST = 0, ME_CLOUD_GT = 1 };

/* kept-pair cut-off of the 1-NN sweep */
enum {
  ME_CUTOFF_SQDIST_LE_R = 0,  /* keep iff d2 <= icp_max_distance  (path A as written, map_eval.cpp:1219,1232) */
  ME_CUTOFF_DIST_LT_R = 1     /* keep iff d2 <  icp_max_distance^2 (Open3D EvaluateRegistration, map_eval.cpp:1168) */
};

/* how the gt->est pairs are looked up by the accumulator pass */
enum {
  ME_PAIRING_AS_WRITTEN = 0,  /* map_eval.cpp:1233 stores (nn_est, i_gt) but :1241 passes (source=gt,target=est),
                                 so :1093-1094 reads gt[nn_est] and est[i_gt]; reproduced verbatim.  Pairs whose
                                 swapped indices fall outside the clouds (undefined behaviour in the reference)
                                 are dropped and counted in n_ub. */
  ME_PAIRING_GEOMETRIC = 1    /* distance between the gt point and its nearest est point */
};

typedef struct me_ctx me_ctx;

typedef struct {
  int32_t abi_version;     /* must be ME_ABI_VERSION                                       */
  int32_t device;          /* CUDA device ordinal                                          */
  int32_t rank, world;     /* this context evaluates queries [rank*N/world,(rank+1)*N/world) */
  void   *stream;          /* cudaStream_t to launch on; NULL = library-owned stream       */
  double  nn_cell_size;    /* edge of the lattice cells in metres; <= 0 = auto         */
  int64_t max_grid_cells;  /* budget for the dense cell table; <= 0 = automatic (2^28 .. 2^31 with the cloud sizes) */
  double  vmd_voxel_size;  /* voxel edge the lattices should align to (the config's vmd_voxel_size);
                              <= 0 = unknown: me_eval_awd then re-lays the clouds out once     */
} me_options;

/* ---- a1/a2/a4: AC / COM / CD inlier statistics and full Chamfer ------------------------------- */

typedef struct {
  double  tau[5];            /* accuracy_level / trunc_dist_ (map_eval.h:85)               */
  double  icp_max_distance;  /* R (map_eval.h:69)                                          */
  int32_t cutoff_mode;       /* ME_CUTOFF_*                                                */
  int32_t pairing;           /* ME_PAIRING_* (gt->est direction only)                      */
  int32_t want_full_cd;      /* also accumulate sum sqrt(d2) over ALL queries (map_eval.cpp:1398-1431) */
  int32_t directions;        /* bit0: est->gt, bit1: gt->est; 0 = both                     */
} me_nn_params;

/* sum-reducible partial accumulators of one direction (int64 block first, then fp64 block) */
#define ME_NN_ACCUM_I64 9
#define ME_NN_ACCUM_F64 13
typedef struct {
  int64_t n_query;           /* queries evaluated by this rank                             */
  int64_t n_corr;            /* kept pairs |C|                                             */
  int64_t n_inlier[5];       /* pairs with d <= tau_k (number_vec, map_eval.cpp:1099-1123) */
  int64_t n_ub;              /* ME_PAIRING_AS_WRITTEN: dropped out-of-range pairs          */
  int64_t n_far;             /* queries resolved by the far (ring-expansion) kernel        */
  double  sum_d[5];          /* sum of d   over pairs with d <= tau_k (mean_vec)           */
  double  sum_d2[5];         /* sum of d^2 over pairs with d <= tau_k (rmse_vec)           */
  double  sum_d_all;         /* sum of d   over all kept pairs  (sigma closed form)        */
  double  sum_d2_all;        /* sum of d^2 over all kept pairs                             */
  double  sum_nn_dist;       /* sum of sqrt(d2_nn) over ALL queries (full CD)              */
} me_nn_accum;

typedef struct {             /* == est_gt_results / gt_est_results (map_eval.cpp:1140-1144) */
  int64_t n_source, n_corr;
  int64_t n_inlier[5];
  int64_t n_ub;
  double  mean[5], rmse[5], fitness[5], sigma[5];
  double  sum_nn_dist;
} me_dir_result;

typedef struct {
  me_dir_result est_to_gt, gt_to_est;
  double cd[5], f1[5], iou[5];   /* map_eval.cpp:1245-1253 */
  double full_cd;                /* map_eval.cpp:1429      */
} me_nn_result;

/* ---- a5-a9: mean map entropy -------------------------------------------------------------------- */

typedef struct {
  int64_t n_query, n_valid;
  double  sum_entropy;
  double  min_entropy, max_entropy;  /* over entropies != 0 (map_eval.cpp:697-701); +inf / -inf if none */
} me_mme_accum;

typedef struct {
  double  mme;                       /* sum/n_valid, 0 if none (map_eval.cpp:1720-1724) */
  int64_t n_valid, n_total;
  double  min_abs_entropy, max_abs_entropy;  /* |max|, |min| (map_eval.cpp:700-701); NaN if no non-zero entropy */
} me_mme_result;

/* ---- a10-a16: voxel Gaussians, AWD ("VMD"), SCS ------------------------------------------------ */

typedef struct {
  double  awd, scs;                  /* vmd (map_eval.cpp:324-325), scs_overall (:387); NaN when the reference divides 0/0 */
  int64_t n_pairs, n_scs;            /* |wasserstein_distances|, scs_count */
  int64_t n_voxels_est, n_voxels_gt; /* voxel_map_.size() after buildVoxelMap (voxel_calculator.cpp:55) */
  int64_t n_active, n_old, n_new;    /* voxel_calculator.cpp:170 */
} me_awd_result;

/* ---- N2: registration (registration_methods: 0 point-to-point, 1 point-to-plane, 2 generalized ICP) --------- */
#define ME_ICP_POINT_TO_POINT 0     /* map_eval.cpp:1370-1374 */
#define ME_ICP_POINT_TO_PLANE 1     /* :1375-1379 */
#define ME_ICP_GENERALIZED    2     /* :1380-1385, the value every shipped config uses */

typedef struct {             /* == open3d RegistrationResult (map_eval.cpp:1367-1394) */
  double  transformation[16];        /* row-major 4x4: registration_result.transformation_ -> trans            */
  double  fitness, inlier_rmse;      /* fitness_ (|corr| / |est|), inlier_rmse_                                  */
  int64_t n_corr;                    /* correspondence_set_.size()                                               */
  int32_t iterations, converged;     /* updates applied; 1 if the relative criteria stopped the loop             */
} me_icp_result;

/* one planning step of the cell lattice (me_plan_lattice) */
typedef struct {
  double  v, h;            /* voxel edge; cell edge h = v / m                                  */
  int32_t m;               /* cells per voxel edge                                             */
  int32_t sparse;          /* 1: occupied 32-cell row segments + hash; 0: dense table          */
  int32_t nvox[3];         /* voxels per axis                                                  */
  int32_t dims[3];         /* cells per axis = nvox * m                                        */
  int64_t ncells;          /* dense table: cells (4 B each); sparse: 0                         */
} me_lattice_plan;

/* ---- lifecycle ---------------------------------------------------------------------------------- */

ME_API int  me_abi_version(void);
ME_API int  me_create(const me_options *opt, me_ctx **out);
ME_API void me_destroy(me_ctx *ctx);
ME_API const char *me_last_error(const me_ctx *ctx);      /* ctx may be NULL: last error of me_create on this thread */
ME_API int  me_set_stream(me_ctx *ctx, void *cuda_stream);
ME_API int  me_set_shard(me_ctx *ctx, int32_t rank, int32_t world);
/* How a context of a multi-GPU job (world > 1) lays the clouds out.  ME_LAYOUT_REPLICATED (default): both whole clouds on
 * every rank, the query ranges of the sweeps sharded.  ME_LAYOUT_SLAB: every rank lays out only the voxel layers it owns
 * along y or z (whichever balances), plus a halo of four lattice cells, of BOTH clouds, and evaluates the points of its
 * layers — the lattice builds and the voxel stage shard with the sweeps.  The clouds themselves stay replicated (caller
 * order, set with me_set_cloud*): searches that leave the halo are finished exactly over the whole cloud.  The partial
 * accumulators are reduced as before (me_accum_block); the voxel stage runs as me_voxel_begin / all-reduce(MAX) of
 * me_voxel_w_table / me_voxel_finish_accum_device.  A scene that cannot be cut (sparse cell table, fewer than 2 x world voxel
 * layers, one rank holding > 75 %) silently stays replicated — me_layout_active tells.  me_eval_awd, the tile sweep and
 * per-point outputs of other ranks' points are not available on an active slab layout. */
#define ME_LAYOUT_REPLICATED 0
#define ME_LAYOUT_SLAB 1
ME_API int  me_set_layout(me_ctx *ctx, int32_t layout);
/* after the lattices were built (any sweep): *layout = the layout in force, *axis = 1 (y) / 2 (z) / 0, n_laid_out[2] = the
 * points of (est, gt) this rank laid out, n_owned[2] = the points it evaluates.  Any pointer may be NULL. */
ME_API int  me_layout_active(me_ctx *ctx, int32_t *layout, int32_t *axis, int64_t n_laid_out[2], int64_t n_owned[2]);
/* The planner of the slab layout as a pure host function (no context, no device): given the number of points in every
 * lattice plane along y and along z (cells_per_voxel planes per voxel layer), cut the voxel layers of one axis into `world`
 * contiguous groups of about equal point count.  *axis = 1 (y) / 2 (z), or 0 when the scene cannot be cut (fewer than
 * 2 x world layers on both axes, or the busiest rank — owned layers + halo_cells planes either side — would lay out more than
 * 75 % of the points); layer_bounds[0..world] = first owned layer of every rank; *busiest_share = that rank's share.  This
 * is the arithmetic every rank runs on the (replicated) cloud's histogram, which is why the ranks agree without talking. */
/* One planning step of the cell lattice as a pure host function (no context, no device): the lattice a cloud of n points
 * with this bounding box gets for a wanted cell edge (<= 0: ~2 points per cell of the box volume), aligned with voxel_size
 * (<= 0: free), fitting `other` (nullable bounding box of the second cloud) under the same spec, within max_grid_cells
 * (<= 0: the automatic budget, 2^28 .. 2^31 growing with the clouds).  Dense if the dense table at that edge fits; sparse
 * when the budget would force cells > 1.25x coarser (allow_sparse).  ME_ERR_RANGE if nothing fits.  The library runs this
 * step 1-3 times per cloud pair, refining the edge on the measured occupancy in between. */
ME_API int  me_plan_lattice(const double bbox_min[3], const double bbox_max[3], int64_t n, const double *other_bbox_min,
                            const double *other_bbox_max, int64_t other_n, double voxel_size, double cell_edge_target,
                            int64_t max_grid_cells, int32_t allow_sparse, me_lattice_plan *out);
ME_API int  me_plan_slab_cut(const uint64_t *plane_counts_y, int32_t n_planes_y, const uint64_t *plane_counts_z,
                             int32_t n_planes_z, int32_t cells_per_voxel, int32_t world, int32_t halo_cells, int32_t *axis,
                             int32_t *layer_bounds, double *busiest_share);
ME_API int  me_synchronize(me_ctx *ctx);

/* replaces: io::ReadPointCloud* results held in map_3d_/gt_3d_ (map_eval.cpp:10-21) — the clouds the path reads.
 * me_set_cloud copies host memory (async on the stream; pinned memory makes it truly async);
 * me_set_cloud_device borrows an fp64 N x 3 device buffer that must outlive the context's use of it. */
ME_API int  me_set_cloud(me_ctx *ctx, int which, const double *xyz_host, int64_t n);
ME_API int  me_set_cloud_device(me_ctx *ctx, int which, const double *xyz_device, int64_t n);
/* replaces: map_3d_->Transform(initial_matrix) (map_eval.cpp:1206); T is a row-major 4x4 */
ME_API int  me_transform(me_ctx *ctx, int which, const double T[16]);
/* replaces: map_3d_ = map_3d_->VoxelDownSample(param_.downsample_size) and the same for gt_3d_ (map_eval.cpp:38-39;
 * open3d::geometry::PointCloud::VoxelDownSample: voxel index floor((p - (min_bound - s/2)) / s), output = mean of the
 * voxel's points accumulated in input order).  The cloud held by the context is replaced by its down-sampled version;
 * *n_out (nullable) receives the new point count.  The output points are bit-identical to the CPU path as a set; they
 * are ordered by increasing voxel index (Open3D: std::unordered_map iteration order, implementation-defined). */
ME_API int  me_voxel_downsample(me_ctx *ctx, int which, double voxel_size, int64_t *n_out);
/* the cloud currently held by the context (after me_transform / me_voxel_downsample), caller order, N x 3 fp64.
 * xyz_host == NULL: size query only (*n).  capacity_points < N is an error. */
ME_API int  me_get_cloud(me_ctx *ctx, int which, double *xyz_host, int64_t capacity_points, int64_t *n);
/* replaces: MapEval::performICPRegistration case 0 (map_eval.cpp:1366-1394): open3d RegistrationICP(est, gt,
 * icp_max_distance, initial_matrix, TransformationEstimationPointToPoint(), ICPConvergenceCriteria{relative_fitness,
 * relative_rmse, max_iteration} = {1e-6, 1e-6, 30}).  On return the estimated cloud held by the context is the original
 * cloud transformed once by the result (map_3d_->Transform(trans), :1392), ready for me_eval_nn with
 * ME_CUTOFF_DIST_LT_R / ME_PAIRING_GEOMETRIC / want_full_cd (calculateMetrics, :1147-1202).  world must be 1. */
ME_API int  me_icp_point_to_point(me_ctx *ctx, double max_correspondence_distance, int32_t max_iteration,
                           double relative_fitness, double relative_rmse, const double T_init[16], me_icp_result *out);
/* replaces: MapEval::performICPRegistration (map_eval.cpp:1366-1394), all three cases; method = param_.evaluation_method_
 * (ME_ICP_*).  ME_ICP_POINT_TO_PLANE = RegistrationICP with TransformationEstimationPointToPlane: the ground-truth cloud
 * needs normals (me_set_normals, e.g. from a PCD that carries them, or me_estimate_normals) — without them the call fails
 * like Open3D does.  ME_ICP_GENERALIZED = RegistrationGeneralizedICP with TransformationEstimationForGeneralizedICP()
 * (epsilon 1e-3): per-point covariances from EstimateNormals(KNN 20) on both clouds, computed inside.  Same post-condition
 * as me_icp_point_to_point.  Normals handed in or estimated before the call do not survive ME_ICP_GENERALIZED. */
ME_API int  me_icp(me_ctx *ctx, int32_t method, double max_correspondence_distance, int32_t max_iteration,
                   double relative_fitness, double relative_rmse, const double T_init[16], me_icp_result *out);
/* normals of a cloud held by the context, caller order, N x 3 fp64 (open3d PointCloud::normals_): upload / estimate as
 * PointCloud::EstimateNormals(KDTreeSearchParamKNN(knn)) does (k nearest neighbours incl. the point, covariance, the
 * eigenvector of the smallest eigenvalue; sign as Open3D's FastEigen3x3 leaves it) / read back.  me_transform rotates them. */
ME_API int  me_set_normals(me_ctx *ctx, int which, const double *normals_host, int64_t n);
ME_API int  me_estimate_normals(me_ctx *ctx, int which, int32_t knn);
ME_API int  me_get_normals(me_ctx *ctx, int which, double *normals_host);
/* builds (or rebuilds) the cell-sorted grid of one cloud; the eval calls build lazily if needed */
ME_API int  me_build_grid(me_ctx *ctx, int which);

/* replaces: MapEval::calculateMetricsWithInitialMatrix (map_eval.cpp:1204-1260), getDiffRegResultWithCorrespondence
 * (:1069-1145), the metric half of calculateMetrics (:1147-1202) and computeChamferDistance (:1398-1431). */
ME_API int  me_eval_nn_accum(me_ctx *ctx, const me_nn_params *p, me_nn_accum *est_to_gt, me_nn_accum *gt_to_est);
ME_API int  me_nn_finalize(const me_nn_params *p, const me_nn_accum *est_to_gt, const me_nn_accum *gt_to_est,
                    int64_t n_est, int64_t n_gt, me_nn_result *out);      /* host only, no device work */
ME_API int  me_eval_nn(me_ctx *ctx, const me_nn_params *p, me_nn_result *out);   /* accum + finalize, world must be 1 */
/* per-query nearest neighbour of the last me_eval_nn*, original point order of the QUERY cloud;
 * which_query = ME_CLOUD_EST -> est->gt sweep.  Only this rank's query range is filled (others -1 / NaN). */
ME_API int  me_get_nn(me_ctx *ctx, int which_query, int32_t *nn_index, double *nn_sqdist);

/* replaces: ComputeMeanMapEntropyUsingNormalTBB / ...UsingNormal (min_neighbors 10, map_eval.cpp:1608-1737, 1538-1606),
 * ComputeMeanMapEntropy (min_neighbors 5, :1438-1535), ComputeEntropy (:1433-1436) and the min/max side effect of
 * ColorPointCloudByMME (:697-701).  entropies_host (nullable) receives N fp64, zeros where invalid. */
ME_API int  me_eval_mme_accum(me_ctx *ctx, int which, double radius, int32_t min_neighbors, me_mme_accum *out);
ME_API int  me_mme_finalize(const me_mme_accum *acc, int64_t n_total, me_mme_result *out);   /* host only */
ME_API int  me_eval_mme(me_ctx *ctx, int which, double radius, int32_t min_neighbors, me_mme_result *out,
                 double *entropies_host);
ME_API int  me_get_entropies(me_ctx *ctx, int which, double *entropies_host);

/* Device-resident accumulators, for multi-GPU passes without host round trips: me_accum_reset clears the context's
 * accumulator block; the *_device calls run the same sweeps as me_eval_nn_accum / me_eval_mme_accum but leave the partial
 * accumulators in that block (fp64; counts ride as fp64, exact below 2^53); me_accum_block hands out its DEVICE address:
 * the first *n_sum values are SUM-reducible, the following *n_max values MAX-reducible (entropy maximum and negated
 * minimum per cloud) — the caller all-reduces them in place (ncclAllReduce / torch.distributed on the context's stream);
 * me_accum_fetch copies the block back ONCE and fills the accumulator structs (any of them may be NULL), ready for
 * me_nn_finalize / me_mme_finalize.  Replaces the reductions of map_eval.cpp:1411,1420,1704-1708 across GPUs. */
ME_API int  me_accum_reset(me_ctx *ctx);
ME_API int  me_eval_nn_accum_device(me_ctx *ctx, const me_nn_params *p);
ME_API int  me_eval_mme_accum_device(me_ctx *ctx, int which, double radius, int32_t min_neighbors);
ME_API int  me_accum_block(me_ctx *ctx, double **device_block, int32_t *n_sum, int32_t *n_max);
ME_API int  me_accum_fetch(me_ctx *ctx, me_nn_accum *est_to_gt, me_nn_accum *gt_to_est, me_mme_accum *mme_est,
                           me_mme_accum *mme_gt);
/* The voxel stage (calculateVMD, map_eval.cpp:240-390) in two halves, for either layout.  me_voxel_begin: lattices,
 * per-voxel Gaussians, pairing and Wasserstein distances of the voxels this rank owns (all of them on a replicated layout).
 * me_voxel_w_table: DEVICE address and length of the W table over the estimated cloud's voxels (-1 = no pair) — on an
 * active slab layout the caller MAX-all-reduces it in place, so that every rank sees the W of its neighbours' voxels (SCS
 * reads 11^3 voxels around each pair, map_eval.cpp:353).  me_voxel_finish_accum_device: the SCS sweep over this rank's
 * pairs; leaves the stage's eight SUM-reducible values in the accumulator block.  me_accum_fetch_awd: after the block's
 * all-reduce, AWD / SCS / the voxel counts (map_eval.cpp:324-325,387) from the block. */
ME_API int  me_voxel_begin(me_ctx *ctx, double voxel_size, int32_t min_points);
ME_API int  me_voxel_w_table(me_ctx *ctx, double **device_w, int64_t *n);
ME_API int  me_voxel_finish_accum_device(me_ctx *ctx, int32_t scs_radius);
ME_API int  me_accum_fetch_awd(me_ctx *ctx, me_awd_result *out);

/* replaces: MapEval::calculateVMD (map_eval.cpp:240-390) with VoxelCalculator::buildVoxelMap / computeVoxelEntropy /
 * updateVoxelMap / computeWassersteinDistanceGaussian / getNeighborIndices (voxel_calculator.cpp:7-56,97-172,241-245).
 * rows27 (nullable): library-allocated n_rows x 27 table = the columns of voxel_errors.txt (map_eval.cpp:292-302),
 * release with me_free.  Replicated layout only: every rank computes the whole voxel stage (me_voxel_* shards it). */
ME_API int  me_eval_awd(me_ctx *ctx, double voxel_size, int32_t min_points, int32_t scs_radius, me_awd_result *out,
                 int64_t *n_rows, double **rows27);
/* The voxel-pair half of calculateVMD on GIVEN voxel Gaussians: rows27 = n_rows x 27 in the column layout of
 * voxel_errors.txt (map_eval.cpp:292-302: vmin[3] vmax[3] mu_est[3] W n_gt n_est sigma_est[6] mu_gt[3] sigma_gt[6]).
 * Recomputes W per row (w_out, nullable) with the device implementation of computeWassersteinDistanceGaussian
 * (voxel_calculator.cpp:115-140) and AWD / SCS over those voxels (map_eval.cpp:324-325, 351-387).  This is how the
 * tests pin the kernels to the sample output the reference ships (map_eval/scripts/voxel_errors.txt). */
ME_API int  me_awd_from_rows(me_ctx *ctx, const double *rows27, int64_t n_rows, double voxel_size, int32_t scs_radius,
                      double *w_out, me_awd_result *out);
ME_API void me_free(void *p);

/* timing of the last call of each stage in milliseconds (CUDA events on the context's stream):
 * [0] grid est [1] grid gt [2] nn est->gt [3] nn gt->est [4] mme est [5] mme gt [6] voxel moments [7] awd [8] scs */
#define ME_N_STAGE_TIMES 9
ME_API int  me_get_stage_times(me_ctx *ctx, double ms[ME_N_STAGE_TIMES]);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
ME_API int64_t me_launch_count(const me_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MAPEVAL_B200_H_ */
