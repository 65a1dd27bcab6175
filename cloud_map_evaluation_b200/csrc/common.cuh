// common.cuh — context, lattice and device helpers shared by the kernels of libmapeval_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <string>
#include "../../include/mapeval_b200.h"

namespace me {

// One point of a cell-sorted cloud: absolute fp64 coordinates (the reference's storage type,
// std::vector<Eigen::Vector3d>) + a tag: low 32 bits = index the point had in the caller's array, high 32 bits =
// lattice cell of the point.  32 B, so a point is two aligned 16-byte loads (and a valid TMA bulk-copy unit).
struct __align__(32) P4 {
  double x, y, z;
  long long idx;
};
__host__ __device__ __forceinline__ int orig_of(long long tag) { return (int)(tag & 0xffffffffll); }
__host__ __device__ __forceinline__ unsigned int cell_of(long long tag) { return (unsigned int)((unsigned long long)tag >> 32); }

static constexpr int kTileEdge = 4;   // query tiles are kTileEdge^3 cells

// Dense lattice of a cloud.  Cells are cubes of edge h = v / m whose boundaries coincide with the voxel
// boundaries floor(x / v) of the reference (voxel_calculator.cpp:241-245): every cell belongs to exactly one
// voxel, so the voxel stage needs no hashing and no atomics.  Without a voxel size v := h, m := 1.
struct Lattice {
  double v;          // voxel edge
  double h;          // cell edge = v / m
  double m_over_v;   // m / v
  int m;             // cells per voxel edge
  int k_lo[3];       // lowest voxel index per axis (floor(min / v))
  int nvox[3];       // voxels per axis
  int dims[3];       // cells per axis = nvox * m
  int nb[3];         // tiles (kTileEdge^3 cells) per axis
  long long ncells;  // dense table only
  long long nvoxels;
  int sparse;        // 1: the cell table holds occupied row segments only (CellIndex below)
  int nsegx;         // sparse: segments of kSegCells cells per lattice row = ceil(dims[0] / kSegCells)
};

// Cell table of a lattice = CSR offsets into the cell-sorted cloud, cells in (z, y, x) order, so that x-adjacent cells are
// adjacent in the cloud and the run of cells [xa, xb] of one lattice row is one contiguous range of points.
//   dense : off[c] for every cell c (ncells + 1 entries) — scenes whose bounding box fits the cell budget
//   sparse: every lattice row is cut into segments of 32 cells along x; only OCCUPIED segments are stored, still in
//           (z, y, x) order — off[rank * 32 + (x & 31)] — and an open-addressing hash table maps the segment key
//           (row * nsegx + x / 32) to its rank.  Site-scale surface scans keep ~2 points per occupied cell whatever
//           the extent; the point order, and with it every run-walking kernel, is the same as in the dense case.
static constexpr int kSegCells = 32;
struct CellIndex {
  const uint32_t *off;
  const unsigned long long *hkey;   // sparse: hash slots, ~0 = empty (linear probing)
  const uint32_t *hval;             // rank of the segment
  uint32_t hmask;                   // slots - 1
  int nsegx;
  int sparse;
  int dimx, dimy, dimz;
};

struct Cloud {
  long long n = 0;
  long long ns = 0;             // points in the sorted arrays: n, or the points of this rank's slab (+ halo) in slab mode
  // slab layout: only the lattice planes [sc_lo, sc_hi) along lattice axis sl_axis (1: y, 2: z) are laid out here, and this
  // rank owns the queries of the planes [so_lo, so_hi) — n_owned points; the sweeps skip the halo points
  bool slab = false;
  int sl_axis = 0, so_lo = 0, so_hi = 0, sc_lo = 0, sc_hi = 0;
  long long n_owned = 0;
  double *d_xyz = nullptr;      // caller order, fp64 AoS
  bool owned = false;
  long long cap_xyz = 0;
  bool grid_valid = false;
  bool grid_solo = false;       // the lattice is this cloud's own (MME with a large radius), not the shared voxel-aligned one
  double solo_h = 0.0;
  double bbox_min[3], bbox_max[3];
  bool bbox_valid = false;
  Lattice lat;
  P4 *d_sorted = nullptr;       // cell-sorted points
  long long cap_sorted = 0;
  float4 *d_rel = nullptr;      // cell-sorted fp32 screening copy: xyz relative to the point's own cell origin, w = (float)ix
  long long cap_rel = 0;
  uint32_t *d_cell_off = nullptr;   // dense: ncells + 1 CSR offsets into d_sorted: cell c = [off[c], off[c+1]); sparse: nseg * 32 + 1
  long long cap_cells = 0;
  uint32_t *d_coarse = nullptr;     // point counts of coarse cells (coarse_f^3 lattice cells each): lets far queries skip empty space
  long long cap_coarse = 0;
  int coarse_f = 0;                 // 0: no coarse grid (small lattices)
  int coarse_dims[3] = {0, 0, 0};
  unsigned long long *d_hkey = nullptr;   // sparse lattice: segment hash table (CellIndex)
  uint32_t *d_hval = nullptr;
  long long cap_hash = 0;
  uint32_t hmask = 0;
  long long n_seg = 0;
  uint32_t *d_cell_id = nullptr;    // scratch: cell of each point (caller order)
  long long cap_cell_id = 0;
  uint32_t *d_tiles = nullptr;      // non-empty query tiles (tile id = (bz*nb[1]+by)*nb[0]+bx)
  long long cap_tiles = 0;
  long long n_tiles = 0;
  bool tiles_valid = false;
  uint32_t *d_tile_pos = nullptr;   // tile-list build scratch (flags -> positions)
  long long cap_tile_pos = 0;
  long long shard_b = 0, shard_e = 0;   // this rank's cell-aligned query range (query_shard)
  int shard_rank = -1, shard_world = 0;
  bool shard_valid = false;
  long long max_cell_count = 0;     // points in the fullest cell
  cudaEvent_t upload_done = nullptr;  // recorded on the copy stream after me_set_cloud's H2D
  bool upload_pending = false;
  // per-query results, SORTED order of this cloud (unsorted on demand)
  int32_t *d_nn_idx = nullptr;      // nearest neighbour in the other cloud (caller index there)
  double *d_nn_d2 = nullptr;        // squared distance as nanoflann accumulates it (cut-off tests, full Chamfer)
  double *d_nn_sq = nullptr;        // squared norm as Eigen accumulates it (inlier statistics)
  long long cap_nn = 0, cap_nn_d2 = 0, cap_nn_sq = 0;
  bool nn_valid = false;
  double *d_entropy = nullptr;      // sorted order of the lattice the MME sweep ran on
  long long cap_entropy = 0;
  bool entropy_valid = false;
  double *d_entropy_caller = nullptr;   // caller order, kept when the sweep ran on a solo lattice (which the next build replaces)
  long long cap_entropy_caller = 0;
  bool entropy_caller_valid = false;
  double *d_normal = nullptr;           // caller order, 3 fp64 per point: normals (me_set_normals / me_estimate_normals); during
  long long cap_normal = 0;             // generalized ICP the unit vector that carries the point's covariance (icp.cu)
  bool normal_valid = false;
};

// Lattice planning is data dependent (the cell edge is refined on the measured occupancy; the slabs are cut on the layer
// histogram) and costs extra histogram passes and host round trips.  The outcome for one (cloud, other cloud, settings)
// constellation is remembered: evaluating a cloud pair of the same sizes and bounding boxes again — repeated passes over the
// same maps — re-uses it.  The plan only steers performance; results do not depend on it.
struct PlanCache {
  bool valid = false;
  // key
  long long n = 0, other_n = -1;
  double bmin[3], bmax[3], obmin[3], obmax[3];
  double v_req = 0, nn_cell = 0;
  long long budget = 0;
  bool sp = false, slab_request = false;
  int rank = 0, world = 1;
  // value
  double h_target = 0;
  Lattice lat;
  bool slab_planned = false, slab_on = false;
  int slab_axis = 0;
  long long slab_k0 = 0, slab_k1 = 0;
};

}  // namespace me

struct me_ctx {
  int device = 0;
  int rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  double nn_cell_size = 0.0;
  long long max_grid_cells = 0;     // budget of the dense cell table; 0 = automatic (grid_budget)
  double voxel_hint = 0.0;          // lattice alignment requested by the voxel stage
  // slab layout (me_set_layout, world > 1, dense lattices): every rank lays out only the voxel layers it owns (+ a halo of
  // slab_halo cells) of both clouds instead of the whole clouds; planned from the layer histogram of the cloud that is laid
  // out first after the lattice spec was (re)planned — identical on every rank, the clouds being replicated
  bool slab_request = false, slab_planned = false, slab_on = false;
  int slab_axis = 0;                       // 1: y, 2: z
  long long slab_k0 = 0, slab_k1 = 0;      // owned world voxel layers [k0, k1) along slab_axis (LLONG_MIN/4, LLONG_MAX/4 at the ends)
  int slab_halo = 4;                       // cells; covers the 3 rings of an MME sweep on the shared lattice
  me::PlanCache plan_cache[2];
  // voxel stage split in two (me_voxel_begin / me_voxel_finish_accum_device): state kept between the halves
  bool vox_open = false;
  long long vox_nvox = 0;
  size_t vox_o_w = 0, vox_o_pairs = 0;
  // lattice spec shared by both clouds, so that their cells coincide (same v, m; integer index offsets)
  double spec_v = 0.0;
  int spec_m = 0;
  cudaStream_t copy_stream = nullptr;   // H2D uploads run here and overlap kernels that do not need them yet
  cudaEvent_t compute_mark = nullptr;
  me::Cloud cloud[2];
  // small device scratch for reductions / accumulators
  void *d_scratch = nullptr;
  size_t scratch_bytes = 0;
  void *h_pinned = nullptr;         // pinned host mirror of the scratch
  void *stage = nullptr;            // pinned bounce buffers + streams of the pageable-memory upload path (api.cu)
  double *d_block = nullptr;        // device-resident accumulator block of the *_device calls (ME_BLOCK_* layout, api.cu)
  // large device scratch (scan partials, far list, voxel tables)
  void *d_work = nullptr;
  size_t work_bytes = 0;
  uint32_t *d_scan_tmp = nullptr;   // per-tile partials of exclusive_scan_inplace
  long long cap_scan_tmp = 0;
  uint32_t *d_rs_hist = nullptr;    // radix sort: digit histograms of every tile
  long long cap_rs_hist = 0;
  cudaEvent_t ev[2 * ME_N_STAGE_TIMES];
  bool ev_used[ME_N_STAGE_TIMES];
  int sm_count = 148;
  long long launches = 0;
  std::string err;
};

namespace me {

int fail(me_ctx *ctx, int code, const std::string &msg);
int ensure(me_ctx *ctx, void **ptr, long long *cap, long long need, size_t elem);
int ensure_work(me_ctx *ctx, size_t bytes);

#define ME_CUDA(ctx, call)                                                                         \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return me::fail((ctx), ME_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
  } while (0)

#define ME_TRY(expr)                \
  do {                              \
    int rc__ = (expr);              \
    if (rc__ != ME_OK) return rc__; \
  } while (0)

#define ME_LAUNCH_CHECK(ctx)                                                                       \
  do {                                                                                             \
    (ctx)->launches++;                                                                             \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return me::fail((ctx), ME_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(e__)); \
  } while (0)

// contiguous range r of W over n items (tile lists; the point ranges of the sweeps come from query_shard, which snaps
// them to cell boundaries)
inline void shard_range(const me_ctx *ctx, long long n, long long *b, long long *e) {
  *b = n * ctx->rank / ctx->world;
  *e = n * (ctx->rank + 1) / ctx->world;
}

struct StageTimer {
  me_ctx *ctx; int stage;
  StageTimer(me_ctx *c, int s) : ctx(c), stage(s) { cudaEventRecord(c->ev[2 * s], c->stream); }
  ~StageTimer() { cudaEventRecord(ctx->ev[2 * stage + 1], ctx->stream); ctx->ev_used[stage] = true; }
};

// the queries a sweep evaluates: all of [q_begin, q_end), or in slab layout those whose cell lies in the owned planes
struct Owned {
  int axis;      // 0: everything, 1: y, 2: z
  int lo, hi;
};
__host__ __device__ inline bool owns(const Owned &o, int iy, int iz) {
  if (o.axis == 0) return true;
  const int a = o.axis == 1 ? iy : iz;
  return a >= o.lo && a < o.hi;
}
inline Owned owned_of(const Cloud &c) {
  Owned o;
  o.axis = c.slab ? c.sl_axis : 0; o.lo = c.so_lo; o.hi = c.so_hi;
  return o;
}

// coarse occupancy grid of a lattice (grid.cu build_coarse): cnt[(cz * cd[1] + cy) * cd[0] + cx] points in the block of f^3 cells
struct CoarseGrid {
  const uint32_t *cnt;      // nullptr: none
  int f;
  int cd[3];
};
inline CoarseGrid coarse_of(const Cloud &c) {
  CoarseGrid g;
  g.cnt = c.coarse_f > 0 ? c.d_coarse : nullptr;
  g.f = c.coarse_f;
  for (int a = 0; a < 3; ++a) g.cd[a] = c.coarse_dims[a];
  return g;
}

inline CellIndex index_of(const Cloud &c) {
  CellIndex I;
  I.off = c.d_cell_off; I.hkey = c.d_hkey; I.hval = c.d_hval; I.hmask = c.hmask;
  I.nsegx = c.lat.nsegx; I.sparse = c.lat.sparse;
  I.dimx = c.lat.dims[0]; I.dimy = c.lat.dims[1]; I.dimz = c.lat.dims[2];
  return I;
}

// stage entry points (implemented in grid.cu / nn.cu / mme.cu / voxel.cu)
int wait_upload(me_ctx *ctx, int which);
int compute_bbox(me_ctx *ctx, int which);
int build_grid(me_ctx *ctx, int which, double solo_h = 0.0);
double density_edge(const Cloud &c);
int build_both(me_ctx *ctx);
int build_tiles(me_ctx *ctx, int which);
int query_shard(me_ctx *ctx, int which, long long *b, long long *e);
int exclusive_scan_inplace(me_ctx *ctx, uint32_t *a, long long n);
int voxel_downsample(me_ctx *ctx, int which, double voxel_size, int64_t *n_out);
int radix_sort_pairs(me_ctx *ctx, unsigned long long *keys, uint32_t *vals, unsigned long long *keys_tmp, uint32_t *vals_tmp,
                     long long n, int key_bits, unsigned long long **keys_sorted, uint32_t **vals_sorted);
int run_nn(me_ctx *ctx, const me_nn_params *p, me_nn_accum *e2g, me_nn_accum *g2e, bool to_block = false);
int unsort_nn(me_ctx *ctx, int which_query, int32_t *h_idx, double *h_d2);
int run_mme(me_ctx *ctx, int which, double radius, int min_neighbors, me_mme_accum *out, bool to_block = false);

// layout of me_ctx::d_block (fp64): the SUM-reducible part first, then the MAX-reducible part.  Counts ride as fp64 (exact
// below 2^53, so their sums are exact and order-independent).
static constexpr int kBlkNN = 22;                    // per direction: n_query n_corr n_inlier[5] n_ub n_far | sum_d[5] sum_d2[5] sum_d_all sum_d2_all sum_nn
static constexpr int kBlkMmeSum = 2 * kBlkNN;        // per cloud: n_query n_valid sum_entropy
static constexpr int kBlkAwd = kBlkMmeSum + 6;       // voxel stage: n_pairs n_scs n_voxels_est n_voxels_gt n_active n_new sum_w sum_scs
static constexpr int kBlkSumCount = kBlkAwd + 8;     // = 58 SUM values
static constexpr int kBlkMax = kBlkSumCount;         // per cloud: max_entropy, -min_entropy
static constexpr int kBlkMaxCount = 4;
static constexpr int kBlkTotal = kBlkSumCount + kBlkMaxCount;
int unsort_entropy(me_ctx *ctx, int which, double *h_entropy);
int run_awd(me_ctx *ctx, double voxel_size, int min_points, int scs_radius, me_awd_result *out, int64_t *n_rows,
            double **rows27);
bool plan_lattice_host(const double bmin[3], const double bmax[3], long long n, const double *obmin, const double *obmax,
                       long long other_n, double v_req, double h_target, long long budget, bool allow_sparse, Lattice *out);
static constexpr double kSlabMaxShare = 0.75;      // a slab cut whose busiest rank lays out more than this share is not worth it
void slab_cut(const unsigned long long *planes_y, int n_planes_y, const unsigned long long *planes_z, int n_planes_z, int m,
              int world, int halo, int *axis_out, int *bounds, double *share_out);
int voxel_begin(me_ctx *ctx, double voxel_size, int min_points);
int voxel_w_table(me_ctx *ctx, double **d_w, int64_t *n);
int voxel_finish_block(me_ctx *ctx, int scs_radius);
int transform_cloud(me_ctx *ctx, int which, const double T[16]);
int run_icp(me_ctx *ctx, int method, double max_dist, int max_iter, double rel_fitness, double rel_rmse, const double T_init[16],
            me_icp_result *out);
int estimate_normals(me_ctx *ctx, int which, int knn, int gicp);
void invalidate_cloud(Cloud &c);      // the cloud's coordinates changed: bbox, lattice, NN and entropy results are stale
int run_awd_rows(me_ctx *ctx, const double *rows27, int64_t n_rows, double voxel_size, int scs_radius, double *w_out,
                 me_awd_result *out);

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------

// squared distance exactly as nanoflann's L2_Adaptor accumulates it: ((dx*dx) + dy*dy) + dz*dz, no contraction
__device__ __forceinline__ double d2_kd(double ax, double ay, double az, double bx, double by, double bz) {
  double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
// Eigen fixed-size squaredNorm: x*x + (y*y + z*z)
__device__ __forceinline__ double sqnorm_eigen(double dx, double dy, double dz) {
  return __dadd_rn(__dmul_rn(dx, dx), __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dz, dz)));
}

// lattice coordinate of one axis: voxel index floor(x / v) (IEEE division, as the reference) and the cell inside it
__device__ __forceinline__ long long cell_coord(double x, const Lattice &L, int axis) {
  double kf = floor(__ddiv_rn(x, L.v));
  double fx = __dsub_rn(x, __dmul_rn(kf, L.v));
  int s = (int)floor(fx * L.m_over_v);
  s = s < 0 ? 0 : (s >= L.m ? L.m - 1 : s);
  // clamp far-away queries before the integer conversion can overflow
  double kk = kf - (double)L.k_lo[axis];
  if (kk < -4.0e9) kk = -4.0e9;
  if (kk > 4.0e9) kk = 4.0e9;
  return (long long)kk * L.m + s;
}
// continuous cell coordinate (for face distances)
__device__ __forceinline__ double cell_coord_cont(double x, const Lattice &L, int axis) {
  return (x - (double)L.k_lo[axis] * L.v) / L.h;
}

// offset of coordinate x from the origin of lattice cell `ic` on `axis` (fp64; the caller rounds it to fp32).
// Origin = (k_lo + ic / m) * v + (ic % m) * h, the same products cell_coord() forms, so |offset| < ~h.
__device__ __forceinline__ double cell_rel(double x, uint32_t ic, const Lattice &L, int axis) {
  const uint32_t kv = ic / (uint32_t)L.m, s = ic - kv * (uint32_t)L.m;
  const double fx = __dsub_rn(x, __dmul_rn((double)(L.k_lo[axis] + (int)kv), L.v));
  return fx - (double)s * L.h;
}

__device__ __forceinline__ P4 load_p4(const P4 *p) {
  const double2 *q = reinterpret_cast<const double2 *>(p);
  double2 a = __ldg(q), b = __ldg(q + 1);
  P4 r;
  r.x = a.x; r.y = a.y; r.z = b.x; r.idx = __double_as_longlong(b.y);
  return r;
}

// ---- cell table access -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t seg_hash(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 29;
  return (uint32_t)k;
}
__device__ __forceinline__ bool seg_find(const CellIndex &I, unsigned long long key, uint32_t &rank) {
  uint32_t h = seg_hash(key) & I.hmask;
  for (;;) {
    const unsigned long long k = __ldg(I.hkey + h);
    if (k == key) { rank = __ldg(I.hval + h); return true; }
    if (k == ~0ull) return false;
    h = (h + 1) & I.hmask;
  }
}
// points of the cells [xa, xb] of lattice row (y, z): [s, e) of the sorted cloud; the caller keeps 0 <= xa <= xb < dimx
// and the row inside the lattice
// SP: -1 = decide at run time (I.sparse), 0 / 1 = compile-time dense / sparse (the hot sweeps are instantiated for both, so
// the dense path carries no trace of the hash lookup)
template <int SP = -1>
__device__ __forceinline__ void cell_range(const CellIndex &I, int z, int y, int xa, int xb, uint32_t &s, uint32_t &e) {
  if (SP == 0 || (SP < 0 && !I.sparse)) {
    const uint32_t row = ((uint32_t)z * (uint32_t)I.dimy + (uint32_t)y) * (uint32_t)I.dimx;
    s = __ldg(I.off + row + (uint32_t)xa);
    e = __ldg(I.off + row + (uint32_t)xb + 1);
    return;
  }
  const unsigned long long rowkey = ((unsigned long long)z * (unsigned long long)I.dimy + (unsigned long long)y) * (unsigned long long)I.nsegx;
  const int ga = xa >> 5, gb = xb >> 5;
  s = 0; e = 0;
  bool have = false;
  for (int g = ga; g <= gb; ++g) {
    uint32_t r;
    if (!seg_find(I, rowkey + (unsigned long long)g, r)) continue;
    const uint32_t base = r << 5;
    if (!have) { s = __ldg(I.off + base + (g == ga ? (uint32_t)(xa & 31) : 0u)); have = true; }
    e = __ldg(I.off + base + (g == gb ? (uint32_t)(xb & 31) + 1u : 32u));
  }
}
// lattice cell of a point of the sorted cloud from its tag (dense: the cell id; sparse: the row id z * dimy + y, the x
// index comes from the cell-relative copy or from the coordinate)
__device__ __forceinline__ void cell_from_tag(const CellIndex &I, long long tag, int ix_if_sparse, int &ix, int &iy, int &iz) {
  const uint32_t t = cell_of(tag);
  if (I.sparse) { ix = ix_if_sparse; iy = (int)(t % (uint32_t)I.dimy); iz = (int)(t / (uint32_t)I.dimy); }
  else {
    ix = (int)(t % (uint32_t)I.dimx);
    const uint32_t cyz = t / (uint32_t)I.dimx;
    iy = (int)(cyz % (uint32_t)I.dimy); iz = (int)(cyz / (uint32_t)I.dimy);
  }
}

// order-preserving map double -> uint64 (for atomicMin/atomicMax on doubles)
__device__ __host__ __forceinline__ unsigned long long enc_ordered(double d) {
#ifdef __CUDA_ARCH__
  unsigned long long u = (unsigned long long)__double_as_longlong(d);
#else
  unsigned long long u; memcpy(&u, &d, 8);
#endif
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __host__ __forceinline__ double dec_ordered(unsigned long long u) {
  u = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d; memcpy(&d, &u, 8); return d;
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Eigen 3x3 determinant (cofactor expansion, bruteforce_det3_helper); m row-major
__device__ __forceinline__ double det3(const double *m) {
  double a = m[0] * (m[4] * m[8] - m[5] * m[7]);
  double b = m[1] * (m[3] * m[8] - m[5] * m[6]);
  double c = m[2] * (m[3] * m[7] - m[4] * m[6]);
  return a - b + c;
}

}  // namespace me
