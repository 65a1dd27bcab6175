// nn.cu — exact 1-NN sweeps with fused inlier accumulators.
//
// Replaces (reference, map_eval/src/map_eval.cpp):
//   :1213-1236  two serial KDTreeFlann::SearchKNN(p, 1) loops of calculateMetricsWithInitialMatrix
//   :1069-1145  getDiffRegResultWithCorrespondence (five-threshold accumulators; == :990-1067, :828-897)
//   :1398-1431  computeChamferDistance (sum of sqrt(d2) over ALL points, unbounded NN)
//
// One thread per query, queries walked in their own cell-sorted order so that a warp's 32 queries sit in the same
// few lattice rows and its candidate loads hit the same L1 lines.  Per query the 3x3x3 cell block of the reference
// lattice is read as 9 contiguous x-runs (3 x-adjacent cells are adjacent in the CSR layout).  All candidate
// distances are evaluated in fp64 with exactly the reference's operation order (no FMA contraction), so the
// arg-min, the cut-off test and every inlier comparison are bit-identical to the CPU path.  A query whose best
// distance does not beat the distance to the faces of its searched block is finished by a warp-per-query
// ring-expansion kernel (rare: ~0.1 % of queries on volume-filling clouds).
#include "common.cuh"
#include <algorithm>
#include <cstring>

namespace me {

static constexpr int kThreads = 256;

struct NNConst {
  double tau[5];
  double cutoff;        // R (mode 0) or R*R (mode 1)
  int cutoff_mode;
  int accumulate;       // accumulate pair statistics in the sweep (0 for ME_PAIRING_AS_WRITTEN gt->est)
  int want_full_cd;
  double max_d2;        // nothing beyond this squared distance matters (inf when full CD is wanted)
  double ref_maxabs;    // max |coordinate| of the reference cloud (slack of the face test)
};

// device accumulator block: 8 x int64 then 13 x fp64 (see me_nn_accum)
struct AccBlock {
  unsigned long long n_corr, n_inl[5], n_ub, n_far;
  double sum_d[5], sum_d2[5], sum_d_all, sum_d2_all, sum_nn;
};

struct LocalAcc {
  unsigned int n_corr, n_inl[5], n_ub;
  double sum_d[5], sum_d2[5], sum_d_all, sum_d2_all, sum_nn;
  __device__ void clear() {
    n_corr = 0; n_ub = 0; sum_d_all = 0; sum_d2_all = 0; sum_nn = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) { n_inl[k] = 0; sum_d[k] = 0; sum_d2[k] = 0; }
  }
};

__device__ __forceinline__ bool keep_pair(double d2, const NNConst &c) {
  return c.cutoff_mode == ME_CUTOFF_SQDIST_LE_R ? (d2 <= c.cutoff) : (d2 < c.cutoff);
}

// map_eval.cpp:1095-1123 for one kept pair (source - target)
__device__ __forceinline__ void accum_pair(double dx, double dy, double dz, const NNConst &c, LocalAcc &a) {
  double sq = sqnorm_eigen(dx, dy, dz);
  double nd = __dsqrt_rn(sq);
  a.n_corr++;
  a.sum_d_all += nd;
  a.sum_d2_all += sq;
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (nd <= c.tau[k]) { a.sum_d[k] += nd; a.sum_d2[k] += sq; a.n_inl[k]++; }
}

__device__ void flush_acc(LocalAcc &a, AccBlock *g) {
  __shared__ double sh_d[kThreads / 32][13];
  __shared__ unsigned long long sh_i[kThreads / 32][7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double dv[13];
  long long iv[7];
#pragma unroll
  for (int k = 0; k < 5; ++k) { dv[k] = a.sum_d[k]; dv[5 + k] = a.sum_d2[k]; iv[1 + k] = a.n_inl[k]; }
  dv[10] = a.sum_d_all; dv[11] = a.sum_d2_all; dv[12] = a.sum_nn;
  iv[0] = a.n_corr; iv[6] = a.n_ub;
#pragma unroll
  for (int k = 0; k < 13; ++k) dv[k] = warp_sum(dv[k]);
#pragma unroll
  for (int k = 0; k < 7; ++k) iv[k] = warp_sum_ll(iv[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 13; ++k) sh_d[warp][k] = dv[k];
#pragma unroll
    for (int k = 0; k < 7; ++k) sh_i[warp][k] = (unsigned long long)iv[k];
  }
  __syncthreads();
  if (threadIdx.x < 13) {
    double s = 0;
    for (int w = 0; w < kThreads / 32; ++w) s += sh_d[w][threadIdx.x];
    double *dst = &g->sum_d[0];
    if (s != 0.0) atomicAdd(dst + threadIdx.x, s);
  } else if (threadIdx.x >= 32 && threadIdx.x < 39) {
    int k = threadIdx.x - 32;
    unsigned long long s = 0;
    for (int w = 0; w < kThreads / 32; ++w) s += sh_i[w][k];
    unsigned long long *dst = &g->n_corr;   // n_corr, n_inl[5], n_ub are contiguous
    if (s) atomicAdd(dst + k, s);
  }
}

// face distance of the searched block [ic - r, ic + r] (cells) around continuous coordinate u, in cells;
// faces beyond the lattice are at infinity (nothing lives there)
__device__ __forceinline__ double face_dist_cells(double u, long long ic, int r, int dim) {
  double lo = (ic - r <= 0) ? INFINITY : u - (double)(ic - r);
  double hi = (ic + r + 1 >= dim) ? INFINITY : (double)(ic + r + 1) - u;
  return fmin(lo, hi);
}

// ---------------------------------------------------------------------------------------------------------------
// main sweep
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
nn_sweep_kernel(const P4 *__restrict__ Q, long long q_begin, long long q_end, const P4 *__restrict__ R,
                const uint32_t *__restrict__ cell_off, Lattice L, NNConst C, int32_t *__restrict__ nn_idx,
                double *__restrict__ nn_d2, uint32_t *__restrict__ far_list, unsigned int *__restrict__ far_count,
                AccBlock *__restrict__ acc) {
  LocalAcc a;
  a.clear();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = q_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < q_end; i += stride) {
    const P4 q = load_p4(Q + i);
    long long ix = cell_coord(q.x, L, 0), iy = cell_coord(q.y, L, 1), iz = cell_coord(q.z, L, 2);
    ix = ix < -1 ? -1 : (ix > L.dims[0] ? L.dims[0] : ix);
    iy = iy < -1 ? -1 : (iy > L.dims[1] ? L.dims[1] : iy);
    iz = iz < -1 ? -1 : (iz > L.dims[2] ? L.dims[2] : iz);

    double best = INFINITY;
    long long bidx = 0x7fffffffffffffffll;
    double bdx = 0, bdy = 0, bdz = 0;
    const int x0 = (int)max(ix - 1, 0ll), x1 = (int)min(ix + 1, (long long)L.dims[0] - 1);
    if (x0 <= x1) {
      for (int dz = -1; dz <= 1; ++dz) {
        long long z = iz + dz;
        if (z < 0 || z >= L.dims[2]) continue;
        for (int dy = -1; dy <= 1; ++dy) {
          long long y = iy + dy;
          if (y < 0 || y >= L.dims[1]) continue;
          const long long row = (z * L.dims[1] + y) * (long long)L.dims[0];
          uint32_t s = __ldg(cell_off + row + x0), e = __ldg(cell_off + row + x1 + 1);
          for (uint32_t j = s; j < e; ++j) {
            const P4 p = load_p4(R + j);
            double ddx = __dsub_rn(q.x, p.x), ddy = __dsub_rn(q.y, p.y), ddz = __dsub_rn(q.z, p.z);
            double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz));
            if (d2 < best || (d2 == best && p.idx < bidx)) { best = d2; bidx = p.idx; bdx = ddx; bdy = ddy; bdz = ddz; }
          }
        }
      }
    }
    // is the best provably the global nearest neighbour?
    double ux = cell_coord_cont(q.x, L, 0), uy = cell_coord_cont(q.y, L, 1), uz = cell_coord_cont(q.z, L, 2);
    double g = fmin(fmin(face_dist_cells(ux, ix, 1, L.dims[0]), face_dist_cells(uy, iy, 1, L.dims[1])),
                    face_dist_cells(uz, iz, 1, L.dims[2])) * L.h;
    double slack = 1e-9 * L.h + 1e-14 * (fabs(q.x) + fabs(q.y) + fabs(q.z) + C.ref_maxabs);
    double ge = g - slack;
    double ge2 = ge > 0 ? ge * ge : 0.0;
    bool resolved = best < ge2;
    bool beyond = false;
    if (!resolved && ge2 > C.max_d2) { resolved = true; beyond = best > C.max_d2; }   // nothing farther matters
    if (!resolved) {
      nn_idx[i] = best < INFINITY ? (int32_t)bidx : -1;
      nn_d2[i] = best;
      unsigned int slot = atomicAdd(far_count, 1u);
      far_list[slot] = (uint32_t)(i - q_begin);
      continue;
    }
    if (beyond || !(best < INFINITY)) { nn_idx[i] = -1; nn_d2[i] = INFINITY; continue; }
    nn_idx[i] = (int32_t)bidx;
    nn_d2[i] = best;
    if (C.want_full_cd) a.sum_nn += __dsqrt_rn(best);        // map_eval.cpp:1416
    if (C.accumulate && keep_pair(best, C)) accum_pair(bdx, bdy, bdz, C, a);
  }
  flush_acc(a, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// far queries: one warp per query, Chebyshev rings r = 2, 3, ... until the best beats the block faces
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
nn_far_kernel(const P4 *__restrict__ Q, long long q_begin, const P4 *__restrict__ R,
              const uint32_t *__restrict__ cell_off, Lattice L, NNConst C, int32_t *__restrict__ nn_idx,
              double *__restrict__ nn_d2, const uint32_t *__restrict__ far_list,
              const unsigned int *__restrict__ far_count, AccBlock *__restrict__ acc) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const unsigned int nfar = *far_count;
  for (long long w = warp; w < nfar; w += nwarps) {
    const long long i = q_begin + far_list[w];
    const P4 q = load_p4(Q + i);
    long long ix = cell_coord(q.x, L, 0), iy = cell_coord(q.y, L, 1), iz = cell_coord(q.z, L, 2);
    ix = ix < -1 ? -1 : (ix > L.dims[0] ? L.dims[0] : ix);
    iy = iy < -1 ? -1 : (iy > L.dims[1] ? L.dims[1] : iy);
    iz = iz < -1 ? -1 : (iz > L.dims[2] ? L.dims[2] : iz);
    const double ux = cell_coord_cont(q.x, L, 0), uy = cell_coord_cont(q.y, L, 1), uz = cell_coord_cont(q.z, L, 2);
    const double slack = 1e-9 * L.h + 1e-14 * (fabs(q.x) + fabs(q.y) + fabs(q.z) + C.ref_maxabs);
    double best = nn_d2[i];
    long long bidx = nn_idx[i] >= 0 ? (long long)nn_idx[i] : 0x7fffffffffffffffll;
    bool beyond = false;
    for (int r = 2;; ++r) {
      const int side = 2 * r + 1;
      // rows of the shell: every (dy,dz) in [-r,r]^2; border rows scan the full x range, inner rows two end cells
      for (int t = lane; t < side * side; t += 32) {
        const int dz = t / side - r, dy = t % side - r;
        const long long z = iz + dz, y = iy + dy;
        if (z < 0 || z >= L.dims[2] || y < 0 || y >= L.dims[1]) continue;
        const long long row = (z * L.dims[1] + y) * (long long)L.dims[0];
        const bool border = (dz == -r || dz == r || dy == -r || dy == r);
        for (int part = 0; part < 2; ++part) {
          long long xa, xb;
          if (border) { if (part) break; xa = ix - r; xb = ix + r; }
          else { xa = xb = part ? ix + r : ix - r; }
          xa = max(xa, 0ll); xb = min(xb, (long long)L.dims[0] - 1);
          if (xa > xb) continue;
          uint32_t s = __ldg(cell_off + row + xa), e = __ldg(cell_off + row + xb + 1);
          for (uint32_t j = s; j < e; ++j) {
            const P4 p = load_p4(R + j);
            double d2 = d2_kd(q.x, q.y, q.z, p.x, p.y, p.z);
            if (d2 < best || (d2 == best && p.idx < bidx)) { best = d2; bidx = p.idx; }
          }
        }
      }
      // warp arg-min (distance, then smaller index)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        double ob = __shfl_xor_sync(0xffffffffu, best, o);
        long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob < best || (ob == best && oi < bidx)) { best = ob; bidx = oi; }
      }
      double g = fmin(fmin(face_dist_cells(ux, ix, r, L.dims[0]), face_dist_cells(uy, iy, r, L.dims[1])),
                      face_dist_cells(uz, iz, r, L.dims[2])) * L.h;
      double ge = g - slack;
      double ge2 = ge > 0 ? ge * ge : 0.0;
      if (best < ge2) break;
      if (ge2 > C.max_d2) { beyond = best > C.max_d2; break; }
      if (g == INFINITY) break;   // the block covers the whole lattice
    }
    if (lane == 0) {
      atomicAdd(&acc->n_far, 1ull);
      if (beyond || !(best < INFINITY)) { nn_idx[i] = -1; nn_d2[i] = INFINITY; }
      else {
        nn_idx[i] = (int32_t)bidx;
        nn_d2[i] = best;
        if (C.want_full_cd) atomicAdd(&acc->sum_nn, __dsqrt_rn(best));
        // pair statistics of far queries: nn_far_accum_kernel (needs the winner's coordinates)
      }
    }
  }
}

// far queries' pair statistics (the winner's coordinates come from the caller-order reference array)
__global__ void __launch_bounds__(kThreads)
nn_far_accum_kernel(const P4 *__restrict__ Q, long long q_begin, const double *__restrict__ ref_xyz, NNConst C,
                    const int32_t *__restrict__ nn_idx, const double *__restrict__ nn_d2,
                    const uint32_t *__restrict__ far_list, const unsigned int *__restrict__ far_count,
                    AccBlock *__restrict__ acc) {
  LocalAcc a;
  a.clear();
  const unsigned int nfar = *far_count;
  for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < nfar; w += (long long)gridDim.x * blockDim.x) {
    const long long i = q_begin + far_list[w];
    const int32_t j = nn_idx[i];
    if (j < 0) continue;
    const double d2 = nn_d2[i];
    if (!keep_pair(d2, C)) continue;
    const P4 q = load_p4(Q + i);
    double px = __ldg(ref_xyz + 3ll * j), py = __ldg(ref_xyz + 3ll * j + 1), pz = __ldg(ref_xyz + 3ll * j + 2);
    accum_pair(__dsub_rn(q.x, px), __dsub_rn(q.y, py), __dsub_rn(q.z, pz), C, a);
  }
  flush_acc(a, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// ME_PAIRING_AS_WRITTEN: map_eval.cpp:1233 stores (nn_est, i_gt); :1241 passes (source = gt, target = est), so
// :1093-1094 reads gt[nn_est] and est[i_gt].  Reproduced verbatim; out-of-range (UB in the reference) is counted.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
pair_as_written_kernel(const P4 *__restrict__ Qgt, long long q_begin, long long q_end,
                       const int32_t *__restrict__ nn_idx, const double *__restrict__ nn_d2,
                       const double *__restrict__ gt_xyz, long long n_gt, const double *__restrict__ est_xyz,
                       long long n_est, NNConst C, AccBlock *__restrict__ acc) {
  LocalAcc a;
  a.clear();
  for (long long i = q_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < q_end;
       i += (long long)gridDim.x * blockDim.x) {
    const int32_t nn_est = nn_idx[i];
    if (nn_est < 0) continue;
    if (!keep_pair(nn_d2[i], C)) continue;
    const long long i_gt = __double_as_longlong(__ldg(reinterpret_cast<const double *>(Qgt + i) + 3));
    const long long s = nn_est, t = i_gt;          // source index into gt, target index into est
    if (s >= n_gt || t >= n_est) { a.n_ub++; continue; }
    double dx = __dsub_rn(__ldg(gt_xyz + 3 * s), __ldg(est_xyz + 3 * t));
    double dy = __dsub_rn(__ldg(gt_xyz + 3 * s + 1), __ldg(est_xyz + 3 * t + 1));
    double dz = __dsub_rn(__ldg(gt_xyz + 3 * s + 2), __ldg(est_xyz + 3 * t + 2));
    accum_pair(dx, dy, dz, C, a);
  }
  flush_acc(a, acc);
}

__global__ void fill_nn_kernel(int32_t *idx, double *d2, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    idx[i] = -1; d2[i] = NAN;
  }
}
__global__ void unsort_nn_kernel(const P4 *__restrict__ Q, long long b, long long e, const int32_t *__restrict__ sidx,
                                 const double *__restrict__ sd2, int32_t *__restrict__ oidx, double *__restrict__ od2) {
  for (long long i = b + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e; i += (long long)gridDim.x * blockDim.x) {
    long long o = __double_as_longlong(__ldg(reinterpret_cast<const double *>(Q + i) + 3));
    oidx[o] = sidx[i]; od2[o] = sd2[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------------------------
static int run_direction(me_ctx *ctx, int qwhich, const me_nn_params *p, me_nn_accum *out, int stage) {
  Cloud &Qc = ctx->cloud[qwhich];
  Cloud &Rc = ctx->cloud[1 - qwhich];
  StageTimer timer(ctx, stage);
  long long qb, qe;
  shard_range(ctx, Qc.n, &qb, &qe);
  const long long nq = qe - qb;

  NNConst C;
  for (int k = 0; k < 5; ++k) C.tau[k] = p->tau[k];
  C.cutoff_mode = p->cutoff_mode;
  C.cutoff = p->cutoff_mode == ME_CUTOFF_SQDIST_LE_R ? p->icp_max_distance : p->icp_max_distance * p->icp_max_distance;
  const bool as_written = (qwhich == ME_CLOUD_GT && p->pairing == ME_PAIRING_AS_WRITTEN);
  C.accumulate = as_written ? 0 : 1;
  C.want_full_cd = p->want_full_cd ? 1 : 0;
  C.max_d2 = p->want_full_cd ? INFINITY : C.cutoff;
  C.ref_maxabs = 0;
  for (int a = 0; a < 3; ++a) C.ref_maxabs = std::max(C.ref_maxabs, std::max(std::fabs(Rc.bbox_min[a]), std::fabs(Rc.bbox_max[a])));

  ME_TRY(ensure(ctx, (void **)&Qc.d_nn_idx, &Qc.cap_nn, Qc.n, sizeof(int32_t)));
  ME_TRY(ensure(ctx, (void **)&Qc.d_nn_d2, &Qc.cap_nn_d2, Qc.n, sizeof(double)));
  // work buffer: far list (nq uint32) after a 256-byte header holding the far counter
  ME_TRY(ensure_work(ctx, 256 + (size_t)std::max<long long>(nq, 1) * sizeof(uint32_t)));
  unsigned int *far_count = (unsigned int *)ctx->d_work;
  uint32_t *far_list = (uint32_t *)((char *)ctx->d_work + 256);
  AccBlock *acc = (AccBlock *)ctx->d_scratch;
  ME_CUDA(ctx, cudaMemsetAsync(acc, 0, sizeof(AccBlock), ctx->stream));
  ME_CUDA(ctx, cudaMemsetAsync(far_count, 0, sizeof(unsigned int), ctx->stream));

  if (nq > 0) {
    int blocks = (int)std::min<long long>((nq + kThreads - 1) / kThreads, (long long)ctx->sm_count * 32);
    nn_sweep_kernel<<<blocks, kThreads, 0, ctx->stream>>>(Qc.d_sorted, qb, qe, Rc.d_sorted, Rc.d_cell_off, Rc.lat, C,
                                                         Qc.d_nn_idx, Qc.d_nn_d2, far_list, far_count, acc);
    ME_LAUNCH_CHECK(ctx);
    int fblocks = ctx->sm_count * 4;
    nn_far_kernel<<<fblocks, kThreads, 0, ctx->stream>>>(Qc.d_sorted, qb, Rc.d_sorted, Rc.d_cell_off, Rc.lat, C,
                                                        Qc.d_nn_idx, Qc.d_nn_d2, far_list, far_count, acc);
    ME_LAUNCH_CHECK(ctx);
    if (C.accumulate) {
      nn_far_accum_kernel<<<ctx->sm_count, kThreads, 0, ctx->stream>>>(Qc.d_sorted, qb, Rc.d_xyz, C, Qc.d_nn_idx,
                                                                       Qc.d_nn_d2, far_list, far_count, acc);
      ME_LAUNCH_CHECK(ctx);
    }
    if (as_written) {
      pair_as_written_kernel<<<blocks, kThreads, 0, ctx->stream>>>(Qc.d_sorted, qb, qe, Qc.d_nn_idx, Qc.d_nn_d2,
                                                                  Qc.d_xyz, Qc.n, Rc.d_xyz, Rc.n, C, acc);
      ME_LAUNCH_CHECK(ctx);
    }
  }
  AccBlock *h = (AccBlock *)ctx->h_pinned;
  ME_CUDA(ctx, cudaMemcpyAsync(h, acc, sizeof(AccBlock), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::memset(out, 0, sizeof(*out));
  out->n_query = nq;
  out->n_corr = (int64_t)h->n_corr;
  for (int k = 0; k < 5; ++k) { out->n_inlier[k] = (int64_t)h->n_inl[k]; out->sum_d[k] = h->sum_d[k]; out->sum_d2[k] = h->sum_d2[k]; }
  out->n_ub = (int64_t)h->n_ub;
  out->n_far = (int64_t)h->n_far;
  out->sum_d_all = h->sum_d_all; out->sum_d2_all = h->sum_d2_all; out->sum_nn_dist = h->sum_nn;
  Qc.nn_valid = true;
  return ME_OK;
}

int run_nn(me_ctx *ctx, const me_nn_params *p, me_nn_accum *e2g, me_nn_accum *g2e) {
  if (ctx->cloud[0].n <= 0 || ctx->cloud[1].n <= 0)
    return fail(ctx, ME_ERR_EMPTY, "both clouds must be set (map_eval.cpp:32-35)");
  ME_TRY(build_grid(ctx, ME_CLOUD_EST));
  ME_TRY(build_grid(ctx, ME_CLOUD_GT));
  const int dirs = p->directions ? p->directions : 3;
  if (e2g) std::memset(e2g, 0, sizeof(*e2g));
  if (g2e) std::memset(g2e, 0, sizeof(*g2e));
  if ((dirs & 1) && e2g) ME_TRY(run_direction(ctx, ME_CLOUD_EST, p, e2g, 2));
  if ((dirs & 2) && g2e) ME_TRY(run_direction(ctx, ME_CLOUD_GT, p, g2e, 3));
  return ME_OK;
}

int unsort_nn(me_ctx *ctx, int which_query, int32_t *h_idx, double *h_d2) {
  Cloud &Qc = ctx->cloud[which_query];
  if (!Qc.nn_valid) return fail(ctx, ME_ERR_INVALID, "me_get_nn before me_eval_nn");
  long long qb, qe;
  shard_range(ctx, Qc.n, &qb, &qe);
  size_t bytes = (size_t)Qc.n * (sizeof(int32_t) + sizeof(double)) + 256;
  ME_TRY(ensure_work(ctx, bytes));
  double *od2 = (double *)ctx->d_work;
  int32_t *oidx = (int32_t *)((char *)ctx->d_work + (size_t)Qc.n * sizeof(double));
  int blocks = (int)std::min<long long>((Qc.n + kThreads - 1) / kThreads, (long long)ctx->sm_count * 16);
  fill_nn_kernel<<<blocks, kThreads, 0, ctx->stream>>>(oidx, od2, Qc.n);
  ME_LAUNCH_CHECK(ctx);
  if (qe > qb) {
    unsort_nn_kernel<<<blocks, kThreads, 0, ctx->stream>>>(Qc.d_sorted, qb, qe, Qc.d_nn_idx, Qc.d_nn_d2, oidx, od2);
    ME_LAUNCH_CHECK(ctx);
  }
  if (h_idx) ME_CUDA(ctx, cudaMemcpyAsync(h_idx, oidx, (size_t)Qc.n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  if (h_d2) ME_CUDA(ctx, cudaMemcpyAsync(h_d2, od2, (size_t)Qc.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ME_OK;
}

}  // namespace me
