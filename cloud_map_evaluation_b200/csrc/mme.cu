// mme.cu — mean map entropy: per-point radius neighbourhood -> 3x3 covariance -> 0.5 ln(2 pi e det).
//
// Replaces (reference, map_eval/src/map_eval.cpp):
//   :1608-1737  ComputeMeanMapEntropyUsingNormalTBB   (k >= 10, default for the estimated map)
//   :1538-1606  ComputeMeanMapEntropyUsingNormal      (k >= 10, OpenMP variant)
//   :1438-1535  ComputeMeanMapEntropy                 (k >=  5, serial, ground truth)
//   :1433-1436  ComputeEntropy, and the min/max side effect of ColorPointCloudByMME (:697-701)
//
// The reference materialises every neighbour list (KDTreeFlann::SearchRadius -> Eigen::MatrixXd(3,k)); here no list
// exists.  Accepted points are folded into nine fp64 moments taken about the query point itself (sum d, sum d d^T):
// the query is its own nearest neighbour at d = 0 and contributes nothing, which is exactly the reference's
// "erase the first hit" (:1672-1673), and cov = (S2 - S1 S1^T / k) / (k - 1) equals the reference's centred product.
//
// One thread per query, queries walked in cell-sorted order.  Rows of the (2k+1)^2 neighbourhood are pruned by their y/z distance and their
// x-extent trimmed to the chord of the sphere; every candidate is tested in fp64 with the reference's operation
// order (d2 < r*r, strict).  A shared-memory staged variant (TMA bulk copies + fp32 screening, as the NN sweep uses)
// was measured SLOWER here (9.9 ms vs 6.5 ms on 10 M points, profiles/r01_mme_tile_experiment.txt): with ~50 accepted
// neighbours per query the accept path needs the exact fp64 record of nearly every candidate, so the shared-memory
// pipe (random 32-byte reads, bank conflicts) becomes the limiter instead of L1.
#include "common.cuh"
#include <algorithm>
#include <cstring>

namespace me {

static constexpr int kThreads = 128;

struct MmeAcc {
  unsigned long long n_valid, n_query;
  double sum;
  unsigned long long min_enc, max_enc;   // ordered encodings of the extrema over entropies != 0
};

__global__ void mme_init_kernel(MmeAcc *a) {
  a->n_valid = 0; a->n_query = 0; a->sum = 0.0;
  a->min_enc = enc_ordered(INFINITY); a->max_enc = enc_ordered(-INFINITY);
}

struct Moments {
  double s1x, s1y, s1z, sxx, sxy, sxz, syy, syz, szz;
  unsigned int cnt;
  __device__ __forceinline__ void init() { s1x = s1y = s1z = sxx = sxy = sxz = syy = syz = szz = 0.0; cnt = 0; }
  __device__ __forceinline__ void add(double dx, double dy, double dz) {
    cnt++;
    s1x += dx; s1y += dy; s1z += dz;
    sxx += dx * dx; sxy += dx * dy; sxz += dx * dz;
    syy += dy * dy; syz += dy * dz; szz += dz * dz;
  }
};

struct ThreadStats {
  double sum, mn, mx;
  unsigned int valid, queries;
  __device__ __forceinline__ void init() { sum = 0.0; mn = INFINITY; mx = -INFINITY; valid = 0; queries = 0; }
};

// map_eval.cpp:1675-1697 for one query
__device__ __forceinline__ double finish_entropy(const Moments &m, int min_neighbors, ThreadStats &t) {
  t.queries++;
  if (m.cnt == 0) return 0.0;
  const long long k = (long long)m.cnt - 1;                 // erase(begin()): the query itself (:1672-1673)
  if (k < (long long)min_neighbors) return 0.0;
  const double kd = (double)k, inv = 1.0 / (double)(k - 1);
  double c[9];
  c[0] = (m.sxx - m.s1x * m.s1x / kd) * inv;
  c[1] = (m.sxy - m.s1x * m.s1y / kd) * inv;
  c[2] = (m.sxz - m.s1x * m.s1z / kd) * inv;
  c[4] = (m.syy - m.s1y * m.s1y / kd) * inv;
  c[5] = (m.syz - m.s1y * m.s1z / kd) * inv;
  c[8] = (m.szz - m.s1z * m.s1z / kd) * inv;
  c[3] = c[1]; c[6] = c[2]; c[7] = c[5];
  const double e = 0.5 * log(2 * M_PI * M_E * det3(c));     // map_eval.cpp:1434 / :1656
  if (isnan(e) || isinf(e)) return 0.0;
  t.sum += e; t.valid++;
  if (e != 0.0) { t.mn = fmin(t.mn, e); t.mx = fmax(t.mx, e); }
  return e;
}

__device__ void flush_stats(ThreadStats &t, MmeAcc *acc) {
  __shared__ double sh_sum[32], sh_min[32], sh_max[32];
  __shared__ unsigned long long sh_cnt[32], sh_q[32];
  t.sum = warp_sum(t.sum); t.mn = warp_min(t.mn); t.mx = warp_max(t.mx);
  const long long tv = warp_sum_ll((long long)t.valid), tq = warp_sum_ll((long long)t.queries);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if (lane == 0) { sh_sum[warp] = t.sum; sh_min[warp] = t.mn; sh_max[warp] = t.mx; sh_cnt[warp] = (unsigned long long)tv; sh_q[warp] = (unsigned long long)tq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0, mn = INFINITY, mx = -INFINITY;
    unsigned long long c = 0, q = 0;
    for (int w = 0; w < nwarps; ++w) { s += sh_sum[w]; mn = fmin(mn, sh_min[w]); mx = fmax(mx, sh_max[w]); c += sh_cnt[w]; q += sh_q[w]; }
    if (c) { atomicAdd(&acc->n_valid, c); atomicAdd(&acc->sum, s); }
    if (q) atomicAdd(&acc->n_query, q);
    if (mn <= mx) { atomicMin(&acc->min_enc, enc_ordered(mn)); atomicMax(&acc->max_enc, enc_ordered(mx)); }
  }
}

// the radius walk of one query over the cell-sorted cloud
__device__ __forceinline__ void walk_global(const P4 &q, const P4 *__restrict__ S, const uint32_t *__restrict__ cell_off,
                                            const Lattice &L, double r2, float rc2, int rings, Moments &m) {
  // the query's own cell comes with its record (high half of the tag)
  const unsigned int cq = cell_of(q.idx);
  const long long ix = cq % (unsigned int)L.dims[0], iy = (cq / (unsigned int)L.dims[0]) % (unsigned int)L.dims[1],
                  iz = cq / ((unsigned int)L.dims[0] * (unsigned int)L.dims[1]);
  const float ux = (float)(cell_coord_cont(q.x, L, 0) - (double)ix), uy = (float)(cell_coord_cont(q.y, L, 1) - (double)iy),
              uz = (float)(cell_coord_cont(q.z, L, 2) - (double)iz);   // position inside the own cell, [0,1)
  for (int dz = -rings; dz <= rings; ++dz) {
    const long long z = iz + dz;
    if (z < 0 || z >= L.dims[2]) continue;
    const float mz = dz == 0 ? 0.f : (dz > 0 ? (float)dz - uz : uz - (float)(dz + 1));
    const float remz = rc2 - (mz > 0.f ? mz * mz : 0.f);
    if (remz < 0.f) continue;
    for (int dy = -rings; dy <= rings; ++dy) {
      const long long y = iy + dy;
      if (y < 0 || y >= L.dims[1]) continue;
      const float my = dy == 0 ? 0.f : (dy > 0 ? (float)dy - uy : uy - (float)(dy + 1));
      const float rem = remz - (my > 0.f ? my * my : 0.f);
      if (rem < 0.f) continue;
      const float xw = sqrtf(rem) + 1e-4f;
      const int da = max((int)floorf(ux - xw), -rings), db = min((int)floorf(ux + xw), rings);
      const long long xa = max(ix + da, 0ll), xb = min(ix + db, (long long)L.dims[0] - 1);
      if (xa > xb) continue;
      const long long row = (z * L.dims[1] + y) * (long long)L.dims[0];
      const uint32_t s = __ldg(cell_off + row + xa), e = __ldg(cell_off + row + xb + 1);
      for (uint32_t j = s; j < e; ++j) {
        const P4 p = load_p4(S + j);
        const double dx = __dsub_rn(q.x, p.x), dy2 = __dsub_rn(q.y, p.y), dz2 = __dsub_rn(q.z, p.z);
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy2, dy2)), __dmul_rn(dz2, dz2));
        if (d2 < r2) m.add(dx, dy2, dz2);                     // nanoflann RadiusResultSet: strict <
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// one thread per query, queries in cell-sorted order: the 32 queries of a warp are neighbours along x in one lattice
// row, so they walk the same (2k+1)^2 rows and their candidate loads hit the same L1 lines (measured: handing the
// queries out tile by tile instead costs 20 %)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
mme_kernel(const P4 *__restrict__ S, long long q_begin, long long q_end, const uint32_t *__restrict__ cell_off,
           Lattice L, double r2, float rc2, int rings, int min_neighbors, double *__restrict__ entropy_sorted,
           MmeAcc *__restrict__ acc) {
  ThreadStats ts;
  ts.init();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = q_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < q_end; i += stride) {
    const P4 q = load_p4(S + i);
    Moments m;
    m.init();
    walk_global(q, S, cell_off, L, r2, rc2, rings, m);
    entropy_sorted[i] = finish_entropy(m, min_neighbors, ts);
  }
  flush_stats(ts, acc);
}

__global__ void unsort_f64_kernel(const P4 *__restrict__ S, long long n, const double *__restrict__ src,
                                  double *__restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long o = orig_of(__double_as_longlong(__ldg(reinterpret_cast<const double *>(S + i) + 3)));
    dst[o] = src[i];
  }
}

int run_mme(me_ctx *ctx, int which, double radius, int min_neighbors, me_mme_accum *out) {
  Cloud &c = ctx->cloud[which];
  if (c.n <= 0) return fail(ctx, ME_ERR_EMPTY, "cloud is empty");
  if (!(radius > 0)) return fail(ctx, ME_ERR_INVALID, "nn_radius must be > 0");
  ME_TRY(build_grid(ctx, which));
  StageTimer timer(ctx, which == ME_CLOUD_EST ? 4 : 5);
  long long qb, qe;
  shard_range(ctx, c.n, &qb, &qe);               // contiguous range of the cell-sorted order
  ME_TRY(ensure(ctx, (void **)&c.d_entropy, &c.cap_entropy, c.n, sizeof(double)));
  MmeAcc *acc = (MmeAcc *)ctx->d_scratch;
  mme_init_kernel<<<1, 1, 0, ctx->stream>>>(acc);
  ME_LAUNCH_CHECK(ctx);
  if (ctx->world > 1) ME_CUDA(ctx, cudaMemsetAsync(c.d_entropy, 0, (size_t)c.n * sizeof(double), ctx->stream));
  // a neighbour sits at most ceil(r/h) cells away; the -1e-9 keeps r = k*h (exactly) at k rings
  const double rings_f = std::max(1.0, std::ceil(radius / c.lat.h - 1e-9));
  if (rings_f > 1.0e6) return fail(ctx, ME_ERR_RANGE, "nn_radius spans too many lattice cells");
  const int rings = (int)rings_f;
  const double rc = radius / c.lat.h;
  const float rc2 = (float)(rc * rc * (1.0 + 1e-5) + 1e-4);   // row pruning is conservative; the point test decides
  if (qe > qb) {
    const int blocks = (int)std::min<long long>((qe - qb + kThreads - 1) / kThreads, (long long)ctx->sm_count * 64);
    mme_kernel<<<blocks, kThreads, 0, ctx->stream>>>(c.d_sorted, qb, qe, c.d_cell_off, c.lat, radius * radius, rc2, rings,
                                                    min_neighbors, c.d_entropy, acc);
    ME_LAUNCH_CHECK(ctx);
  }
  MmeAcc *h = (MmeAcc *)ctx->h_pinned;
  ME_CUDA(ctx, cudaMemcpyAsync(h, acc, sizeof(MmeAcc), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->n_query = (int64_t)h->n_query;
  out->n_valid = (int64_t)h->n_valid;
  out->sum_entropy = h->sum;
  out->min_entropy = dec_ordered(h->min_enc);
  out->max_entropy = dec_ordered(h->max_enc);
  c.entropy_valid = true;
  return ME_OK;
}

int unsort_entropy(me_ctx *ctx, int which, double *h_entropy) {
  Cloud &c = ctx->cloud[which];
  if (!c.entropy_valid) return fail(ctx, ME_ERR_INVALID, "me_get_entropies before me_eval_mme");
  ME_TRY(ensure_work(ctx, (size_t)c.n * sizeof(double)));
  double *dst = (double *)ctx->d_work;
  int blocks = (int)std::min<long long>((c.n + 255) / 256, (long long)ctx->sm_count * 16);
  unsort_f64_kernel<<<blocks, 256, 0, ctx->stream>>>(c.d_sorted, c.n, c.d_entropy, dst);
  ME_LAUNCH_CHECK(ctx);
  ME_CUDA(ctx, cudaMemcpyAsync(h_entropy, dst, (size_t)c.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ME_OK;
}

}  // namespace me
