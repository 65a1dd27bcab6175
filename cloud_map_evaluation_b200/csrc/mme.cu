// mme.cu — mean map entropy: per-point radius neighbourhood -> 3x3 covariance -> 0.5 ln(2 pi e det).
//
// Replaces (reference, map_eval/src/map_eval.cpp):
//   :1608-1737  ComputeMeanMapEntropyUsingNormalTBB   (k >= 10, default for the estimated map)
//   :1538-1606  ComputeMeanMapEntropyUsingNormal      (k >= 10, OpenMP variant)
//   :1438-1535  ComputeMeanMapEntropy                 (k >=  5, serial, ground truth)
//   :1433-1436  ComputeEntropy, and the min/max side effect of ColorPointCloudByMME (:697-701)
//
// The reference materialises every neighbour list (KDTreeFlann::SearchRadius -> Eigen::MatrixXd(3,k)); here no list
// exists: each query streams the (2k+1)^2 lattice rows that can intersect its sphere (rows pruned by their y/z
// distance, x-extent trimmed to the chord), tests d2 < r*r in fp64 with the reference's operation order, and folds
// accepted points into nine fp64 moments taken about the query point itself (sum d, sum d d^T).  The query is its
// own nearest neighbour at d = 0 and contributes nothing to the moments, which is exactly the reference's
// "erase the first hit" (:1672-1673).  cov = (S2 - S1 S1^T / k) / (k - 1) equals the reference's centred product.
#include "common.cuh"
#include <algorithm>
#include <cstring>

namespace me {

static constexpr int kThreads = 128;

struct MmeAcc {
  unsigned long long n_valid;
  double sum;
  unsigned long long min_enc, max_enc;   // ordered encodings of the extrema over entropies != 0
};

__global__ void mme_init_kernel(MmeAcc *a) {
  a->n_valid = 0; a->sum = 0.0;
  a->min_enc = enc_ordered(INFINITY); a->max_enc = enc_ordered(-INFINITY);
}

__global__ void __launch_bounds__(kThreads)
mme_kernel(const P4 *__restrict__ S, long long q_begin, long long q_end, const uint32_t *__restrict__ cell_off,
           Lattice L, double radius, double r2, int rings, int min_neighbors, double *__restrict__ entropy_sorted,
           MmeAcc *__restrict__ acc) {
  double t_sum = 0.0, t_min = INFINITY, t_max = -INFINITY;
  unsigned int t_valid = 0;
  const double rc = radius / L.h;              // radius in cells
  const double rc2 = rc * rc * (1.0 + 1e-9) + 1e-6;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = q_begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < q_end; i += stride) {
    const P4 q = load_p4(S + i);
    const long long ix = cell_coord(q.x, L, 0), iy = cell_coord(q.y, L, 1), iz = cell_coord(q.z, L, 2);
    const double ux = cell_coord_cont(q.x, L, 0), uy = cell_coord_cont(q.y, L, 1), uz = cell_coord_cont(q.z, L, 2);
    double s1x = 0, s1y = 0, s1z = 0, sxx = 0, sxy = 0, sxz = 0, syy = 0, syz = 0, szz = 0;
    unsigned int cnt = 0;
    for (int dz = -rings; dz <= rings; ++dz) {
      const long long z = iz + dz;
      if (z < 0 || z >= L.dims[2]) continue;
      const double mz = dz == 0 ? 0.0 : (dz > 0 ? (double)z - uz : uz - (double)(z + 1));
      const double remz = rc2 - (mz > 0 ? mz * mz : 0.0);
      if (remz < 0) continue;
      for (int dy = -rings; dy <= rings; ++dy) {
        const long long y = iy + dy;
        if (y < 0 || y >= L.dims[1]) continue;
        const double my = dy == 0 ? 0.0 : (dy > 0 ? (double)y - uy : uy - (double)(y + 1));
        const double rem = remz - (my > 0 ? my * my : 0.0);
        if (rem < 0) continue;
        const double xw = sqrt(rem) + 1e-6;
        long long xa = (long long)floor(ux - xw), xb = (long long)floor(ux + xw);
        xa = max(max(xa, ix - rings), 0ll);
        xb = min(min(xb, ix + rings), (long long)L.dims[0] - 1);
        if (xa > xb) continue;
        const long long row = (z * L.dims[1] + y) * (long long)L.dims[0];
        const uint32_t s = __ldg(cell_off + row + xa), e = __ldg(cell_off + row + xb + 1);
        for (uint32_t j = s; j < e; ++j) {
          const P4 p = load_p4(S + j);
          const double dx = __dsub_rn(q.x, p.x), dy2 = __dsub_rn(q.y, p.y), dz2 = __dsub_rn(q.z, p.z);
          const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy2, dy2)), __dmul_rn(dz2, dz2));
          if (d2 < r2) {                       // nanoflann RadiusResultSet: strict <
            cnt++;
            s1x += dx; s1y += dy2; s1z += dz2;
            sxx += dx * dx; sxy += dx * dy2; sxz += dx * dz2;
            syy += dy2 * dy2; syz += dy2 * dz2; szz += dz2 * dz2;
          }
        }
      }
    }
    double ent = 0.0;
    if (cnt > 0) {
      const long long k = (long long)cnt - 1;            // erase(begin()): the query itself (:1672-1673)
      if (k >= (long long)min_neighbors) {
        const double kd = (double)k, inv = 1.0 / (double)(k - 1);
        double c[9];
        c[0] = (sxx - s1x * s1x / kd) * inv;
        c[1] = (sxy - s1x * s1y / kd) * inv;
        c[2] = (sxz - s1x * s1z / kd) * inv;
        c[4] = (syy - s1y * s1y / kd) * inv;
        c[5] = (syz - s1y * s1z / kd) * inv;
        c[8] = (szz - s1z * s1z / kd) * inv;
        c[3] = c[1]; c[6] = c[2]; c[7] = c[5];
        const double e = 0.5 * log(2 * M_PI * M_E * det3(c));   // map_eval.cpp:1434 / :1656
        if (!isnan(e) && !isinf(e)) {
          ent = e;
          t_sum += e; t_valid++;
          if (e != 0.0) { t_min = fmin(t_min, e); t_max = fmax(t_max, e); }
        }
      }
    }
    entropy_sorted[i] = ent;
  }
  // block reduction
  __shared__ double sh_sum[kThreads / 32], sh_min[kThreads / 32], sh_max[kThreads / 32];
  __shared__ unsigned long long sh_cnt[kThreads / 32];
  t_sum = warp_sum(t_sum); t_min = warp_min(t_min); t_max = warp_max(t_max);
  long long tv = warp_sum_ll((long long)t_valid);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh_sum[warp] = t_sum; sh_min[warp] = t_min; sh_max[warp] = t_max; sh_cnt[warp] = (unsigned long long)tv; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0, mn = INFINITY, mx = -INFINITY;
    unsigned long long c = 0;
    for (int w = 0; w < kThreads / 32; ++w) { s += sh_sum[w]; mn = fmin(mn, sh_min[w]); mx = fmax(mx, sh_max[w]); c += sh_cnt[w]; }
    if (c) { atomicAdd(&acc->n_valid, c); atomicAdd(&acc->sum, s); }
    if (mn <= mx) { atomicMin(&acc->min_enc, enc_ordered(mn)); atomicMax(&acc->max_enc, enc_ordered(mx)); }
  }
}

__global__ void unsort_f64_kernel(const P4 *__restrict__ S, long long b, long long e, const double *__restrict__ src,
                                  double *__restrict__ dst) {
  for (long long i = b + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < e; i += (long long)gridDim.x * blockDim.x) {
    long long o = __double_as_longlong(__ldg(reinterpret_cast<const double *>(S + i) + 3));
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
  shard_range(ctx, c.n, &qb, &qe);
  ME_TRY(ensure(ctx, (void **)&c.d_entropy, &c.cap_entropy, c.n, sizeof(double)));
  MmeAcc *acc = (MmeAcc *)ctx->d_scratch;
  mme_init_kernel<<<1, 1, 0, ctx->stream>>>(acc);
  ME_LAUNCH_CHECK(ctx);
  const double rings_f = std::ceil(radius / c.lat.h + 1e-9);
  if (rings_f > 1.0e6) return fail(ctx, ME_ERR_RANGE, "nn_radius spans too many lattice cells");
  const int rings = (int)rings_f;
  if (qe > qb) {
    int blocks = (int)std::min<long long>((qe - qb + kThreads - 1) / kThreads, (long long)ctx->sm_count * 64);
    mme_kernel<<<blocks, kThreads, 0, ctx->stream>>>(c.d_sorted, qb, qe, c.d_cell_off, c.lat, radius, radius * radius,
                                                    rings, min_neighbors, c.d_entropy, acc);
    ME_LAUNCH_CHECK(ctx);
  }
  MmeAcc *h = (MmeAcc *)ctx->h_pinned;
  ME_CUDA(ctx, cudaMemcpyAsync(h, acc, sizeof(MmeAcc), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out->n_query = qe - qb;
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
  long long qb, qe;
  shard_range(ctx, c.n, &qb, &qe);
  ME_TRY(ensure_work(ctx, (size_t)c.n * sizeof(double)));
  double *dst = (double *)ctx->d_work;
  ME_CUDA(ctx, cudaMemsetAsync(dst, 0, (size_t)c.n * sizeof(double), ctx->stream));
  if (qe > qb) {
    int blocks = (int)std::min<long long>((qe - qb + 255) / 256, (long long)ctx->sm_count * 16);
    unsort_f64_kernel<<<blocks, 256, 0, ctx->stream>>>(c.d_sorted, qb, qe, c.d_entropy, dst);
    ME_LAUNCH_CHECK(ctx);
  }
  ME_CUDA(ctx, cudaMemcpyAsync(h_entropy, dst, (size_t)c.n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ME_OK;
}

}  // namespace me
