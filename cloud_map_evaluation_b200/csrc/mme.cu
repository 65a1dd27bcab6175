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
// Tile kernel (radius <= 2 cells): one CTA = one 4x4x4-cell query tile.  The (4+2K)^3 cells around it are (4+2K)^2
// contiguous x-runs of the cell-sorted cloud, each brought into shared memory by one TMA bulk copy (raw 32-byte
// records) and re-expressed as fp32 offsets from the tile centre.  Each thread owns one query: rows are pruned by
// their y/z distance, the x-extent trimmed to the chord, candidates screened in fp32 against r^2 -/+ the fp32 error
// bound (the sliver in between is decided in fp64 with the reference's operation order), and accepted points are
// accumulated in fp64 from the raw records.  Larger radii fall back to the same walk over global memory.
#include "common.cuh"
#include "tile.cuh"
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

// the radius walk straight over global memory (any radius; fallback of the tile kernel)
__device__ __forceinline__ void walk_global(const P4 &q, const P4 *__restrict__ S, const uint32_t *__restrict__ cell_off,
                                            const Lattice &L, double r2, float rc2, int rings, Moments &m) {
  const long long ix = cell_coord(q.x, L, 0), iy = cell_coord(q.y, L, 1), iz = cell_coord(q.z, L, 2);
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
// global-memory kernel (radius spanning more than 2 cells), one thread per query, tiles sharded like the tile kernel
// ---------------------------------------------------------------------------------------------------------------
static constexpr int kSegs = kTileEdge * kTileEdge;

__global__ void __launch_bounds__(kThreads)
mme_global_kernel(const P4 *__restrict__ S, const uint32_t *__restrict__ cell_off, Lattice L,
                  const uint32_t *__restrict__ tiles, long long t_begin, double r2, float rc2, int rings,
                  int min_neighbors, double *__restrict__ entropy_sorted, MmeAcc *__restrict__ acc) {
  __shared__ uint32_t seg_g0[kSegs], seg_pref[kSegs + 1];
  const int tid = threadIdx.x;
  const uint32_t tile = tiles[t_begin + blockIdx.x];
  const int bx = (int)(tile % L.nb[0]), by = (int)((tile / L.nb[0]) % L.nb[1]), bz = (int)(tile / ((uint32_t)L.nb[0] * L.nb[1]));
  if (tid < kSegs) {
    const int y = by * kTileEdge + tid % kTileEdge, z = bz * kTileEdge + tid / kTileEdge;
    uint32_t g0 = 0, n = 0;
    if (y < L.dims[1] && z < L.dims[2]) {
      const long long row = ((long long)z * L.dims[1] + y) * L.dims[0];
      const int xa = bx * kTileEdge, xb = min(xa + kTileEdge, L.dims[0]);
      g0 = __ldg(cell_off + row + xa);
      n = __ldg(cell_off + row + xb) - g0;
    }
    seg_g0[tid] = g0; seg_pref[tid + 1] = n;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    seg_pref[0] = 0;
    for (int s = 0; s < kSegs; ++s) { const uint32_t c = seg_pref[s + 1]; seg_pref[s + 1] = run + c; run += c; }
  }
  __syncthreads();
  const uint32_t nq = seg_pref[kSegs];
  ThreadStats ts;
  ts.init();
  for (uint32_t qi = tid; qi < nq; qi += kThreads) {
    int seg = 0;
#pragma unroll
    for (int s = 1; s < kSegs; ++s) seg += (qi >= seg_pref[s]) ? 1 : 0;
    const uint32_t pos = seg_g0[seg] + (qi - seg_pref[seg]);
    const P4 q = load_p4(S + pos);
    Moments m;
    m.init();
    walk_global(q, S, cell_off, L, r2, rc2, rings, m);
    entropy_sorted[pos] = finish_entropy(m, min_neighbors, ts);
  }
  flush_stats(ts, acc);
}

// ---------------------------------------------------------------------------------------------------------------
// tile kernel, K = radius in cells (1 or 2)
// ---------------------------------------------------------------------------------------------------------------
template <int K> struct MmeTile {
  static constexpr int W = kTileEdge + 2 * K;      // region cells per axis
  static constexpr int Rows = W * W;
  static constexpr int Cap = K == 1 ? 512 : 1024;  // staged candidates (48 B each)
};

template <int K>
__global__ void __launch_bounds__(kTileThreads)
mme_tile_kernel(const P4 *__restrict__ S, const uint32_t *__restrict__ cell_off, Lattice L,
                const uint32_t *__restrict__ tiles, long long t_begin, double radius, double r2, float rc2,
                int min_neighbors, double *__restrict__ entropy_sorted, MmeAcc *__restrict__ acc) {
  using T = MmeTile<K>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  P4 *raw = reinterpret_cast<P4 *>(smem_raw);
  float4 *rel = reinterpret_cast<float4 *>(smem_raw + (size_t)T::Cap * sizeof(P4));
  __shared__ uint32_t row_g0[T::Rows];
  __shared__ uint32_t row_pref[T::Rows + 1];
  __shared__ uint16_t cell_rel[T::Rows][T::W + 1];
  __shared__ uint32_t seg_g0[kSegs], seg_pref[kSegs + 1];
  __shared__ __align__(8) uint64_t mbar;

  const int tid = threadIdx.x;
  const uint32_t tile = tiles[t_begin + blockIdx.x];
  const int bx = (int)(tile % L.nb[0]), by = (int)((tile / L.nb[0]) % L.nb[1]), bz = (int)(tile / ((uint32_t)L.nb[0] * L.nb[1]));
  const long long r0x = (long long)bx * kTileEdge - K, r0y = (long long)by * kTileEdge - K, r0z = (long long)bz * kTileEdge - K;

  if (tid == 0) mbar_init(&mbar, 1);
  uint32_t my_cnt = 0;
  if (tid < T::Rows) {
    const long long y = r0y + tid % T::W, z = r0z + tid / T::W;
    uint32_t g0 = 0;
    const long long xa = max(r0x, 0ll), xb = min(r0x + T::W - 1, (long long)L.dims[0] - 1);
    if (y >= 0 && y < L.dims[1] && z >= 0 && z < L.dims[2] && xa <= xb) {
      const long long row = (z * L.dims[1] + y) * (long long)L.dims[0];
      g0 = __ldg(cell_off + row + xa);
#pragma unroll
      for (int c = 0; c <= T::W; ++c) {
        long long x = r0x + c;
        x = x < xa ? xa : (x > xb + 1 ? xb + 1 : x);
        cell_rel[tid][c] = (uint16_t)min(__ldg(cell_off + row + x) - g0, 0xffffu);
      }
      my_cnt = __ldg(cell_off + row + xb + 1) - g0;
    } else {
#pragma unroll
      for (int c = 0; c <= T::W; ++c) cell_rel[tid][c] = 0;
    }
    row_g0[tid] = g0;
    row_pref[tid + 1] = my_cnt;
  }
  if (tid >= 96 && tid < 96 + kSegs) {
    const int s = tid - 96;
    const int y = by * kTileEdge + s % kTileEdge, z = bz * kTileEdge + s / kTileEdge;
    uint32_t g0 = 0, n = 0;
    if (y < L.dims[1] && z < L.dims[2]) {
      const long long row = ((long long)z * L.dims[1] + y) * L.dims[0];
      const int xa = bx * kTileEdge, xb = min(xa + kTileEdge, L.dims[0]);
      g0 = __ldg(cell_off + row + xa);
      n = __ldg(cell_off + row + xb) - g0;
    }
    seg_g0[s] = g0; seg_pref[s + 1] = n;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    row_pref[0] = 0;
    for (int r = 0; r < T::Rows; ++r) { const uint32_t c = row_pref[r + 1]; row_pref[r + 1] = run + c; run += c; }
  } else if (tid == 32) {
    uint32_t run = 0;
    seg_pref[0] = 0;
    for (int s = 0; s < kSegs; ++s) { const uint32_t c = seg_pref[s + 1]; seg_pref[s + 1] = run + c; run += c; }
  }
  __syncthreads();
  const uint32_t nc = row_pref[T::Rows], nq = seg_pref[kSegs];
  const bool staged = nc <= (uint32_t)T::Cap;
  const double ocx = ((double)L.k_lo[0] * L.v) + ((double)bx * kTileEdge + 0.5 * kTileEdge) * L.h;
  const double ocy = ((double)L.k_lo[1] * L.v) + ((double)by * kTileEdge + 0.5 * kTileEdge) * L.h;
  const double ocz = ((double)L.k_lo[2] * L.v) + ((double)bz * kTileEdge + 0.5 * kTileEdge) * L.h;
  if (staged && nc > 0) {
    if (tid == 0) mbar_expect_tx(&mbar, nc * (uint32_t)sizeof(P4));
    if (tid < T::Rows && my_cnt > 0)
      tma_bulk_g2s(raw + row_pref[tid], S + row_g0[tid], my_cnt * (uint32_t)sizeof(P4), &mbar);
    mbar_wait(&mbar, 0);
    for (uint32_t i = tid; i < nc; i += kTileThreads) {
      const P4 p = raw[i];
      rel[i] = make_float4((float)(p.x - ocx), (float)(p.y - ocy), (float)(p.z - ocz), 0.f);
    }
    __syncthreads();
  }

  // fp32 screening band around r^2: |rel| <= (2 + K) h per axis
  const float eta = (float)(7.0 * (2.0 + K) * L.h * 5.9604645e-8);
  const float r2f = (float)r2;
  const float band = 1.5f * (2.f * (float)radius * eta + eta * eta + 1e-6f * r2f);
  const float r2_lo = r2f - band, r2_hi = r2f + band;

  ThreadStats ts;
  ts.init();
  for (uint32_t qb = 0; qb < nq; qb += kTileThreads) {
    const uint32_t qi = qb + tid;
    if (qi >= nq) break;
    int seg = 0;
#pragma unroll
    for (int s = 1; s < kSegs; ++s) seg += (qi >= seg_pref[s]) ? 1 : 0;
    const uint32_t pos = seg_g0[seg] + (qi - seg_pref[seg]);
    const P4 q = load_p4(S + pos);
    Moments m;
    m.init();
    if (staged) {
      const int qcx = (int)(cell_of(q.idx) % (uint32_t)L.dims[0]);
      const int lx = qcx - bx * kTileEdge + K, ly = seg % kTileEdge + K, lz = seg / kTileEdge + K;   // region-local cell
      const float qx = (float)(q.x - ocx), qy = (float)(q.y - ocy), qz = (float)(q.z - ocz);
      // position inside the own cell in cell units, from the fp32 offsets (tile centre = cell boundary 2)
      const float inv_h = (float)(1.0 / L.h);
      const float ux = qx * inv_h + 0.5f * kTileEdge - (float)(lx - K), uy = qy * inv_h + 0.5f * kTileEdge - (float)(ly - K),
                  uz = qz * inv_h + 0.5f * kTileEdge - (float)(lz - K);
#pragma unroll 1
      for (int dz = -K; dz <= K; ++dz) {
        const float mz = dz == 0 ? 0.f : (dz > 0 ? (float)dz - uz : uz - (float)(dz + 1));
        const float remz = rc2 - (mz > 0.f ? mz * mz : 0.f);
        if (remz < 0.f) continue;
#pragma unroll 1
        for (int dy = -K; dy <= K; ++dy) {
          const float my = dy == 0 ? 0.f : (dy > 0 ? (float)dy - uy : uy - (float)(dy + 1));
          const float rem = remz - (my > 0.f ? my * my : 0.f);
          if (rem < 0.f) continue;
          const float xw = sqrtf(rem) + 1e-4f;
          const int da = max((int)floorf(ux - xw), -K), db = min((int)floorf(ux + xw), K);
          const int rr = (lz + dz) * T::W + (ly + dy);
          const uint32_t base = row_pref[rr];
          const uint32_t so = base + cell_rel[rr][lx + da], eo = base + cell_rel[rr][lx + db + 1];
          for (uint32_t j = so; j < eo; ++j) {
            const float4 c = rel[j];
            const float ddx = c.x - qx, ddy = c.y - qy, ddz = c.z - qz;
            const float d32 = fmaf(ddz, ddz, fmaf(ddy, ddy, ddx * ddx));
            if (d32 < r2_hi) {
              const P4 p = raw[j];
              const double dx = __dsub_rn(q.x, p.x), dy2 = __dsub_rn(q.y, p.y), dz2 = __dsub_rn(q.z, p.z);
              bool in = true;
              if (d32 > r2_lo) {   // inside the error band: decide exactly as nanoflann does
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy2, dy2)), __dmul_rn(dz2, dz2));
                in = d2 < r2;
              }
              if (in) m.add(dx, dy2, dz2);
            }
          }
        }
      }
    } else {
      walk_global(q, S, cell_off, L, r2, rc2, K, m);
    }
    entropy_sorted[pos] = finish_entropy(m, min_neighbors, ts);
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

template <int K>
static int launch_tile(me_ctx *ctx, Cloud &c, long long tb, long long te, double radius, float rc2, int min_neighbors,
                       MmeAcc *acc) {
  using T = MmeTile<K>;
  const size_t dyn = (size_t)T::Cap * (sizeof(P4) + sizeof(float4));
  static bool attr_done = false;
  if (!attr_done) {
    ME_CUDA(ctx, cudaFuncSetAttribute(mme_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    attr_done = true;
  }
  mme_tile_kernel<K><<<(unsigned)(te - tb), kTileThreads, dyn, ctx->stream>>>(
      c.d_sorted, c.d_cell_off, c.lat, c.d_tiles, tb, radius, radius * radius, rc2, min_neighbors, c.d_entropy, acc);
  ME_LAUNCH_CHECK(ctx);
  return ME_OK;
}

int run_mme(me_ctx *ctx, int which, double radius, int min_neighbors, me_mme_accum *out) {
  Cloud &c = ctx->cloud[which];
  if (c.n <= 0) return fail(ctx, ME_ERR_EMPTY, "cloud is empty");
  if (!(radius > 0)) return fail(ctx, ME_ERR_INVALID, "nn_radius must be > 0");
  ME_TRY(build_grid(ctx, which));
  StageTimer timer(ctx, which == ME_CLOUD_EST ? 4 : 5);
  long long tb, te;
  shard_range(ctx, c.n_tiles, &tb, &te);
  ME_TRY(ensure(ctx, (void **)&c.d_entropy, &c.cap_entropy, c.n, sizeof(double)));
  MmeAcc *acc = (MmeAcc *)ctx->d_scratch;
  mme_init_kernel<<<1, 1, 0, ctx->stream>>>(acc);
  ME_LAUNCH_CHECK(ctx);
  if (ctx->world > 1) ME_CUDA(ctx, cudaMemsetAsync(c.d_entropy, 0, (size_t)c.n * sizeof(double), ctx->stream));
  const double rings_f = std::ceil(radius / c.lat.h + 1e-9);
  if (rings_f > 1.0e6) return fail(ctx, ME_ERR_RANGE, "nn_radius spans too many lattice cells");
  const int rings = (int)rings_f;
  const double rc = radius / c.lat.h;
  const float rc2 = (float)(rc * rc * (1.0 + 1e-5) + 1e-4);   // row pruning is conservative; the point test decides
  if (te > tb) {
    if (rings == 1) ME_TRY(launch_tile<1>(ctx, c, tb, te, radius, rc2, min_neighbors, acc));
    else if (rings == 2) ME_TRY(launch_tile<2>(ctx, c, tb, te, radius, rc2, min_neighbors, acc));
    else {
      mme_global_kernel<<<(unsigned)(te - tb), kThreads, 0, ctx->stream>>>(c.d_sorted, c.d_cell_off, c.lat, c.d_tiles, tb,
                                                                          radius * radius, rc2, rings, min_neighbors,
                                                                          c.d_entropy, acc);
      ME_LAUNCH_CHECK(ctx);
    }
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
