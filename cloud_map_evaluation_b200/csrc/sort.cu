// sort.cu — stable LSD radix sort of (64-bit key, 32-bit value) pairs, written for this library (no CUB / Thrust).
//
// Users: the sparse lattice build (grid.cu: points sorted by their 64-bit cell key when the dense cell table would not
// fit) and the voxel down-sampling (downsample.cu: 63-bit voxel keys; the STABLE order is what makes the per-voxel sums
// run in input order, like the reference's std::unordered_map accumulation).
//
// 8 bits per pass, three kernels per pass over tiles of kTile keys:
//   rs_hist_kernel     per-tile digit histogram (shared-memory atomics) -> ghist[digit][tile]
//   (exclusive scan of ghist, grid.cu's scan)                           -> global base of every (digit, tile)
//   rs_scatter_kernel  each warp owns a contiguous slice of the tile and walks it 32 keys at a time: MATCH.ANY on the
//                      digit gives a key's rank among the equal digits of its chunk, a per-warp running counter in shared
//                      memory the rank among the earlier chunks, a prefix over the warps the rank inside the tile.
// HBM-streaming integer work: every pass reads and writes 12 B per pair once, coalesced on the read side.
#include "common.cuh"
#include <algorithm>

namespace me {

static constexpr int kRsThreads = 256;
static constexpr int kRsWarps = kRsThreads / 32;
static constexpr int kRsChunks = 8;                               // chunks of 32 keys per warp
static constexpr int kRsTile = kRsThreads * kRsChunks;            // 2048 keys per block

__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const unsigned long long *__restrict__ keys, long long n, int shift, long long ntiles, uint32_t *__restrict__ ghist) {
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kRsTile;
#pragma unroll
  for (int j = 0; j < kRsChunks; ++j) {
    const long long i = base + (long long)j * kRsThreads + threadIdx.x;
    if (i < n) atomicAdd(&hist[(unsigned)(__ldg(keys + i) >> shift) & 255u], 1u);
  }
  __syncthreads();
  ghist[(long long)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ vals, long long n, int shift,
                  long long ntiles, const uint32_t *__restrict__ gbase, unsigned long long *__restrict__ keys_out,
                  uint32_t *__restrict__ vals_out) {
  __shared__ uint32_t wcnt[kRsWarps][256];      // first: digit counts of each warp's slice; then: running destinations
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k = threadIdx.x; k < kRsWarps * 256; k += kRsThreads) (&wcnt[0][0])[k] = 0;
  __syncthreads();
  // warp w owns keys [base + w * 32 * kRsChunks, ...): chunk c = 32 consecutive keys
  const long long wbase = (long long)blockIdx.x * kRsTile + (long long)warp * 32 * kRsChunks;
  unsigned long long key[kRsChunks];
#pragma unroll
  for (int c = 0; c < kRsChunks; ++c) {
    const long long i = wbase + c * 32 + lane;
    key[c] = i < n ? __ldg(keys + i) : ~0ull;
    if (i < n) atomicAdd(&wcnt[warp][(unsigned)(key[c] >> shift) & 255u], 1u);
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps + the global base of (digit, tile)
  {
    const int d = threadIdx.x;
    uint32_t run = __ldg(gbase + (long long)d * ntiles + blockIdx.x);
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) { const uint32_t c = wcnt[w][d]; wcnt[w][d] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kRsChunks; ++c) {
    const long long i = wbase + c * 32 + lane;
    const bool live = i < n;
    const unsigned d = live ? ((unsigned)(key[c] >> shift) & 255u) : 256u + (unsigned)lane;      // dead lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(peers) - 1;
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t dst = 0;
    if (live && lane == leader) { dst = wcnt[warp][d]; wcnt[warp][d] = dst + __popc(peers); }
    dst = __shfl_sync(0xffffffffu, dst, leader) + rank;
    if (live) {
      keys_out[dst] = key[c];
      vals_out[dst] = __ldg(vals + i);
    }
    __syncwarp();
  }
}

// Sorts n (key, value) pairs by the key bits [0, key_bits), stable.  The pairs ping-pong between (keys, vals) and
// (keys_tmp, vals_tmp); *keys_sorted / *vals_sorted point at the buffers that hold the result.
int radix_sort_pairs(me_ctx *ctx, unsigned long long *keys, uint32_t *vals, unsigned long long *keys_tmp, uint32_t *vals_tmp,
                     long long n, int key_bits, unsigned long long **keys_sorted, uint32_t **vals_sorted) {
  *keys_sorted = keys; *vals_sorted = vals;
  if (n <= 1) return ME_OK;
  if (n >= 0xffffffffll) return fail(ctx, ME_ERR_RANGE, "radix sort: more than 2^32-1 elements");
  const long long ntiles = (n + kRsTile - 1) / kRsTile;
  uint32_t *ghist = nullptr;
  ME_TRY(ensure(ctx, (void **)&ctx->d_rs_hist, &ctx->cap_rs_hist, 256 * ntiles, sizeof(uint32_t)));
  ghist = ctx->d_rs_hist;
  unsigned long long *ka = keys, *kb = keys_tmp;
  uint32_t *va = vals, *vb = vals_tmp;
  for (int shift = 0; shift < key_bits; shift += 8) {
    rs_hist_kernel<<<(unsigned)ntiles, kRsThreads, 0, ctx->stream>>>(ka, n, shift, ntiles, ghist);
    ME_LAUNCH_CHECK(ctx);
    ME_TRY(exclusive_scan_inplace(ctx, ghist, 256 * ntiles));
    rs_scatter_kernel<<<(unsigned)ntiles, kRsThreads, 0, ctx->stream>>>(ka, va, n, shift, ntiles, ghist, kb, vb);
    ME_LAUNCH_CHECK(ctx);
    std::swap(ka, kb);
    std::swap(va, vb);
  }
  *keys_sorted = ka; *vals_sorted = va;
  return ME_OK;
}

}  // namespace me
