// flat.cuh — per-thread run tables in shared memory and the flattened candidate walk shared by the NN and MME sweeps.
//
// A query's neighbourhood in the cell-sorted reference cloud is a set of x-runs, one per lattice row (y, z) of the
// searched block.  Each thread writes the runs of ITS query to a private column of a shared-memory table
// ([slot][thread]: conflict-free) and then walks all candidates of all runs as one flattened sequence, so that the lanes
// of a warp — x-neighbours with near-identical run tables — stay in step whatever the occupancy of the individual rows.
// Per run the walk needs the run bounds and the row's y/z offset constants (cy, cz); two encodings:
//   E16: 16-byte entries {begin, end, cy, cz}          — refill is one LDS.128
//   E8 :  8-byte entries {begin, len << 8 | dz << 4 | dy} — half the shared memory (more L1), refill decodes cy, cz
#pragma once
#include "common.cuh"

namespace me {

static constexpr int kFlatThreads = 128;

template <bool E16>
struct RunTab;

template <>
struct RunTab<true> {
  static constexpr int kEntryBytes = 16;
  uint4 *tab;
  __device__ __forceinline__ explicit RunTab(void *smem) : tab(reinterpret_cast<uint4 *>(smem)) {}
  __device__ __forceinline__ void put(int slot, int tid, uint32_t s, uint32_t e, int /*dyi*/, int /*dzi*/, float cy, float cz) {
    tab[slot * kFlatThreads + tid] = make_uint4(s, e, __float_as_uint(cy), __float_as_uint(cz));
  }
  __device__ __forceinline__ void get(int slot, int tid, uint32_t &s, uint32_t &e, float &cy, float &cz, float /*h*/,
                                      float /*qy*/, float /*qz*/, int /*R*/) const {
    const uint4 t = tab[slot * kFlatThreads + tid];
    s = t.x; e = t.y; cy = __uint_as_float(t.z); cz = __uint_as_float(t.w);
  }
};

template <>
struct RunTab<false> {
  static constexpr int kEntryBytes = 8;
  uint2 *tab;
  __device__ __forceinline__ explicit RunTab(void *smem) : tab(reinterpret_cast<uint2 *>(smem)) {}
  // runs longer than 2^24 - 1 points are split by the caller (put_split)
  __device__ __forceinline__ void put(int slot, int tid, uint32_t s, uint32_t e, int dyi, int dzi, float /*cy*/, float /*cz*/) {
    tab[slot * kFlatThreads + tid] = make_uint2(s, ((e - s) << 8) | ((uint32_t)dzi << 4) | (uint32_t)dyi);
  }
  __device__ __forceinline__ void get(int slot, int tid, uint32_t &s, uint32_t &e, float &cy, float &cz, float h, float qy,
                                      float qz, int R) const {
    const uint2 t = tab[slot * kFlatThreads + tid];
    s = t.x; e = t.x + (t.y >> 8);
    cy = fmaf((float)((int)(t.y & 15u) - R), h, -qy);
    cz = fmaf((float)((int)((t.y >> 4) & 15u) - R), h, -qz);
  }
};

// walk state of one thread
struct RunWalk {
  uint32_t j, e;
  int r, nrun;
  float cy, cz;
  __device__ __forceinline__ void start(int n) { j = 0; e = 0; r = 0; nrun = n; cy = 0.f; cz = 0.f; }
  // next candidate of the flattened sequence: index + the y/z constants of its row; false when exhausted
  template <bool E16>
  __device__ __forceinline__ bool next(const RunTab<E16> &T, int tid, float h, float qy, float qz, int R, uint32_t &idx,
                                       float &ocy, float &ocz) {
    if (j >= e) {
      if (r >= nrun) return false;
      T.get(r, tid, j, e, cy, cz, h, qy, qz, R);
      ++r;
    }
    idx = j++; ocy = cy; ocz = cz;
    return true;
  }
};

}  // namespace me
