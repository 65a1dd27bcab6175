// api.cu — the C-ABI of libmapeval_b200.so (include/mapeval_b200.h): context lifecycle, cloud upload, host-side
// finalisation of the sum-reducible accumulators.  No CPU fallback exists: every evaluation runs the sm_100a kernels.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <thread>
#include <vector>

static thread_local std::string g_create_error;

namespace me {

int fail(me_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}

int ensure(me_ctx *ctx, void **ptr, long long *cap, long long need, size_t elem) {
  if (*ptr && *cap >= need) return ME_OK;
  if (*ptr) { cudaFree(*ptr); *ptr = nullptr; *cap = 0; }
  long long want = std::max<long long>(need, 1);
  cudaError_t e = cudaMalloc(ptr, (size_t)want * elem);
  if (e != cudaSuccess) {
    *ptr = nullptr;
    cudaGetLastError();
    return fail(ctx, ME_ERR_NOMEM, std::string("cudaMalloc of ") + std::to_string((size_t)want * elem) + " bytes: " + cudaGetErrorString(e));
  }
  *cap = want;
  return ME_OK;
}

int ensure_work(me_ctx *ctx, size_t bytes) {
  if (ctx->d_work && ctx->work_bytes >= bytes) return ME_OK;
  if (ctx->d_work) { cudaStreamSynchronize(ctx->stream); cudaFree(ctx->d_work); ctx->d_work = nullptr; ctx->work_bytes = 0; }
  size_t want = std::max<size_t>(bytes, 1 << 20);
  cudaError_t e = cudaMalloc(&ctx->d_work, want);
  if (e != cudaSuccess) {
    ctx->d_work = nullptr;
    cudaGetLastError();
    return fail(ctx, ME_ERR_NOMEM, std::string("cudaMalloc(work) of ") + std::to_string(want) + " bytes: " + cudaGetErrorString(e));
  }
  ctx->work_bytes = want;
  return ME_OK;
}

void invalidate_cloud(Cloud &c) {
  c.grid_valid = false; c.bbox_valid = false; c.nn_valid = false; c.entropy_valid = false; c.entropy_caller_valid = false;
}
static void invalidate(Cloud &c) {      // a new cloud: its normals go as well
  invalidate_cloud(c);
  c.normal_valid = false;
}

static void free_cloud(Cloud &c) {
  if (c.owned && c.d_xyz) cudaFree(c.d_xyz);
  if (c.d_sorted) cudaFree(c.d_sorted);
  if (c.d_rel) cudaFree(c.d_rel);
  if (c.d_cell_off) cudaFree(c.d_cell_off);
  if (c.d_cell_id) cudaFree(c.d_cell_id);
  if (c.d_coarse) cudaFree(c.d_coarse);
  if (c.d_hkey) cudaFree(c.d_hkey);
  if (c.d_hval) cudaFree(c.d_hval);
  if (c.d_nn_idx) cudaFree(c.d_nn_idx);
  if (c.d_nn_d2) cudaFree(c.d_nn_d2);
  if (c.d_nn_sq) cudaFree(c.d_nn_sq);
  if (c.d_entropy) cudaFree(c.d_entropy);
  if (c.d_entropy_caller) cudaFree(c.d_entropy_caller);
  if (c.d_normal) cudaFree(c.d_normal);
  if (c.d_tiles) cudaFree(c.d_tiles);
  if (c.d_tile_pos) cudaFree(c.d_tile_pos);
  if (c.upload_done) cudaEventDestroy(c.upload_done);
  c = Cloud();
}

}  // namespace me

using namespace me;

static constexpr size_t kScratchBytes = 4096;

// ---- upload from PAGEABLE host memory ------------------------------------------------------------------------------
// cudaMemcpyAsync from pageable memory goes through the driver's single staging buffer (~7 GB/s on this box: 65 ms for the
// 480 MB of the 10 M vs 10 M pair, 4x the whole metric pass).  A std::vector / numpy caller cannot be asked to pin its
// clouds, so the library stages them itself: kStageThreads host threads copy interleaved chunks into pinned bounce
// buffers (two per thread) and push each chunk with its own async H2D copy; the cloud's upload event is recorded once
// all chunks are queued.  Like the driver's path, the call returns when the caller's buffer has been read.
static constexpr int kStageThreads = 8;
static constexpr size_t kStageChunk = 8u << 20;

struct StagePool {
  void *pinned[kStageThreads][2] = {};
  cudaStream_t stream[kStageThreads] = {};
  cudaEvent_t done[kStageThreads][2] = {};
  bool ok = false;
};

static StagePool *stage_pool(me_ctx *ctx) {
  if (ctx->stage) return (StagePool *)ctx->stage;
  StagePool *p = new (std::nothrow) StagePool();
  if (!p) return nullptr;
  bool ok = true;
  for (int t = 0; ok && t < kStageThreads; ++t) {
    ok = cudaStreamCreateWithFlags(&p->stream[t], cudaStreamNonBlocking) == cudaSuccess;
    for (int b = 0; ok && b < 2; ++b)
      ok = cudaMallocHost(&p->pinned[t][b], kStageChunk) == cudaSuccess &&
           cudaEventCreateWithFlags(&p->done[t][b], cudaEventDisableTiming) == cudaSuccess;
  }
  p->ok = ok;
  ctx->stage = p;
  if (!ok) cudaGetLastError();
  return p;
}

static void stage_pool_free(me_ctx *ctx) {
  StagePool *p = (StagePool *)ctx->stage;
  if (!p) return;
  for (int t = 0; t < kStageThreads; ++t) {
    if (p->stream[t]) { cudaStreamSynchronize(p->stream[t]); cudaStreamDestroy(p->stream[t]); }
    for (int b = 0; b < 2; ++b) {
      if (p->pinned[t][b]) cudaFreeHost(p->pinned[t][b]);
      if (p->done[t][b]) cudaEventDestroy(p->done[t][b]);
    }
  }
  delete p;
  ctx->stage = nullptr;
}

// true if the staged upload ran (the copy stream then waits for every chunk); false = use the plain cudaMemcpyAsync
static bool staged_upload(me_ctx *ctx, void *dst, const void *src, size_t bytes) {
  if (bytes < 4 * kStageChunk || getenv("ME_NO_STAGED_UPLOAD")) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, src) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeUnregistered) return false;      // pinned / managed / device memory: the direct copy is async
  StagePool *p = stage_pool(ctx);
  if (!p || !p->ok) return false;
  const size_t nchunks = (bytes + kStageChunk - 1) / kStageChunk;
  const int dev = ctx->device;
  std::vector<int> rc(kStageThreads, 0);
  std::vector<std::thread> th;
  for (int t = 0; t < kStageThreads; ++t)
    th.emplace_back([=, &rc] {
      if (cudaSetDevice(dev) != cudaSuccess) { rc[t] = 1; return; }
      int use = 0;
      for (size_t c = (size_t)t; c < nchunks; c += kStageThreads, ++use) {
        const int b = use & 1;
        const size_t off = c * kStageChunk, len = std::min(kStageChunk, bytes - off);
        // the bounce buffer is free once the copy that last used it has run (also one queued by an earlier call; an event
        // never recorded counts as complete)
        if (cudaEventSynchronize(p->done[t][b]) != cudaSuccess) { rc[t] = 1; return; }
        std::memcpy(p->pinned[t][b], (const char *)src + off, len);
        if (cudaMemcpyAsync((char *)dst + off, p->pinned[t][b], len, cudaMemcpyHostToDevice, p->stream[t]) != cudaSuccess ||
            cudaEventRecord(p->done[t][b], p->stream[t]) != cudaSuccess) { rc[t] = 1; return; }
      }
    });
  for (auto &x : th) x.join();
  for (int t = 0; t < kStageThreads; ++t) if (rc[t]) { cudaGetLastError(); return false; }
  // the copy stream (whose event the consumers wait on) waits for the last chunk of every staging stream
  for (int t = 0; t < kStageThreads; ++t)
    for (int b = 0; b < 2; ++b)
      if (cudaStreamWaitEvent(ctx->copy_stream, p->done[t][b], 0) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

#define ME_ENTER(ctx)                                                                    \
  if (!(ctx)) return ME_ERR_INVALID;                                                     \
  do {                                                                                   \
    cudaError_t e__ = cudaSetDevice((ctx)->device);                                      \
    if (e__ != cudaSuccess) return me::fail((ctx), ME_ERR_CUDA, cudaGetErrorString(e__)); \
  } while (0)

extern "C" {

int me_abi_version(void) { return ME_ABI_VERSION; }

int me_create(const me_options *opt, me_ctx **out) {
  if (!out) return ME_ERR_INVALID;
  *out = nullptr;
  if (!opt || opt->abi_version != ME_ABI_VERSION) return fail(nullptr, ME_ERR_INVALID, "me_options.abi_version mismatch");
  if (opt->world < 1 || opt->rank < 0 || opt->rank >= opt->world) return fail(nullptr, ME_ERR_INVALID, "bad rank/world");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    return fail(nullptr, ME_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                               " (libmapeval_b200 has no CPU path)");
  }
  if (opt->device < 0 || opt->device >= ndev) return fail(nullptr, ME_ERR_NO_DEVICE, "device ordinal out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, opt->device) != cudaSuccess) return fail(nullptr, ME_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, ME_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                               std::to_string(prop.minor) + "; this library carries sm_100a code only");
  if (cudaSetDevice(opt->device) != cudaSuccess) return fail(nullptr, ME_ERR_CUDA, "cudaSetDevice failed");
  me_ctx *ctx = new (std::nothrow) me_ctx();
  if (!ctx) return fail(nullptr, ME_ERR_NOMEM, "out of host memory");
  ctx->device = opt->device;
  ctx->rank = opt->rank; ctx->world = opt->world;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->nn_cell_size = opt->nn_cell_size > 0 ? opt->nn_cell_size : 0.0;
  ctx->max_grid_cells = opt->max_grid_cells > 0 ? std::min<long long>(opt->max_grid_cells, 0xfffffff0ll) : 0;   // 0 = automatic
  ctx->voxel_hint = opt->vmd_voxel_size > 0 ? opt->vmd_voxel_size : 0.0;
  if (opt->stream) { ctx->stream = (cudaStream_t)opt->stream; ctx->own_stream = false; }
  else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return fail(nullptr, ME_ERR_CUDA, "cudaStreamCreate failed"); }
    ctx->own_stream = true;
  }
  bool ok = cudaMalloc(&ctx->d_scratch, kScratchBytes) == cudaSuccess && cudaMallocHost(&ctx->h_pinned, kScratchBytes) == cudaSuccess;
  ok = ok && cudaMalloc((void **)&ctx->d_block, kBlkTotal * sizeof(double)) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&ctx->compute_mark, cudaEventDisableTiming) == cudaSuccess;
  for (int w = 0; ok && w < 2; ++w) ok = cudaEventCreateWithFlags(&ctx->cloud[w].upload_done, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; ok && i < 2 * ME_N_STAGE_TIMES; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
  for (int i = 0; i < ME_N_STAGE_TIMES; ++i) ctx->ev_used[i] = false;
  if (!ok) { me_destroy(ctx); return fail(nullptr, ME_ERR_NOMEM, "context allocation failed"); }
  ctx->scratch_bytes = kScratchBytes;
  *out = ctx;
  return ME_OK;
}

void me_destroy(me_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->compute_mark) cudaEventDestroy(ctx->compute_mark);
  free_cloud(ctx->cloud[0]);
  free_cloud(ctx->cloud[1]);
  stage_pool_free(ctx);
  if (ctx->d_scratch) cudaFree(ctx->d_scratch);
  if (ctx->d_block) cudaFree(ctx->d_block);
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  if (ctx->d_work) cudaFree(ctx->d_work);
  if (ctx->d_scan_tmp) cudaFree(ctx->d_scan_tmp);
  if (ctx->d_rs_hist) cudaFree(ctx->d_rs_hist);
  for (int i = 0; i < 2 * ME_N_STAGE_TIMES; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char *me_last_error(const me_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int me_set_stream(me_ctx *ctx, void *cuda_stream) {
  ME_ENTER(ctx);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
  if (cuda_stream) ctx->stream = (cudaStream_t)cuda_stream;
  else {
    ME_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
  }
  return ME_OK;
}

int me_set_shard(me_ctx *ctx, int32_t rank, int32_t world) {
  ME_ENTER(ctx);
  if (world < 1 || rank < 0 || rank >= world) return fail(ctx, ME_ERR_INVALID, "bad rank/world");
  ctx->rank = rank; ctx->world = world;
  ctx->slab_planned = false; ctx->slab_on = false; ctx->vox_open = false;
  ctx->cloud[0].grid_valid = ctx->cloud[1].grid_valid = false;      // slab lattices belong to one (rank, world)
  ctx->cloud[0].nn_valid = ctx->cloud[1].nn_valid = false;
  ctx->cloud[0].entropy_valid = ctx->cloud[1].entropy_valid = false;
  ctx->cloud[0].entropy_caller_valid = ctx->cloud[1].entropy_caller_valid = false;
  ctx->cloud[0].shard_valid = ctx->cloud[1].shard_valid = false;
  return ME_OK;
}

int me_set_layout(me_ctx *ctx, int32_t layout) {
  ME_ENTER(ctx);
  if (layout != ME_LAYOUT_REPLICATED && layout != ME_LAYOUT_SLAB) return fail(ctx, ME_ERR_INVALID, "bad layout");
  const bool want = layout == ME_LAYOUT_SLAB;
  if (want == ctx->slab_request) return ME_OK;
  ctx->slab_request = want;
  ctx->slab_planned = false; ctx->slab_on = false; ctx->vox_open = false;
  for (int w = 0; w < 2; ++w) {      // the lattices are laid out again under the new layout
    ctx->cloud[w].grid_valid = false; ctx->cloud[w].nn_valid = false; ctx->cloud[w].entropy_valid = false;
    ctx->cloud[w].shard_valid = false;
  }
  return ME_OK;
}

int me_plan_lattice(const double bbox_min[3], const double bbox_max[3], int64_t n, const double *other_bbox_min,
                    const double *other_bbox_max, int64_t other_n, double voxel_size, double cell_edge_target,
                    int64_t max_grid_cells, int32_t allow_sparse, me_lattice_plan *out) {
  if (!bbox_min || !bbox_max || !out || n < 1) return ME_ERR_INVALID;
  for (int a = 0; a < 3; ++a)
    if (!(bbox_min[a] <= bbox_max[a])) return ME_ERR_INVALID;
  me::Lattice L;
  if (!me::plan_lattice_host(bbox_min, bbox_max, n, other_bbox_min, other_bbox_max, other_n, voxel_size, cell_edge_target,
                             max_grid_cells, allow_sparse != 0, &L))
    return ME_ERR_RANGE;
  out->v = L.v; out->h = L.h; out->m = L.m; out->sparse = L.sparse;
  for (int a = 0; a < 3; ++a) { out->nvox[a] = L.nvox[a]; out->dims[a] = L.dims[a]; }
  out->ncells = L.sparse ? 0 : L.ncells;
  return ME_OK;
}

int me_plan_slab_cut(const uint64_t *plane_counts_y, int32_t n_planes_y, const uint64_t *plane_counts_z, int32_t n_planes_z,
                     int32_t cells_per_voxel, int32_t world, int32_t halo_cells, int32_t *axis, int32_t *layer_bounds,
                     double *busiest_share) {
  if (!plane_counts_y || !plane_counts_z || !axis || !layer_bounds || !busiest_share) return ME_ERR_INVALID;
  if (n_planes_y < 0 || n_planes_z < 0 || cells_per_voxel < 1 || world < 1 || halo_cells < 0) return ME_ERR_INVALID;
  static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "plane counts are 64-bit");
  int ax = 0;
  double share = 2.0;
  std::vector<int> b((size_t)world + 1, 0);
  me::slab_cut((const unsigned long long *)plane_counts_y, n_planes_y, (const unsigned long long *)plane_counts_z, n_planes_z,
               cells_per_voxel, world, halo_cells, &ax, b.data(), &share);
  if (ax && share > me::kSlabMaxShare) ax = 0;
  *axis = ax;
  *busiest_share = share;
  for (int r = 0; r <= world; ++r) layer_bounds[r] = ax ? b[r] : 0;
  return ME_OK;
}

int me_layout_active(me_ctx *ctx, int32_t *layout, int32_t *axis, int64_t n_laid_out[2], int64_t n_owned[2]) {
  ME_ENTER(ctx);
  const bool slab = (ctx->cloud[0].grid_valid && ctx->cloud[0].slab) || (ctx->cloud[1].grid_valid && ctx->cloud[1].slab);
  if (layout) *layout = slab ? ME_LAYOUT_SLAB : ME_LAYOUT_REPLICATED;
  if (axis) *axis = slab ? ctx->slab_axis : 0;
  for (int w = 0; w < 2; ++w) {
    const Cloud &c = ctx->cloud[w];
    if (n_laid_out) n_laid_out[w] = c.grid_valid ? c.ns : 0;
    if (n_owned) n_owned[w] = c.grid_valid ? c.n_owned : 0;
  }
  return ME_OK;
}

int me_synchronize(me_ctx *ctx) {
  ME_ENTER(ctx);
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ME_OK;
}

int me_set_cloud(me_ctx *ctx, int which, const double *xyz_host, int64_t n) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "which must be ME_CLOUD_EST or ME_CLOUD_GT");
  if (n < 0 || (n > 0 && !xyz_host)) return fail(ctx, ME_ERR_INVALID, "bad cloud pointer/size");
  if (n >= 0x7fffffffll) return fail(ctx, ME_ERR_RANGE, "more than 2^31-1 points per cloud (the reference indexes with int)");
  Cloud &c = ctx->cloud[which];
  if (!c.owned) { c.d_xyz = nullptr; c.cap_xyz = 0; }
  ME_TRY(ensure(ctx, (void **)&c.d_xyz, &c.cap_xyz, 3 * n, sizeof(double)));
  c.owned = true;
  c.n = n;
  invalidate(c);
  if (n > 0) {
    // the copy runs on its own stream: it may overlap kernels that work on the other cloud, but must not overtake
    // kernels already queued that still read this buffer
    ME_CUDA(ctx, cudaEventRecord(ctx->compute_mark, ctx->stream));
    ME_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->compute_mark, 0));
    // the staging streams must not overtake kernels that still read this buffer either
    bool staged = false;
    {
      StagePool *sp = (StagePool *)ctx->stage;
      if (sp && sp->ok)
        for (int t = 0; t < kStageThreads; ++t) cudaStreamWaitEvent(sp->stream[t], ctx->compute_mark, 0);
      staged = staged_upload(ctx, c.d_xyz, xyz_host, (size_t)n * 3 * sizeof(double));
    }
    if (!staged)
      ME_CUDA(ctx, cudaMemcpyAsync(c.d_xyz, xyz_host, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->copy_stream));
    ME_CUDA(ctx, cudaEventRecord(c.upload_done, ctx->copy_stream));
    c.upload_pending = true;
  }
  return ME_OK;
}

int me_set_cloud_device(me_ctx *ctx, int which, const double *xyz_device, int64_t n) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "which must be ME_CLOUD_EST or ME_CLOUD_GT");
  if (n < 0 || (n > 0 && !xyz_device)) return fail(ctx, ME_ERR_INVALID, "bad cloud pointer/size");
  if (n >= 0x7fffffffll) return fail(ctx, ME_ERR_RANGE, "more than 2^31-1 points per cloud");
  Cloud &c = ctx->cloud[which];
  if (c.owned && c.d_xyz) { cudaStreamSynchronize(ctx->stream); cudaFree(c.d_xyz); }
  c.d_xyz = const_cast<double *>(xyz_device);
  c.cap_xyz = 0;
  c.owned = false;
  c.upload_pending = false;
  c.n = n;
  invalidate(c);
  return ME_OK;
}

int me_transform(me_ctx *ctx, int which, const double T[16]) {
  ME_ENTER(ctx);
  if ((which != ME_CLOUD_EST && which != ME_CLOUD_GT) || !T) return fail(ctx, ME_ERR_INVALID, "bad arguments");
  return transform_cloud(ctx, which, T);
}

int me_voxel_downsample(me_ctx *ctx, int which, double voxel_size, int64_t *n_out) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "bad cloud id");
  return voxel_downsample(ctx, which, voxel_size, n_out);
}

int me_get_cloud(me_ctx *ctx, int which, double *xyz_host, int64_t capacity_points, int64_t *n) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "bad cloud id");
  Cloud &c = ctx->cloud[which];
  if (n) *n = c.n;
  if (!xyz_host) return ME_OK;                       // size query
  if (capacity_points < c.n) return fail(ctx, ME_ERR_INVALID, "me_get_cloud: host buffer too small");
  ME_TRY(wait_upload(ctx, which));
  if (c.n > 0) {
    ME_CUDA(ctx, cudaMemcpyAsync(xyz_host, c.d_xyz, (size_t)c.n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return ME_OK;
}

int me_icp_point_to_point(me_ctx *ctx, double max_correspondence_distance, int32_t max_iteration, double relative_fitness,
                          double relative_rmse, const double T_init[16], me_icp_result *out) {
  ME_ENTER(ctx);
  if (!T_init || !out) return fail(ctx, ME_ERR_INVALID, "null argument");
  return run_icp(ctx, ME_ICP_POINT_TO_POINT, max_correspondence_distance, max_iteration, relative_fitness, relative_rmse, T_init, out);
}

int me_icp(me_ctx *ctx, int32_t method, double max_correspondence_distance, int32_t max_iteration, double relative_fitness,
           double relative_rmse, const double T_init[16], me_icp_result *out) {
  ME_ENTER(ctx);
  if (!T_init || !out) return fail(ctx, ME_ERR_INVALID, "null argument");
  return run_icp(ctx, method, max_correspondence_distance, max_iteration, relative_fitness, relative_rmse, T_init, out);
}

int me_set_normals(me_ctx *ctx, int which, const double *normals_host, int64_t n) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "bad cloud id");
  Cloud &c = ctx->cloud[which];
  if (!normals_host || n != c.n || n <= 0) return fail(ctx, ME_ERR_INVALID, "me_set_normals: one normal per point of the cloud held by the context");
  ME_TRY(ensure(ctx, (void **)&c.d_normal, &c.cap_normal, 3 * c.n, sizeof(double)));
  ME_CUDA(ctx, cudaMemcpyAsync(c.d_normal, normals_host, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  c.normal_valid = true;
  return ME_OK;
}

int me_estimate_normals(me_ctx *ctx, int which, int32_t knn) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "bad cloud id");
  return estimate_normals(ctx, which, knn, 0);
}

int me_get_normals(me_ctx *ctx, int which, double *normals_host) {
  ME_ENTER(ctx);
  if ((which != ME_CLOUD_EST && which != ME_CLOUD_GT) || !normals_host) return fail(ctx, ME_ERR_INVALID, "bad arguments");
  Cloud &c = ctx->cloud[which];
  if (!c.normal_valid) return fail(ctx, ME_ERR_INVALID, "me_get_normals: the cloud has no normals");
  ME_CUDA(ctx, cudaMemcpyAsync(normals_host, c.d_normal, (size_t)c.n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ME_OK;
}

int me_build_grid(me_ctx *ctx, int which) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "bad cloud id");
  ctx->cloud[which].grid_valid = false;
  return build_grid(ctx, which);
}

int me_eval_nn_accum(me_ctx *ctx, const me_nn_params *p, me_nn_accum *est_to_gt, me_nn_accum *gt_to_est) {
  ME_ENTER(ctx);
  if (!p) return fail(ctx, ME_ERR_INVALID, "null params");
  if (p->cutoff_mode != ME_CUTOFF_SQDIST_LE_R && p->cutoff_mode != ME_CUTOFF_DIST_LT_R) return fail(ctx, ME_ERR_INVALID, "bad cutoff_mode");
  if (p->pairing != ME_PAIRING_AS_WRITTEN && p->pairing != ME_PAIRING_GEOMETRIC) return fail(ctx, ME_ERR_INVALID, "bad pairing");
  const int dirs = p->directions ? p->directions : 3;
  if (((dirs & 1) && !est_to_gt) || ((dirs & 2) && !gt_to_est)) return fail(ctx, ME_ERR_INVALID, "null accumulator for a requested direction");
  return run_nn(ctx, p, est_to_gt, gt_to_est);
}

// ---- device-resident accumulators: sweep -> (all-reduce on the device by the caller) -> one fetch per pass ----------------
__global__ void block_init_kernel(double *blk) {
  const int t = threadIdx.x;
  if (t < me::kBlkSumCount) blk[t] = 0.0;
  else if (t < me::kBlkTotal) blk[t] = -INFINITY;
}

extern "C" {

int me_accum_reset(me_ctx *ctx) {
  ME_ENTER(ctx);
  block_init_kernel<<<1, 64, 0, ctx->stream>>>(ctx->d_block);
  ME_LAUNCH_CHECK(ctx);
  return ME_OK;
}

int me_eval_nn_accum_device(me_ctx *ctx, const me_nn_params *p) {
  ME_ENTER(ctx);
  if (!p) return fail(ctx, ME_ERR_INVALID, "null params");
  if (p->cutoff_mode != ME_CUTOFF_SQDIST_LE_R && p->cutoff_mode != ME_CUTOFF_DIST_LT_R) return fail(ctx, ME_ERR_INVALID, "bad cutoff_mode");
  if (p->pairing != ME_PAIRING_AS_WRITTEN && p->pairing != ME_PAIRING_GEOMETRIC) return fail(ctx, ME_ERR_INVALID, "bad pairing");
  return run_nn(ctx, p, nullptr, nullptr, true);
}

int me_eval_mme_accum_device(me_ctx *ctx, int which, double radius, int32_t min_neighbors) {
  ME_ENTER(ctx);
  if (which != ME_CLOUD_EST && which != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "bad arguments");
  return run_mme(ctx, which, radius, min_neighbors, nullptr, true);
}

int me_accum_block(me_ctx *ctx, double **device_block, int32_t *n_sum, int32_t *n_max) {
  ME_ENTER(ctx);
  if (device_block) *device_block = ctx->d_block;
  if (n_sum) *n_sum = kBlkSumCount;
  if (n_max) *n_max = kBlkMaxCount;
  return ME_OK;
}

int me_voxel_begin(me_ctx *ctx, double voxel_size, int32_t min_points) {
  ME_ENTER(ctx);
  return voxel_begin(ctx, voxel_size, min_points);
}

int me_voxel_w_table(me_ctx *ctx, double **device_w, int64_t *n) {
  ME_ENTER(ctx);
  if (!device_w || !n) return fail(ctx, ME_ERR_INVALID, "null argument");
  return voxel_w_table(ctx, device_w, n);
}

int me_voxel_finish_accum_device(me_ctx *ctx, int32_t scs_radius) {
  ME_ENTER(ctx);
  return voxel_finish_block(ctx, scs_radius);
}

int me_accum_fetch_awd(me_ctx *ctx, me_awd_result *out) {
  ME_ENTER(ctx);
  if (!out) return fail(ctx, ME_ERR_INVALID, "null result");
  double *h = (double *)((char *)ctx->h_pinned + 3072);
  ME_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_block + kBlkAwd, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  auto i64 = [](double v) { return (int64_t)std::llround(v); };
  std::memset(out, 0, sizeof(*out));
  out->n_pairs = i64(h[0]); out->n_scs = i64(h[1]);
  out->n_voxels_est = i64(h[2]); out->n_voxels_gt = i64(h[3]);
  out->n_active = i64(h[4]); out->n_new = i64(h[5]);
  out->n_old = out->n_voxels_gt - out->n_active;
  out->awd = h[6] / (double)out->n_pairs;        // 0/0 -> NaN as map_eval.cpp:324
  out->scs = h[7] / (double)out->n_scs;          // map_eval.cpp:387
  return ME_OK;
}

int me_accum_fetch(me_ctx *ctx, me_nn_accum *est_to_gt, me_nn_accum *gt_to_est, me_mme_accum *mme_est, me_mme_accum *mme_gt) {
  ME_ENTER(ctx);
  double *h = (double *)((char *)ctx->h_pinned + 3072);
  ME_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_block, kBlkTotal * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  auto i64 = [](double v) { return (int64_t)std::llround(v); };
  me_nn_accum *dirs[2] = {est_to_gt, gt_to_est};
  for (int d = 0; d < 2; ++d) {
    me_nn_accum *a = dirs[d];
    if (!a) continue;
    const double *b = h + d * kBlkNN;
    a->n_query = i64(b[0]); a->n_corr = i64(b[1]);
    for (int k = 0; k < 5; ++k) a->n_inlier[k] = i64(b[2 + k]);
    a->n_ub = i64(b[7]); a->n_far = i64(b[8]);
    for (int k = 0; k < 5; ++k) { a->sum_d[k] = b[9 + k]; a->sum_d2[k] = b[14 + k]; }
    a->sum_d_all = b[19]; a->sum_d2_all = b[20]; a->sum_nn_dist = b[21];
  }
  me_mme_accum *mm[2] = {mme_est, mme_gt};
  for (int w = 0; w < 2; ++w) {
    me_mme_accum *m = mm[w];
    if (!m) continue;
    const double *b = h + kBlkMmeSum + 3 * w, *x = h + kBlkMax + 2 * w;
    m->n_query = i64(b[0]); m->n_valid = i64(b[1]); m->sum_entropy = b[2];
    m->max_entropy = x[0]; m->min_entropy = -x[1];
  }
  return ME_OK;
}

}  // extern "C"

static void finalize_dir(const me_nn_accum *a, int64_t n_source, me_dir_result *r) {
  std::memset(r, 0, sizeof(*r));
  r->n_source = n_source;
  r->n_corr = a->n_corr;
  r->n_ub = a->n_ub;
  r->sum_nn_dist = a->sum_nn_dist;
  const double nc = (double)a->n_corr;
  const int target_num = (int)n_source;                       // map_eval.cpp:1128
  for (int k = 0; k < 5; ++k) {
    r->n_inlier[k] = a->n_inlier[k];
    const double mean = a->sum_d[k] / nc;                      // :1125
    r->mean[k] = mean;
    r->rmse[k] = std::sqrt(a->sum_d2[k] / nc);                 // :1126,1131
    r->fitness[k] = (double)a->n_inlier[k] * 1.0 / target_num; // :1130
    // :1132-1138, sum over ALL kept pairs of (d - mean_k)^2, in closed form
    double ss = a->sum_d2_all - 2.0 * mean * a->sum_d_all + nc * mean * mean;
    if (ss < 0 && ss > -1e-9 * a->sum_d2_all) ss = 0;
    r->sigma[k] = std::sqrt(ss / nc);
  }
}

int me_nn_finalize(const me_nn_params *p, const me_nn_accum *est_to_gt, const me_nn_accum *gt_to_est, int64_t n_est,
                   int64_t n_gt, me_nn_result *out) {
  if (!p || !out) return ME_ERR_INVALID;
  std::memset(out, 0, sizeof(*out));
  me_nn_accum zero;
  std::memset(&zero, 0, sizeof(zero));
  const int dirs = p->directions ? p->directions : 3;
  if (dirs & 1) finalize_dir(est_to_gt ? est_to_gt : &zero, n_est, &out->est_to_gt);
  if (dirs & 2) finalize_dir(gt_to_est ? gt_to_est : &zero, n_gt, &out->gt_to_est);
  for (int k = 0; k < 5; ++k) {                                 // map_eval.cpp:1245-1253
    out->cd[k] = out->est_to_gt.rmse[k] + out->gt_to_est.rmse[k];
    const double overlap = out->est_to_gt.fitness[k], rmse = out->est_to_gt.rmse[k];
    out->f1[k] = 2 * overlap * rmse / (overlap + rmse);
    const int num_intersection = (int)out->est_to_gt.n_inlier[k];
    const int num_union = (int)(n_est + n_gt - num_intersection);
    out->iou[k] = (double)num_intersection / num_union;
  }
  out->full_cd = p->want_full_cd ? (out->est_to_gt.sum_nn_dist / (double)n_est + out->gt_to_est.sum_nn_dist / (double)n_gt) : 0.0;
  return ME_OK;
}

int me_eval_nn(me_ctx *ctx, const me_nn_params *p, me_nn_result *out) {
  ME_ENTER(ctx);
  if (!p || !out) return fail(ctx, ME_ERR_INVALID, "null argument");
  if (ctx->world != 1) return fail(ctx, ME_ERR_INVALID, "me_eval_nn needs world == 1; use me_eval_nn_accum + all-reduce + me_nn_finalize");
  me_nn_accum a, b;
  ME_TRY(me_eval_nn_accum(ctx, p, &a, &b));
  return me_nn_finalize(p, &a, &b, ctx->cloud[0].n, ctx->cloud[1].n, out);
}

int me_get_nn(me_ctx *ctx, int which_query, int32_t *nn_index, double *nn_sqdist) {
  ME_ENTER(ctx);
  if (which_query != ME_CLOUD_EST && which_query != ME_CLOUD_GT) return fail(ctx, ME_ERR_INVALID, "bad cloud id");
  return unsort_nn(ctx, which_query, nn_index, nn_sqdist);
}

int me_eval_mme_accum(me_ctx *ctx, int which, double radius, int32_t min_neighbors, me_mme_accum *out) {
  ME_ENTER(ctx);
  if ((which != ME_CLOUD_EST && which != ME_CLOUD_GT) || !out) return fail(ctx, ME_ERR_INVALID, "bad arguments");
  return run_mme(ctx, which, radius, min_neighbors, out);
}

int me_mme_finalize(const me_mme_accum *acc, int64_t n_total, me_mme_result *out) {
  if (!acc || !out) return ME_ERR_INVALID;
  out->n_total = n_total;
  out->n_valid = acc->n_valid;
  out->mme = acc->n_valid > 0 ? acc->sum_entropy / (double)acc->n_valid : 0.0;   // map_eval.cpp:1720-1724
  if (acc->min_entropy <= acc->max_entropy) {                                     // :700-701
    out->max_abs_entropy = std::fabs(acc->min_entropy);
    out->min_abs_entropy = std::fabs(acc->max_entropy);
  } else {
    out->max_abs_entropy = out->min_abs_entropy = std::numeric_limits<double>::quiet_NaN();
  }
  return ME_OK;
}

int me_get_entropies(me_ctx *ctx, int which, double *entropies_host) {
  ME_ENTER(ctx);
  if ((which != ME_CLOUD_EST && which != ME_CLOUD_GT) || !entropies_host) return fail(ctx, ME_ERR_INVALID, "bad arguments");
  return unsort_entropy(ctx, which, entropies_host);
}

int me_eval_mme(me_ctx *ctx, int which, double radius, int32_t min_neighbors, me_mme_result *out, double *entropies_host) {
  ME_ENTER(ctx);
  if (!out) return fail(ctx, ME_ERR_INVALID, "null result");
  if (ctx->world != 1) return fail(ctx, ME_ERR_INVALID, "me_eval_mme needs world == 1; use me_eval_mme_accum + all-reduce + me_mme_finalize");
  me_mme_accum acc;
  ME_TRY(me_eval_mme_accum(ctx, which, radius, min_neighbors, &acc));
  ME_TRY(me_mme_finalize(&acc, ctx->cloud[which].n, out));
  if (entropies_host) ME_TRY(unsort_entropy(ctx, which, entropies_host));
  return ME_OK;
}

int me_eval_awd(me_ctx *ctx, double voxel_size, int32_t min_points, int32_t scs_radius, me_awd_result *out,
                int64_t *n_rows, double **rows27) {
  ME_ENTER(ctx);
  if (!out) return fail(ctx, ME_ERR_INVALID, "null result");
  return run_awd(ctx, voxel_size, min_points, scs_radius, out, n_rows, rows27);
}

int me_awd_from_rows(me_ctx *ctx, const double *rows27, int64_t n_rows, double voxel_size, int32_t scs_radius, double *w_out,
                     me_awd_result *out) {
  ME_ENTER(ctx);
  if (!out) return fail(ctx, ME_ERR_INVALID, "null result");
  return run_awd_rows(ctx, rows27, n_rows, voxel_size, scs_radius, w_out, out);
}

void me_free(void *p) { std::free(p); }

int me_get_stage_times(me_ctx *ctx, double ms[ME_N_STAGE_TIMES]) {
  ME_ENTER(ctx);
  ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int s = 0; s < ME_N_STAGE_TIMES; ++s) {
    ms[s] = 0.0;
    if (!ctx->ev_used[s]) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ctx->ev[2 * s], ctx->ev[2 * s + 1]) == cudaSuccess) ms[s] = t;
    else cudaGetLastError();
  }
  return ME_OK;
}

int64_t me_launch_count(const me_ctx *ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
