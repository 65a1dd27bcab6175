// icp.cu — point-to-point ICP on top of the 1-NN sweep (registration_methods: 0).
//
// Replaces (reference): MapEval::performICPRegistration case 0 (map_eval.cpp:1366-1394), i.e.
// open3d::pipelines::registration::RegistrationICP(source = est, target = gt, icp_max_distance, initial_matrix,
// TransformationEstimationPointToPoint(), ICPConvergenceCriteria()) [ext, Open3D 0.15-0.17 Registration.cpp]:
//   pcd = est transformed by init;  result = correspondences(pcd)          (SearchHybrid(p, R, 1): NN kept iff d2 < R^2)
//   loop (max_iteration 30): update = Eigen::umeyama(corr, no scaling); T = update T; pcd.Transform(update);
//                            backup = result; result = correspondences(pcd);
//                            stop when |d fitness| < 1e-6 and |d inlier_rmse| < 1e-6
//   fitness = |corr| / |est|, inlier_rmse = sqrt(sum d2 / |corr|)
// and afterwards map_3d_ = map_3d_->Transform(trans) on the ORIGINAL cloud (:1392).
//
// Per iteration: the est lattice is laid out again (the cloud moved), the flat NN sweep finds the correspondences, and
// icp_accum_kernel reduces count, sum p, sum q, sum q p^T and sum d2 over them (17 doubles); the 3x3 algebra of
// umeyama runs on the host (polar factor via the eigen-decomposition of sigma^T sigma).  Point-to-plane (needs target
// normals) and generalized ICP (the configs' default: a third-party nonlinear solver) are not built.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace me {

static constexpr int kThreads = 256;

struct IcpAcc {
  unsigned long long n;
  double sp[3], sq[3], sqp[9], err2;
};

// sums about `c` (a point near the clouds: keeps the magnitudes, and with them the cancellation in sigma, small)
__global__ void __launch_bounds__(kThreads)
icp_accum_kernel(const P4 *__restrict__ Q, long long n, const int32_t *__restrict__ nn_idx, const double *__restrict__ nn_d2,
                 const double *__restrict__ gt_xyz, double r2, double cx, double cy, double cz, IcpAcc *__restrict__ acc) {
  double v[17];
#pragma unroll
  for (int k = 0; k < 17; ++k) v[k] = 0.0;
  unsigned long long cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int32_t j = __ldg(nn_idx + i);
    if (j < 0) continue;
    const double d2 = __ldg(nn_d2 + i);
    if (!(d2 < r2)) continue;                              // SearchHybrid: lower_bound(..., radius^2) keeps d2 < R^2
    const P4 p = load_p4(Q + i);
    const double px = p.x - cx, py = p.y - cy, pz = p.z - cz;
    const double qx = __ldg(gt_xyz + 3ll * j) - cx, qy = __ldg(gt_xyz + 3ll * j + 1) - cy, qz = __ldg(gt_xyz + 3ll * j + 2) - cz;
    cnt++;
    v[0] += px; v[1] += py; v[2] += pz; v[3] += qx; v[4] += qy; v[5] += qz;
    v[6] += qx * px; v[7] += qx * py; v[8] += qx * pz;
    v[9] += qy * px; v[10] += qy * py; v[11] += qy * pz;
    v[12] += qz * px; v[13] += qz * py; v[14] += qz * pz;
    v[15] += d2;
  }
  __shared__ double sh[kThreads / 32][17];
  __shared__ unsigned long long shc[kThreads / 32];
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = warp_sum(v[k]);
  cnt = (unsigned long long)warp_sum_ll((long long)cnt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) sh[warp][k] = v[k];
    shc[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double s = 0;
    for (int w = 0; w < kThreads / 32; ++w) s += sh[w][threadIdx.x];
    double *dst = &acc->sp[0];                             // sp, sq, sqp, err2 are contiguous
    if (s != 0.0) atomicAdd(dst + threadIdx.x, s);
  } else if (threadIdx.x == 32) {
    unsigned long long s = 0;
    for (int w = 0; w < kThreads / 32; ++w) s += shc[w];
    if (s) atomicAdd(&acc->n, s);
  }
}

// ---- host 3x3 algebra -------------------------------------------------------------------------------------------
static void eig3_sym_host(const double a_in[9], double w[3], double v[9]) {      // cyclic Jacobi, ascending not guaranteed
  double a[9];
  for (int i = 0; i < 9; ++i) { a[i] = a_in[i]; v[i] = (i % 4 == 0) ? 1.0 : 0.0; }
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[1] * a[1] + a[2] * a[2] + a[5] * a[5], diag = a[0] * a[0] + a[4] * a[4] + a[8] * a[8];
    if (off <= 1e-40 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p * 3 + q];
        if (apq == 0.0) continue;
        const double theta = (a[q * 3 + q] - a[p * 3 + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { const double x = a[k * 3 + p], y = a[k * 3 + q]; a[k * 3 + p] = c * x - s * y; a[k * 3 + q] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = a[p * 3 + k], y = a[q * 3 + k]; a[p * 3 + k] = c * x - s * y; a[q * 3 + k] = s * x + c * y; }
        for (int k = 0; k < 3; ++k) { const double x = v[k * 3 + p], y = v[k * 3 + q]; v[k * 3 + p] = c * x - s * y; v[k * 3 + q] = s * x + c * y; }
      }
  }
  w[0] = a[0]; w[1] = a[4]; w[2] = a[8];
}
static double det3_host(const double *m) {
  return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// Eigen::umeyama without scaling from the reduced sums: R = U S V^T of sigma = (1/n) sum (q - qm)(p - pm)^T
static void umeyama_from_sums(const IcpAcc &a, const double c[3], double upd[16]) {
  const double n = (double)a.n;
  double pm[3], qm[3], sigma[9];
  for (int k = 0; k < 3; ++k) { pm[k] = a.sp[k] / n; qm[k] = a.sq[k] / n; }
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 3; ++col) sigma[r * 3 + col] = a.sqp[r * 3 + col] / n - qm[r] * pm[col];
  // SVD through the symmetric eigen-problem of sigma^T sigma
  double ata[9], w[3], V[9];
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 3; ++col) { double s = 0; for (int k = 0; k < 3; ++k) s += sigma[k * 3 + r] * sigma[k * 3 + col]; ata[r * 3 + col] = s; }
  eig3_sym_host(ata, w, V);
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int x, int y) { return w[x] > w[y]; });
  double Vs[9], U[9], sv[3];
  for (int j = 0; j < 3; ++j) { sv[j] = std::sqrt(std::max(w[ord[j]], 0.0)); for (int k = 0; k < 3; ++k) Vs[k * 3 + j] = V[k * 3 + ord[j]]; }
  for (int j = 0; j < 3; ++j)
    for (int r = 0; r < 3; ++r) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += sigma[r * 3 + k] * Vs[k * 3 + j];
      U[r * 3 + j] = sv[j] > 0 ? s / sv[j] : 0.0;
    }
  // re-orthonormalise: the third column from the first two whenever the smallest singular value is not well separated
  if (!(sv[2] > 1e-8 * sv[0])) {
    U[2] = U[3] * U[7] - U[6] * U[4]; U[5] = U[6] * U[1] - U[0] * U[7]; U[8] = U[0] * U[4] - U[3] * U[1];
    if (det3_host(Vs) < 0) { U[2] = -U[2]; U[5] = -U[5]; U[8] = -U[8]; }      // keep det(U) det(V) > 0: no reflection is forced
  }
  double S[3] = {1, 1, 1};
  if (det3_host(U) * det3_host(Vs) < 0) S[2] = -1;
  std::memset(upd, 0, 16 * sizeof(double));
  upd[15] = 1.0;
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 3; ++col) { double s = 0; for (int k = 0; k < 3; ++k) s += U[r * 3 + k] * S[k] * Vs[col * 3 + k]; upd[r * 4 + col] = s; }
  // t = dst_mean - R src_mean, means back in absolute coordinates
  for (int r = 0; r < 3; ++r) {
    double rp = 0;
    for (int k = 0; k < 3; ++k) rp += upd[r * 4 + k] * (pm[k] + c[k]);
    upd[r * 4 + 3] = (qm[r] + c[r]) - rp;
  }
}

int run_icp(me_ctx *ctx, double max_dist, int max_iter, double rel_fitness, double rel_rmse, const double T_init[16],
            me_icp_result *out) {
  std::memset(out, 0, sizeof(*out));
  Cloud &E = ctx->cloud[ME_CLOUD_EST], &G = ctx->cloud[ME_CLOUD_GT];
  if (E.n <= 0 || G.n <= 0) return fail(ctx, ME_ERR_EMPTY, "both clouds must be set (map_eval.cpp:32-35)");
  if (ctx->world != 1) return fail(ctx, ME_ERR_INVALID, "me_icp_point_to_point needs world == 1 (the update needs all correspondences)");
  if (!E.owned) return fail(ctx, ME_ERR_INVALID, "me_icp_point_to_point needs a library-owned estimated cloud (use me_set_cloud)");
  if (!(max_dist > 0) || max_iter < 0) return fail(ctx, ME_ERR_INVALID, "bad ICP parameters");
  ME_TRY(wait_upload(ctx, ME_CLOUD_EST));
  // the original cloud: RegistrationICP iterates on a copy, the caller's cloud is transformed once at the end (:1392)
  double *orig = nullptr;
  long long cap_orig = 0;
  ME_TRY(ensure(ctx, (void **)&orig, &cap_orig, 3 * E.n, sizeof(double)));
  ME_CUDA(ctx, cudaMemcpyAsync(orig, E.d_xyz, (size_t)E.n * 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));

  double T[16];
  std::memcpy(T, T_init, sizeof(T));
  int rc = transform_cloud(ctx, ME_CLOUD_EST, T);
  me_nn_params p;
  std::memset(&p, 0, sizeof(p));
  p.icp_max_distance = max_dist;
  p.cutoff_mode = ME_CUTOFF_DIST_LT_R;
  p.pairing = ME_PAIRING_GEOMETRIC;
  p.want_full_cd = 0;
  p.directions = 1;
  IcpAcc *d_acc = (IcpAcc *)((char *)ctx->d_scratch + 1024), *h_acc = (IcpAcc *)((char *)ctx->h_pinned + 1536);
  IcpAcc res;
  std::memset(&res, 0, sizeof(res));
  double fitness = 0, rmse = 0, c[3] = {0, 0, 0};
  auto evaluate = [&]() -> int {
    me_nn_accum e2g;
    ME_TRY(run_nn(ctx, &p, &e2g, nullptr));                                  // lays the moved cloud out again, then sweeps
    for (int a = 0; a < 3; ++a) c[a] = 0.5 * (G.bbox_min[a] + G.bbox_max[a]);
    ME_CUDA(ctx, cudaMemsetAsync(d_acc, 0, sizeof(IcpAcc), ctx->stream));
    const int blocks = (int)std::min<long long>((E.n + kThreads - 1) / kThreads, (long long)ctx->sm_count * 8);
    icp_accum_kernel<<<blocks, kThreads, 0, ctx->stream>>>(E.d_sorted, E.n, E.d_nn_idx, E.d_nn_d2, G.d_xyz, max_dist * max_dist,
                                                          c[0], c[1], c[2], d_acc);
    ME_LAUNCH_CHECK(ctx);
    ME_CUDA(ctx, cudaMemcpyAsync(h_acc, d_acc, sizeof(IcpAcc), cudaMemcpyDeviceToHost, ctx->stream));
    ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    res = *h_acc;
    fitness = (double)res.n / (double)E.n;
    rmse = res.n > 0 ? std::sqrt(res.err2 / (double)res.n) : 0.0;
    return ME_OK;
  };
  int it = 0, converged = 0;
  if (rc == ME_OK) rc = evaluate();
  for (; rc == ME_OK && it < max_iter; ++it) {
    if (res.n == 0) break;                                                   // no correspondences: nothing to estimate
    double upd[16], Tn[16];
    umeyama_from_sums(res, c, upd);
    for (int r = 0; r < 4; ++r)
      for (int col = 0; col < 4; ++col) { double s = 0; for (int k = 0; k < 4; ++k) s += upd[r * 4 + k] * T[k * 4 + col]; Tn[r * 4 + col] = s; }
    std::memcpy(T, Tn, sizeof(T));
    rc = transform_cloud(ctx, ME_CLOUD_EST, upd);
    if (rc != ME_OK) break;
    const double f0 = fitness, r0 = rmse;
    rc = evaluate();
    if (rc != ME_OK) break;
    if (std::fabs(f0 - fitness) < rel_fitness && std::fabs(r0 - rmse) < rel_rmse) { ++it; converged = 1; break; }
  }
  // est := Transform(original, T) (map_eval.cpp:1392), whatever happened above
  cudaMemcpyAsync(E.d_xyz, orig, (size_t)E.n * 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
  int rc2 = rc == ME_OK ? transform_cloud(ctx, ME_CLOUD_EST, T) : ME_OK;
  cudaStreamSynchronize(ctx->stream);
  cudaFree(orig);
  if (rc != ME_OK) return rc;
  if (rc2 != ME_OK) return rc2;
  std::memcpy(out->transformation, T, sizeof(T));
  out->fitness = fitness; out->inlier_rmse = rmse; out->n_corr = (int64_t)res.n; out->iterations = it; out->converged = converged;
  return ME_OK;
}

}  // namespace me
