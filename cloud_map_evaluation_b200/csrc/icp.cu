// icp.cu — the three registration methods of MapEval::performICPRegistration (map_eval.cpp:1366-1394) on top of the
// 1-NN sweep: registration_methods 0 = point-to-point, 1 = point-to-plane, 2 = generalized ICP (what
// every shipped config uses).
//
// Replaces (reference) open3d::pipelines::registration::RegistrationICP / RegistrationGeneralizedICP(source = est,
// target = gt, icp_max_distance, initial_matrix, estimation, ICPConvergenceCriteria()) [ext, Open3D 0.15-0.17]:
//   pcd = est transformed by init;  result = correspondences(pcd)          (SearchHybrid(p, R, 1): NN kept iff d2 < R^2)
//   loop (max_iteration 30): update = estimation.ComputeTransformation(corr); T = update T; pcd.Transform(update);
//                            backup = result; result = correspondences(pcd);
//                            stop when |d fitness| < 1e-6 and |d inlier_rmse| < 1e-6
//   fitness = |corr| / |est|, inlier_rmse = sqrt(sum d2 / |corr|) (point-to-point quantities for every estimation)
// and afterwards map_3d_ = map_3d_->Transform(trans) on the ORIGINAL cloud (:1392).
//   point-to-point:  Eigen::umeyama without scaling — icp_accum_kernel reduces count, sum p, sum q, sum q p^T, sum d2
//                    (17 doubles); the 3x3 algebra runs on the host.
//   point-to-plane:  r = (vs - vt) . nt, J = [vs x nt, nt]; needs target normals (me_set_normals / me_estimate_normals).
//   generalized:     InitializePointCloudForGeneralizedICP = EstimateNormals(KNN 20) on both clouds (knn_normals_kernel:
//                    exact k-NN by ring expansion over the lattice, covariance, Eberly's robust 3x3 eigenvector =
//                    Open3D's FastEigen3x3), covariance_i = Rx diag(1e-3, 1, 1) Rx^T = I + (eps - 1) n n^T with n := e1
//                    where Open3D's GetRotationFromE1ToX falls back to the identity (n.x < -0.99, sic) — so a unit vector
//                    per point carries the covariance, and PointCloud::Transform's R C R^T is a rotation of that vector;
//                    ComputeTransformation: M = Cs + Ct, J = M^-1/2 [-skew(vs) | I], r = M^-1/2 d, i.e. the 6x6 normal
//                    equations sum A^T M^-1 A, sum A^T M^-1 d, reduced on the device (icp_ne_accum_kernel, 28 doubles);
//                    the 6x6 solve and TransformVector6dToMatrix4d run on the host.
// Per iteration: the est lattice is laid out again (the cloud moved), the flat NN sweep finds the correspondences, one
// reduction kernel, one transform.
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

// ---------------------------------------------------------------------------------------------------------------
// Open3D FastEigen3x3 (Eberly, "A Robust Eigensolver for 3x3 Symmetric Matrices"): unit eigenvector of the smallest
// eigenvalue of a symmetric 3x3 (row-major, 9 entries); the zero vector for the zero matrix
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cross3d(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3d(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ void eberly_evec0(const double A[9], double eval0, double out[3]) {
  const double row0[3] = {A[0] - eval0, A[1], A[2]}, row1[3] = {A[1], A[4] - eval0, A[5]}, row2[3] = {A[2], A[5], A[8] - eval0};
  double r01[3], r02[3], r12[3];
  cross3d(row0, row1, r01); cross3d(row0, row2, r02); cross3d(row1, row2, r12);
  const double d0 = dot3d(r01, r01), d1 = dot3d(r02, r02), d2 = dot3d(r12, r12);
  double dmax = d0; int imax = 0;
  if (d1 > dmax) { dmax = d1; imax = 1; }
  if (d2 > dmax) imax = 2;
  const double s = sqrt(imax == 0 ? d0 : (imax == 1 ? d1 : d2));
#pragma unroll
  for (int k = 0; k < 3; ++k) out[k] = (imax == 0 ? r01[k] : (imax == 1 ? r02[k] : r12[k])) / s;
}
__device__ void eberly_evec1(const double A[9], const double e0[3], double eval1, double out[3]) {
  double U[3], V[3];
  if (fabs(e0[0]) > fabs(e0[1])) {
    const double inv = 1.0 / sqrt(e0[0] * e0[0] + e0[2] * e0[2]);
    U[0] = -e0[2] * inv; U[1] = 0; U[2] = e0[0] * inv;
  } else {
    const double inv = 1.0 / sqrt(e0[1] * e0[1] + e0[2] * e0[2]);
    U[0] = 0; U[1] = e0[2] * inv; U[2] = -e0[1] * inv;
  }
  cross3d(e0, U, V);
  const double AU[3] = {A[0] * U[0] + A[1] * U[1] + A[2] * U[2], A[1] * U[0] + A[4] * U[1] + A[5] * U[2], A[2] * U[0] + A[5] * U[1] + A[8] * U[2]};
  const double AV[3] = {A[0] * V[0] + A[1] * V[1] + A[2] * V[2], A[1] * V[0] + A[4] * V[1] + A[5] * V[2], A[2] * V[0] + A[5] * V[1] + A[8] * V[2]};
  double m00 = dot3d(U, AU) - eval1, m01 = dot3d(U, AV), m11 = dot3d(V, AV) - eval1;
  const double a00 = fabs(m00), a01 = fabs(m01), a11 = fabs(m11);
  if (a00 >= a11) {
    if (fmax(a00, a01) > 0) {
      if (a00 >= a01) { m01 /= m00; m00 = 1 / sqrt(1 + m01 * m01); m01 *= m00; }
      else { m00 /= m01; m01 = 1 / sqrt(1 + m00 * m00); m00 *= m01; }
#pragma unroll
      for (int k = 0; k < 3; ++k) out[k] = m01 * U[k] - m00 * V[k];
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) out[k] = U[k];
    }
  } else {
    if (fmax(a11, a01) > 0) {
      if (a11 >= a01) { m01 /= m11; m11 = 1 / sqrt(1 + m01 * m01); m01 *= m11; }
      else { m11 /= m01; m01 = 1 / sqrt(1 + m11 * m11); m11 *= m01; }
#pragma unroll
      for (int k = 0; k < 3; ++k) out[k] = m11 * U[k] - m01 * V[k];
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) out[k] = U[k];
    }
  }
}
__device__ void fast_eigen3x3_normal(const double cov[9], double nrm[3]) {
  double max_coeff = cov[0];
#pragma unroll
  for (int k = 1; k < 9; ++k) max_coeff = fmax(max_coeff, cov[k]);
  nrm[0] = nrm[1] = nrm[2] = 0.0;
  if (max_coeff == 0) return;
  double A[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) A[k] = cov[k] / max_coeff;
  const double norm = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
  if (norm > 0) {
    const double q = (A[0] + A[4] + A[8]) / 3;
    const double b00 = A[0] - q, b11 = A[4] - q, b22 = A[8] - q;
    const double p = sqrt((b00 * b00 + b11 * b11 + b22 * b22 + norm * 2) / 6);
    const double c00 = b11 * b22 - A[5] * A[5], c01 = A[1] * b22 - A[5] * A[2], c02 = A[1] * A[5] - b11 * A[2];
    const double det = (b00 * c00 - A[1] * c01 + A[2] * c02) / (p * p * p);
    const double half_det = fmin(fmax(det * 0.5, -1.0), 1.0);
    const double angle = acos(half_det) / 3.0;
    const double two_thirds_pi = 2.09439510239319549;
    const double beta2 = cos(angle) * 2, beta0 = cos(angle + two_thirds_pi) * 2, beta1 = -(beta0 + beta2);
    const double ev0 = q + p * beta0, ev1 = q + p * beta1, ev2 = q + p * beta2;
    double ea[3], eb[3];
    if (half_det >= 0) {
      eberly_evec0(A, ev2, ea);                                   // ea = evec2
      if (ev2 < ev0 && ev2 < ev1) { nrm[0] = ea[0]; nrm[1] = ea[1]; nrm[2] = ea[2]; return; }
      eberly_evec1(A, ea, ev1, eb);                               // eb = evec1
      if (ev1 < ev0 && ev1 < ev2) { nrm[0] = eb[0]; nrm[1] = eb[1]; nrm[2] = eb[2]; return; }
      cross3d(eb, ea, nrm);                                       // evec0 = evec1 x evec2
    } else {
      eberly_evec0(A, ev0, ea);                                   // ea = evec0
      if (ev0 < ev1 && ev0 < ev2) { nrm[0] = ea[0]; nrm[1] = ea[1]; nrm[2] = ea[2]; return; }
      eberly_evec1(A, ea, ev1, eb);                               // eb = evec1
      if (ev1 < ev0 && ev1 < ev2) { nrm[0] = eb[0]; nrm[1] = eb[1]; nrm[2] = eb[2]; return; }
      cross3d(ea, eb, nrm);                                       // evec2 = evec0 x evec1
    }
  } else {
    if (cov[0] < cov[4] && cov[0] < cov[8]) nrm[0] = 1;
    else if (cov[4] < cov[0] && cov[4] < cov[8]) nrm[1] = 1;
    else nrm[2] = 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// PointCloud::EstimateNormals(KDTreeSearchParamKNN(k)) [ext]: per point the k nearest neighbours (the point itself
// included), their covariance, the eigenvector of its smallest eigenvalue.  One thread per point; the k best (d2,
// position) pairs sit sorted in shared memory ([slot][thread]); the lattice is searched shell by shell (Chebyshev rings)
// until the k-th distance beats the distance to the border of the searched block.  Ties go to the smaller caller index.
// gicp != 0: store e1 where Open3D's GetRotationFromE1ToX falls back to the identity (n.x < -0.99).
// ---------------------------------------------------------------------------------------------------------------
static constexpr int kKnnThreads = 128;

__global__ void __launch_bounds__(kKnnThreads)
knn_normals_kernel(const P4 *__restrict__ S, long long n, CellIndex I, Lattice L, CoarseGrid CG, int k, double slack,
                   int gicp, double *__restrict__ normals) {
  extern __shared__ __align__(16) unsigned char knn_smem[];
  double *dk = reinterpret_cast<double *>(knn_smem) + threadIdx.x;                                     // dk[t * kKnnThreads]
  uint32_t *ik = reinterpret_cast<uint32_t *>(knn_smem + (size_t)k * kKnnThreads * sizeof(double)) + threadIdx.x;
  for (long long i = blockIdx.x * (long long)kKnnThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kKnnThreads) {
    const P4 q = load_p4(S + i);
    int cix, ciy, ciz;
    long long xs = I.sparse ? cell_coord(q.x, L, 0) : 0;
    xs = xs < 0 ? 0 : (xs >= L.dims[0] ? L.dims[0] - 1 : xs);
    cell_from_tag(I, q.idx, (int)xs, cix, ciy, ciz);
    const long long ix = cix, iy = ciy, iz = ciz;
    int cnt = 0;
    double kth = INFINITY;
    // one candidate: keep the k best (d2, position) pairs sorted; ties go to the smaller caller index
    auto offer = [&](uint32_t j) {
      const P4 p = load_p4(S + j);
      const double d2 = d2_kd(q.x, q.y, q.z, p.x, p.y, p.z);
      if (cnt == k) {
        if (d2 > kth) return;
        if (d2 == kth && orig_of(p.idx) > orig_of(__double_as_longlong(__ldg(reinterpret_cast<const double *>(S + ik[(k - 1) * kKnnThreads]) + 3)))) return;
      }
      int pos = cnt < k ? cnt++ : k - 1;
      while (pos > 0) {
        const double dp = dk[(pos - 1) * kKnnThreads];
        if (dp < d2) break;
        if (dp == d2 && orig_of(__double_as_longlong(__ldg(reinterpret_cast<const double *>(S + ik[(pos - 1) * kKnnThreads]) + 3))) < orig_of(p.idx)) break;
        dk[pos * kKnnThreads] = dp; ik[pos * kKnnThreads] = ik[(pos - 1) * kKnnThreads];
        --pos;
      }
      dk[pos * kKnnThreads] = d2; ik[pos * kKnnThreads] = j;
      if (cnt == k) kth = dk[(k - 1) * kKnnThreads];
    };
    bool resolved = false;
    const int fine_rings = CG.cnt ? 8 : 0x7fffffff;      // isolated points: blocks of f^3 cells take over (empty space is skipped)
    for (int r = 0; r <= fine_rings; ++r) {
      for (int dz = -r; dz <= r; ++dz) {
        const long long z = iz + dz;
        if (z < 0 || z >= L.dims[2]) continue;
        for (int dy = -r; dy <= r; ++dy) {
          const long long y = iy + dy;
          if (y < 0 || y >= L.dims[1]) continue;
          const bool border = (dz == -r || dz == r || dy == -r || dy == r);
          for (int part = 0; part < 2; ++part) {
            long long xa, xb;
            if (border) { if (part) break; xa = ix - r; xb = ix + r; }
            else { xa = xb = part ? ix + r : ix - r; }
            xa = max(xa, 0ll); xb = min(xb, (long long)L.dims[0] - 1);
            if (xa > xb) continue;
            uint32_t s, e;
            cell_range(I, (int)z, (int)y, (int)xa, (int)xb, s, e);
            for (uint32_t j = s; j < e; ++j) offer(j);
          }
        }
      }
      // everything closer than the border of the searched block [i - r, i + r]^3 has been seen
      const double g = (double)r * L.h - slack;
      if (cnt == k && g > 0 && kth < g * g) { resolved = true; break; }
      if (ix - r <= 0 && ix + r >= L.dims[0] - 1 && iy - r <= 0 && iy + r >= L.dims[1] - 1 && iz - r <= 0 && iz + r >= L.dims[2] - 1) { resolved = true; break; }
    }
    if (!resolved) {
      // coarse phase, from scratch (the blocks cover the cells of the fine phase again): Chebyshev rings of blocks, empty
      // blocks cost one count load, blocks that cannot hold a point closer than the k-th best are skipped
      const int f = CG.f;
      const long long cqx = ix / f, cqy = iy / f, cqz = iz / f;
      const double ux = cell_coord_cont(q.x, L, 0), uy = cell_coord_cont(q.y, L, 1), uz = cell_coord_cont(q.z, L, 2);
      cnt = 0; kth = INFINITY;
      for (int rc = 0;; ++rc) {
        for (int dz = -rc; dz <= rc; ++dz) {
          const long long bz = cqz + dz;
          if (bz < 0 || bz >= CG.cd[2]) continue;
          for (int dy = -rc; dy <= rc; ++dy) {
            const long long by = cqy + dy;
            if (by < 0 || by >= CG.cd[1]) continue;
            const bool border = (dz == -rc || dz == rc || dy == -rc || dy == rc);
            for (int dx = -rc; dx <= rc; dx += (border || rc == 0) ? 1 : 2 * rc) {
              const long long bx = cqx + dx;
              if (bx < 0 || bx >= CG.cd[0]) continue;
              if (__ldg(CG.cnt + (bz * CG.cd[1] + by) * (long long)CG.cd[0] + bx) == 0) continue;
              if (cnt == k) {
                const double ex = fmax(0.0, fmax((double)(bx * f) - ux, ux - (double)((bx + 1) * f)));
                const double ey = fmax(0.0, fmax((double)(by * f) - uy, uy - (double)((by + 1) * f)));
                const double ez = fmax(0.0, fmax((double)(bz * f) - uz, uz - (double)((bz + 1) * f)));
                const double lb = sqrt(ex * ex + ey * ey + ez * ez) * L.h - slack;
                if (lb > 0 && lb * lb > kth) continue;
              }
              const int xa = (int)(bx * f), xb = (int)min(bx * f + f - 1, (long long)L.dims[0] - 1);
              for (int row = 0; row < f * f; ++row) {
                const long long z = bz * f + row / f, y = by * f + row % f;
                if (z >= L.dims[2] || y >= L.dims[1]) continue;
                uint32_t s, e;
                cell_range(I, (int)z, (int)y, xa, xb, s, e);
                for (uint32_t j = s; j < e; ++j) offer(j);
              }
            }
          }
        }
        // the point lies inside its own block: everything closer than rc blocks has been seen
        const double g = (double)rc * (double)f * L.h - slack;
        if (cnt == k && g > 0 && kth < g * g) break;
        if (cqx - rc <= 0 && cqx + rc >= CG.cd[0] - 1 && cqy - rc <= 0 && cqy + rc >= CG.cd[1] - 1 && cqz - rc <= 0 && cqz + rc >= CG.cd[2] - 1) break;
      }
    }
    // utility::ComputeCovariance over the neighbours (cumulants about the query point instead of the origin: same
    // covariance, without the cancellation of raw coordinates); fewer than 3 neighbours -> identity
    double cov[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (cnt >= 3) {
      double c1[3] = {0, 0, 0}, c2[6] = {0, 0, 0, 0, 0, 0};
      for (int t = 0; t < cnt; ++t) {
        const P4 p = load_p4(S + ik[t * kKnnThreads]);
        const double dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        c1[0] += dx; c1[1] += dy; c1[2] += dz;
        c2[0] += dx * dx; c2[1] += dx * dy; c2[2] += dx * dz; c2[3] += dy * dy; c2[4] += dy * dz; c2[5] += dz * dz;
      }
      const double inv = 1.0 / (double)cnt;
#pragma unroll
      for (int a = 0; a < 3; ++a) c1[a] *= inv;
#pragma unroll
      for (int a = 0; a < 6; ++a) c2[a] *= inv;
      cov[0] = c2[0] - c1[0] * c1[0]; cov[4] = c2[3] - c1[1] * c1[1]; cov[8] = c2[5] - c1[2] * c1[2];
      cov[1] = cov[3] = c2[1] - c1[0] * c1[1]; cov[2] = cov[6] = c2[2] - c1[0] * c1[2]; cov[5] = cov[7] = c2[4] - c1[1] * c1[2];
    }
    double nr[3];
    fast_eigen3x3_normal(cov, nr);
    if (sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]) == 0.0) { nr[0] = 0; nr[1] = 0; nr[2] = 1; }
    if (gicp && nr[0] < -0.99) { nr[0] = 1; nr[1] = 0; nr[2] = 0; }
    const long long o = orig_of(q.idx);
    normals[3 * o] = nr[0]; normals[3 * o + 1] = nr[1]; normals[3 * o + 2] = nr[2];
  }
}

int estimate_normals(me_ctx *ctx, int which, int knn, int gicp) {
  Cloud &c = ctx->cloud[which];
  if (c.n <= 0) return fail(ctx, ME_ERR_EMPTY, "cloud is empty");
  if (knn < 1 || knn > 64) return fail(ctx, ME_ERR_INVALID, "knn must be in 1..64");
  ME_TRY(build_grid(ctx, which, c.grid_valid && c.grid_solo ? c.solo_h : 0.0));
  if (c.slab) return fail(ctx, ME_ERR_INVALID, "me_estimate_normals needs the replicated layout (ME_LAYOUT_REPLICATED)");
  ME_TRY(ensure(ctx, (void **)&c.d_normal, &c.cap_normal, 3 * c.n, sizeof(double)));
  const int k = (int)std::min<long long>(knn, c.n);
  const size_t smem = (size_t)k * kKnnThreads * (sizeof(double) + sizeof(uint32_t));
  if (smem > 48 * 1024)
    ME_CUDA(ctx, cudaFuncSetAttribute(knn_normals_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double maxabs = 0;
  for (int a = 0; a < 3; ++a) maxabs = std::max(maxabs, std::max(std::fabs(c.bbox_min[a]), std::fabs(c.bbox_max[a])));
  const double slack = 1e-9 * c.lat.h + 4e-14 * maxabs;
  const int blocks = (int)std::min<long long>((c.n + kKnnThreads - 1) / kKnnThreads, (long long)ctx->sm_count * 32);
  knn_normals_kernel<<<blocks, kKnnThreads, smem, ctx->stream>>>(c.d_sorted, c.n, index_of(c), c.lat, coarse_of(c), k, slack, gicp, c.d_normal);
  ME_LAUNCH_CHECK(ctx);
  c.normal_valid = true;
  return ME_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// normal equations of point-to-plane (method 1) and generalized ICP (method 2) over the kept correspondences:
// v[0..20] upper triangle of JTJ (row-major), v[21..26] JTr, v[27] sum d2
// ---------------------------------------------------------------------------------------------------------------
struct NeAcc {
  unsigned long long n;
  double v[28];
};

__global__ void __launch_bounds__(kThreads)
icp_ne_accum_kernel(const P4 *__restrict__ Q, long long n, const int32_t *__restrict__ nn_idx, const double *__restrict__ nn_d2,
                    const double *__restrict__ gt_xyz, const double *__restrict__ gt_nrm, const double *__restrict__ est_nrm,
                    double r2, double eps, int method, NeAcc *__restrict__ acc) {
  double v[28];
#pragma unroll
  for (int k = 0; k < 28; ++k) v[k] = 0.0;
  unsigned long long cnt = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int32_t j = __ldg(nn_idx + i);
    if (j < 0) continue;
    const double d2 = __ldg(nn_d2 + i);
    if (!(d2 < r2)) continue;                              // SearchHybrid keeps d2 < R^2
    const P4 p = load_p4(Q + i);
    const double vs[3] = {p.x, p.y, p.z};
    const double d[3] = {p.x - __ldg(gt_xyz + 3ll * j), p.y - __ldg(gt_xyz + 3ll * j + 1), p.z - __ldg(gt_xyz + 3ll * j + 2)};
    const double nt[3] = {__ldg(gt_nrm + 3ll * j), __ldg(gt_nrm + 3ll * j + 1), __ldg(gt_nrm + 3ll * j + 2)};
    cnt++;
    v[27] += d2;
    if (method == ME_ICP_POINT_TO_PLANE) {
      double J[6];
      cross3d(vs, nt, J);
      J[3] = nt[0]; J[4] = nt[1]; J[5] = nt[2];
      const double r = dot3d(d, nt);
      int t = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = a; b < 6; ++b) v[t++] += J[a] * J[b];
        v[21 + a] += J[a] * r;
      }
    } else {
      const long long o = orig_of(p.idx);
      const double ns[3] = {__ldg(est_nrm + 3 * o), __ldg(est_nrm + 3 * o + 1), __ldg(est_nrm + 3 * o + 2)};
      // M = Cs + Ct = 2 I + (eps - 1)(ns ns^T + nt nt^T); B = M^-1 (symmetric, by cofactors)
      const double w = eps - 1.0;
      const double m00 = 2.0 + w * (ns[0] * ns[0] + nt[0] * nt[0]), m01 = w * (ns[0] * ns[1] + nt[0] * nt[1]), m02 = w * (ns[0] * ns[2] + nt[0] * nt[2]);
      const double m11 = 2.0 + w * (ns[1] * ns[1] + nt[1] * nt[1]), m12 = w * (ns[1] * ns[2] + nt[1] * nt[2]), m22 = 2.0 + w * (ns[2] * ns[2] + nt[2] * nt[2]);
      const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
      const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
      const double idet = 1.0 / (m00 * c00 + m01 * c01 + m02 * c02);
      const double B[9] = {c00 * idet, c01 * idet, c02 * idet, c01 * idet, c11 * idet, c12 * idet, c02 * idet, c12 * idet, c22 * idet};
      // A = [-skew(vs) | I] (3 x 6), BA = B A, JTJ = A^T B A, JTr = A^T B d
      const double A[18] = {0.0, vs[2], -vs[1], 1.0, 0.0, 0.0, -vs[2], 0.0, vs[0], 0.0, 1.0, 0.0, vs[1], -vs[0], 0.0, 0.0, 0.0, 1.0};
      double BA[18], g[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int col = 0; col < 6; ++col) BA[r * 6 + col] = B[r * 3] * A[col] + B[r * 3 + 1] * A[6 + col] + B[r * 3 + 2] * A[12 + col];
        g[r] = B[r * 3] * d[0] + B[r * 3 + 1] * d[1] + B[r * 3 + 2] * d[2];
      }
      int t = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = a; b < 6; ++b) v[t++] += A[a] * BA[b] + A[6 + a] * BA[6 + b] + A[12 + a] * BA[12 + b];
        v[21 + a] += A[a] * g[0] + A[6 + a] * g[1] + A[12 + a] * g[2];
      }
    }
  }
  __shared__ double sh[kThreads / 32][28];
  __shared__ unsigned long long shc[kThreads / 32];
#pragma unroll
  for (int k = 0; k < 28; ++k) v[k] = warp_sum(v[k]);
  cnt = (unsigned long long)warp_sum_ll((long long)cnt);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 28; ++k) sh[warp][k] = v[k];
    shc[warp] = cnt;
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    double s = 0;
    for (int w2 = 0; w2 < kThreads / 32; ++w2) s += sh[w2][threadIdx.x];
    if (s != 0.0) atomicAdd(&acc->v[threadIdx.x], s);
  } else if (threadIdx.x == 32) {
    unsigned long long s = 0;
    for (int w2 = 0; w2 < kThreads / 32; ++w2) s += shc[w2];
    if (s) atomicAdd(&acc->n, s);
  }
}

// symmetric 6x6 solve (Open3D: JTJ.ldlt().solve(-JTr)); Gaussian elimination with partial pivoting
static bool solve6_host(const double A_in[36], const double b_in[6], double x[6]) {
  double A[36], b[6];
  std::memcpy(A, A_in, sizeof(A)); std::memcpy(b, b_in, sizeof(b));
  for (int c = 0; c < 6; ++c) {
    int piv = c;
    for (int r = c + 1; r < 6; ++r) if (std::fabs(A[r * 6 + c]) > std::fabs(A[piv * 6 + c])) piv = r;
    if (A[piv * 6 + c] == 0.0) return false;
    if (piv != c) { for (int k = 0; k < 6; ++k) std::swap(A[c * 6 + k], A[piv * 6 + k]); std::swap(b[c], b[piv]); }
    for (int r = c + 1; r < 6; ++r) {
      const double f = A[r * 6 + c] / A[c * 6 + c];
      for (int k = c; k < 6; ++k) A[r * 6 + k] -= f * A[c * 6 + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = 5; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < 6; ++k) s -= A[r * 6 + k] * x[k];
    x[r] = s / A[r * 6 + r];
  }
  for (int k = 0; k < 6; ++k) if (!std::isfinite(x[k])) return false;
  return true;
}
// utility::TransformVector6dToMatrix4d: Rz(x2) Ry(x1) Rx(x0), translation x[3..5]
static void vec6_to_mat4_host(const double x[6], double T[16]) {
  const double ca = std::cos(x[0]), sa = std::sin(x[0]), cb = std::cos(x[1]), sb = std::sin(x[1]), cg = std::cos(x[2]), sg = std::sin(x[2]);
  const double Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca}, Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb}, Rz[9] = {cg, -sg, 0, sg, cg, 0, 0, 0, 1};
  double t[9], R[9];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[r * 3 + c] = Rz[r * 3] * Ry[c] + Rz[r * 3 + 1] * Ry[3 + c] + Rz[r * 3 + 2] * Ry[6 + c];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r * 3 + c] = t[r * 3] * Rx[c] + t[r * 3 + 1] * Rx[3 + c] + t[r * 3 + 2] * Rx[6 + c];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[r * 4 + c] = R[r * 3 + c]; T[r * 4 + 3] = x[3 + r]; }
  T[12] = T[13] = T[14] = 0; T[15] = 1;
}

int run_icp(me_ctx *ctx, int method, double max_dist, int max_iter, double rel_fitness, double rel_rmse, const double T_init[16],
            me_icp_result *out) {
  std::memset(out, 0, sizeof(*out));
  Cloud &E = ctx->cloud[ME_CLOUD_EST], &G = ctx->cloud[ME_CLOUD_GT];
  if (E.n <= 0 || G.n <= 0) return fail(ctx, ME_ERR_EMPTY, "both clouds must be set (map_eval.cpp:32-35)");
  if (method != ME_ICP_POINT_TO_POINT && method != ME_ICP_POINT_TO_PLANE && method != ME_ICP_GENERALIZED)
    return fail(ctx, ME_ERR_INVALID, "Invalid registration type specified (map_eval.cpp:1387)");
  if (ctx->world != 1) return fail(ctx, ME_ERR_INVALID, "me_icp needs world == 1 (the update needs all correspondences)");
  if (!E.owned) return fail(ctx, ME_ERR_INVALID, "me_icp needs a library-owned estimated cloud (use me_set_cloud)");
  if (!(max_dist > 0) || max_iter < 0) return fail(ctx, ME_ERR_INVALID, "bad ICP parameters");
  if (method == ME_ICP_POINT_TO_PLANE && !G.normal_valid)
    return fail(ctx, ME_ERR_INVALID, "TransformationEstimationPointToPlane requires pre-computed normal vectors for the target "
                                     "PointCloud (me_set_normals / me_estimate_normals on the ground-truth cloud)");
  ME_TRY(wait_upload(ctx, ME_CLOUD_EST));
  const double eps = 1e-3;      // TransformationEstimationForGeneralizedICP() default epsilon
  if (method == ME_ICP_GENERALIZED) {
    // InitializePointCloudForGeneralizedICP on both clouds (before the initial transform, as Open3D does)
    ME_TRY(build_both(ctx));
    ME_TRY(estimate_normals(ctx, ME_CLOUD_EST, 20, 1));
    ME_TRY(estimate_normals(ctx, ME_CLOUD_GT, 20, 1));
  }
  // the original cloud: RegistrationICP iterates on a copy, the caller's cloud is transformed once at the end (:1392)
  double *orig = nullptr;
  long long cap_orig = 0;
  ME_TRY(ensure(ctx, (void **)&orig, &cap_orig, 3 * E.n, sizeof(double)));
  int rc = ME_OK;
  if (cudaMemcpyAsync(orig, E.d_xyz, (size_t)E.n * 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
    rc = fail(ctx, ME_ERR_CUDA, "copy of the estimated cloud failed");

  double T[16];
  std::memcpy(T, T_init, sizeof(T));
  if (rc == ME_OK) rc = transform_cloud(ctx, ME_CLOUD_EST, T);      // rotates the (effective) normals along
  me_nn_params p;
  std::memset(&p, 0, sizeof(p));
  p.icp_max_distance = max_dist;
  p.cutoff_mode = ME_CUTOFF_DIST_LT_R;
  p.pairing = ME_PAIRING_GEOMETRIC;
  p.want_full_cd = 0;
  p.directions = 1;
  IcpAcc *d_acc = (IcpAcc *)((char *)ctx->d_scratch + 1024), *h_acc = (IcpAcc *)((char *)ctx->h_pinned + 1536);
  NeAcc *d_ne = (NeAcc *)((char *)ctx->d_scratch + 2048), *h_ne = (NeAcc *)((char *)ctx->h_pinned + 2048);
  IcpAcc res;
  NeAcc ne;
  std::memset(&res, 0, sizeof(res));
  std::memset(&ne, 0, sizeof(ne));
  unsigned long long n_corr = 0;
  double fitness = 0, rmse = 0, c[3] = {0, 0, 0};
  auto evaluate = [&]() -> int {
    me_nn_accum e2g;
    ME_TRY(run_nn(ctx, &p, &e2g, nullptr));                                  // lays the moved cloud out again, then sweeps
    const int blocks = (int)std::min<long long>((E.n + kThreads - 1) / kThreads, (long long)ctx->sm_count * 8);
    double err2 = 0;
    if (method == ME_ICP_POINT_TO_POINT) {
      for (int a = 0; a < 3; ++a) c[a] = 0.5 * (G.bbox_min[a] + G.bbox_max[a]);
      ME_CUDA(ctx, cudaMemsetAsync(d_acc, 0, sizeof(IcpAcc), ctx->stream));
      icp_accum_kernel<<<blocks, kThreads, 0, ctx->stream>>>(E.d_sorted, E.n, E.d_nn_idx, E.d_nn_d2, G.d_xyz, max_dist * max_dist,
                                                            c[0], c[1], c[2], d_acc);
      ME_LAUNCH_CHECK(ctx);
      ME_CUDA(ctx, cudaMemcpyAsync(h_acc, d_acc, sizeof(IcpAcc), cudaMemcpyDeviceToHost, ctx->stream));
      ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      res = *h_acc;
      n_corr = res.n; err2 = res.err2;
    } else {
      ME_CUDA(ctx, cudaMemsetAsync(d_ne, 0, sizeof(NeAcc), ctx->stream));
      icp_ne_accum_kernel<<<blocks, kThreads, 0, ctx->stream>>>(E.d_sorted, E.n, E.d_nn_idx, E.d_nn_d2, G.d_xyz, G.d_normal,
                                                               E.d_normal, max_dist * max_dist, eps, method, d_ne);
      ME_LAUNCH_CHECK(ctx);
      ME_CUDA(ctx, cudaMemcpyAsync(h_ne, d_ne, sizeof(NeAcc), cudaMemcpyDeviceToHost, ctx->stream));
      ME_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      ne = *h_ne;
      n_corr = ne.n; err2 = ne.v[27];
    }
    fitness = (double)n_corr / (double)E.n;
    rmse = n_corr > 0 ? std::sqrt(err2 / (double)n_corr) : 0.0;
    return ME_OK;
  };
  int it = 0, converged = 0;
  if (rc == ME_OK) rc = evaluate();
  for (; rc == ME_OK && it < max_iter; ++it) {
    // ComputeTransformation: the identity when there are no correspondences or the solve fails (Open3D)
    double upd[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    if (n_corr > 0) {
      if (method == ME_ICP_POINT_TO_POINT) umeyama_from_sums(res, c, upd);
      else {
        double JTJ[36], nb[6], x[6];
        int t = 0;
        for (int a = 0; a < 6; ++a)
          for (int b = a; b < 6; ++b) { JTJ[a * 6 + b] = ne.v[t]; JTJ[b * 6 + a] = ne.v[t]; ++t; }
        for (int a = 0; a < 6; ++a) nb[a] = -ne.v[21 + a];
        if (solve6_host(JTJ, nb, x)) vec6_to_mat4_host(x, upd);
      }
    }
    double Tn[16];
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
  // est := Transform(original, T) (map_eval.cpp:1392).  On an error the original cloud comes back untouched; either way
  // every derived structure of the estimated cloud is stale now.
  int rc2 = ME_OK;
  if (cudaMemcpyAsync(E.d_xyz, orig, (size_t)E.n * 3 * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
    rc2 = fail(ctx, ME_ERR_CUDA, "restoring the estimated cloud failed");
  invalidate_cloud(E);
  E.normal_valid = false;                         // the working normals of generalized ICP do not outlive the call
  if (rc == ME_OK && rc2 == ME_OK) rc2 = transform_cloud(ctx, ME_CLOUD_EST, T);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc2 == ME_OK) rc2 = fail(ctx, ME_ERR_CUDA, "synchronize after ICP failed");
  cudaFree(orig);
  if (method == ME_ICP_GENERALIZED) G.normal_valid = false;      // they hold the e1 substitution: not plain normals
  if (rc != ME_OK) return rc;
  if (rc2 != ME_OK) return rc2;
  std::memcpy(out->transformation, T, sizeof(T));
  out->fitness = fitness; out->inlier_rmse = rmse; out->n_corr = (int64_t)n_corr; out->iterations = it; out->converged = converged;
  return ME_OK;
}

}  // namespace me
