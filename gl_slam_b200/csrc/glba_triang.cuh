// glba_triang.cuh — two-view triangulation + cheirality / reprojection filter, one thread per match.
// Replaces slam_core::triangulate_and_filter_3d_points (GL-SLAM src/core/slam_core.cpp:173-256):
//   P0 = K [R1|t1], P1 = K [R2|t2] (:177-183); cv::triangulatePoints (:194) = DLT: the right singular vector of the
//   smallest singular value of the 4x4 system [x P(2,:) - P(0,:); y P(2,:) - P(1,:)] over both views;
//   reject |w| < 1e-9, depth <= 0 or > distance_threshold in either view, reprojection error > threshold in either view.
// The SVD is a one-sided (Hestenes) Jacobi on the 4x4 matrix itself (not on A'A: no squaring of the condition number),
// fully unrolled so everything stays in registers.
#pragma once
#include "glba_kernels.cuh"

namespace glba {

struct TriArgs {
  double P0[12], P1[12];     // K [R|t], row-major 3x4
  double T0[12], T1[12];     // [R|t], row-major 3x4
  Intr K;
  double dist_thr, reproj_thr;
};

template <int P, int Q>
__device__ __forceinline__ bool jacobi_pair(double (&A)[4][4], double (&V)[4][4]) {
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
  for (int r = 0; r < 4; ++r) { alpha += A[r][P] * A[r][P]; beta += A[r][Q] * A[r][Q]; gamma += A[r][P] * A[r][Q]; }
  if (fabs(gamma) <= 1e-17 * sqrt(alpha * beta) || gamma == 0.0) return false;
  const double zeta = (beta - alpha) / (2.0 * gamma);
  const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = rsqrt(1.0 + t * t), s = c * t;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const double ap = A[r][P], aq = A[r][Q];
    A[r][P] = c * ap - s * aq; A[r][Q] = s * ap + c * aq;
    const double vp = V[r][P], vq = V[r][Q];
    V[r][P] = c * vp - s * vq; V[r][Q] = s * vp + c * vq;
  }
  return true;
}

__global__ void k_triangulate(const int n, const TriArgs T, const double* __restrict__ p0, const double* __restrict__ p1,
                              double* __restrict__ X3 /* 3n: X/w */, uint8_t* __restrict__ keep) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x0 = p0[2 * i], y0 = p0[2 * i + 1], x1 = p1[2 * i], y1 = p1[2 * i + 1];
  double A[4][4], V[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    A[0][c] = x0 * T.P0[8 + c] - T.P0[c];
    A[1][c] = y0 * T.P0[8 + c] - T.P0[4 + c];
    A[2][c] = x1 * T.P1[8 + c] - T.P1[c];
    A[3][c] = y1 * T.P1[8 + c] - T.P1[4 + c];
#pragma unroll
    for (int r = 0; r < 4; ++r) V[r][c] = (r == c) ? 1.0 : 0.0;
  }
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool any = false;
    any |= jacobi_pair<0, 1>(A, V); any |= jacobi_pair<0, 2>(A, V); any |= jacobi_pair<0, 3>(A, V);
    any |= jacobi_pair<1, 2>(A, V); any |= jacobi_pair<1, 3>(A, V); any |= jacobi_pair<2, 3>(A, V);
    if (!any) break;
  }
  double best = 1e300;
  double h[4] = {0, 0, 0, 1};
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double nn = 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) nn += A[r][c] * A[r][c];
    if (nn < best) {
      best = nn;
#pragma unroll
      for (int r = 0; r < 4; ++r) h[r] = V[r][c];
    }
  }
  const double w = h[3];
  bool ok = !(fabs(w) < 1e-9);                                         // :213
  const double iw = ok ? 1.0 / w : 0.0;
  double c0[3], c1[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c0[r] = (T.T0[4 * r] * h[0] + T.T0[4 * r + 1] * h[1] + T.T0[4 * r + 2] * h[2] + T.T0[4 * r + 3] * h[3]) * iw;
    c1[r] = (T.T1[4 * r] * h[0] + T.T1[4 * r + 1] * h[1] + T.T1[4 * r + 2] * h[2] + T.T1[4 * r + 3] * h[3]) * iw;
  }
  ok = ok && !(c0[2] <= 0.0 || c0[2] > T.dist_thr) && !(c1[2] <= 0.0 || c1[2] > T.dist_thr);   // :216-221
  if (ok) {
    const double u0 = T.K.fx * c0[0] / c0[2] + T.K.cx, v0 = T.K.fy * c0[1] / c0[2] + T.K.cy;
    const double u1 = T.K.fx * c1[0] / c1[2] + T.K.cx, v1 = T.K.fy * c1[1] / c1[2] + T.K.cy;
    if (sqrt((u0 - x0) * (u0 - x0) + (v0 - y0) * (v0 - y0)) > T.reproj_thr) ok = false;         // :230-231
    if (sqrt((u1 - x1) * (u1 - x1) + (v1 - y1) * (v1 - y1)) > T.reproj_thr) ok = false;         // :240-241
  }
  X3[3 * i] = h[0] * iw; X3[3 * i + 1] = h[1] * iw; X3[3 * i + 2] = h[2] * iw;
  keep[i] = ok ? 1 : 0;
}

}  // namespace glba
