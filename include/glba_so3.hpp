// glba_so3.hpp — ProjectToSO3 / ComputeDeltaPose_SO3 of GL-SLAM (src/core/slam_core.cpp:885-912), one restatement shared
// by the host adaptor (glba_slam.hpp) and the device kernel behind glba_map_propagate (glba_map.cuh).
//
// The reference takes cv::SVD of the 3x3 input, returns U V', and if det(U V') < 0 flips the LAST column of U (the singular
// vector of the smallest singular value; cv::SVD orders them descending) and returns that product instead (:890-895).
// Here the SVD is obtained from a cyclic Jacobi eigen-decomposition of A'A = V S^2 V' (3x3, converges in a few sweeps; the
// inputs are rotations up to numerical drift, so S ~ I and U = A V S^-1 is well conditioned), eigenpairs ordered by
// descending singular value as OpenCV orders them.  Plain C++, no dependencies; row-major 3x3 in double[9].
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GLBA_HD __host__ __device__ inline
#else
#define GLBA_HD inline
#endif

namespace glba_so3 {

GLBA_HD double det3(const double* A) {
  return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
GLBA_HD void mul3(const double* A, const double* B, double* C) {          // C = A B
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
GLBA_HD void mul3_bt(const double* A, const double* B, double* C) {       // C = A B'
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c * 3] + A[r * 3 + 1] * B[c * 3 + 1] + A[r * 3 + 2] * B[c * 3 + 2];
}

// Eigen-decomposition of the symmetric 3x3 M (destroyed): M = V diag(e) V', columns of V = eigenvectors.
GLBA_HD void jacobi_eig3(double* M, double* V, double* e) {
  for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    const double off = fabs(M[1]) + fabs(M[2]) + fabs(M[5]);
    if (off < 1e-300) break;
    if (off <= 1e-17 * (fabs(M[0]) + fabs(M[4]) + fabs(M[8]))) break;
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      const double apq = M[p * 3 + q];
      if (apq == 0.0) continue;
      const double theta = (M[q * 3 + q] - M[p * 3 + p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; ++k) {          // M <- M J  (columns p, q)
        const double mkp = M[k * 3 + p], mkq = M[k * 3 + q];
        M[k * 3 + p] = c * mkp - s * mkq; M[k * 3 + q] = s * mkp + c * mkq;
      }
      for (int k = 0; k < 3; ++k) {          // M <- J' M  (rows p, q)
        const double mpk = M[p * 3 + k], mqk = M[q * 3 + k];
        M[p * 3 + k] = c * mpk - s * mqk; M[q * 3 + k] = s * mpk + c * mqk;
      }
      for (int k = 0; k < 3; ++k) {
        const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
        V[k * 3 + p] = c * vkp - s * vkq; V[k * 3 + q] = s * vkp + c * vkq;
      }
    }
  }
  e[0] = M[0]; e[1] = M[4]; e[2] = M[8];
}

// ProjectToSO3 (slam_core.cpp:885-897).
GLBA_HD void project_to_so3(const double* A, double* R) {
  double M[9], V[9], e[3];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M[r * 3 + c] = A[r] * A[c] + A[3 + r] * A[3 + c] + A[6 + r] * A[6 + c];   // A'A
  jacobi_eig3(M, V, e);
  int ord[3] = {0, 1, 2};                                   // descending singular values, as cv::SVD returns them
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2 - a; ++b)
      if (e[ord[b]] < e[ord[b + 1]]) { const int t = ord[b]; ord[b] = ord[b + 1]; ord[b + 1] = t; }
  double U[9], Vs[9];
  for (int k = 0; k < 3; ++k) {
    const int col = ord[k];
    const double sv = sqrt(e[col] > 0.0 ? e[col] : 0.0);
    double u[3];
    for (int r = 0; r < 3; ++r) u[r] = A[r * 3] * V[col] + A[r * 3 + 1] * V[3 + col] + A[r * 3 + 2] * V[6 + col];
    const double inv = sv > 0.0 ? 1.0 / sv : 0.0;
    for (int r = 0; r < 3; ++r) { U[r * 3 + k] = u[r] * inv; Vs[r * 3 + k] = V[r * 3 + col]; }
  }
  mul3_bt(U, Vs, R);                                        // U V'
  if (det3(R) < 0.0) {                                      // "Flip last column of U and recompute" (:892-895)
    for (int r = 0; r < 3; ++r) U[r * 3 + 2] = -U[r * 3 + 2];
    mul3_bt(U, Vs, R);
  }
}

// ComputeDeltaPose_SO3 (slam_core.cpp:899-912): dR = Proj(Proj(Ra) Proj(Rb)'), dt = ta - dR tb.
GLBA_HD void compute_delta_pose_so3(const double* Rb_in, const double* tb, const double* Ra_in, const double* ta, double* dR, double* dt) {
  double Rb[9], Ra[9], P[9];
  project_to_so3(Rb_in, Rb);
  project_to_so3(Ra_in, Ra);
  mul3_bt(Ra, Rb, P);
  project_to_so3(P, dR);
  for (int r = 0; r < 3; ++r) dt[r] = ta[r] - (dR[r * 3] * tb[0] + dR[r * 3 + 1] * tb[1] + dR[r * 3 + 2] * tb[2]);
}

}  // namespace glba_so3
