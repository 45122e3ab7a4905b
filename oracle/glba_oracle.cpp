/*
 * glba_oracle.cpp — CPU restatement of GL-SLAM's bundle-adjustment path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (gl_slam_b200/csrc) never links, loads or calls it.
 *
 * PARITY UNPINNED: the arithmetic of the reference path lives in Ceres Solver, which the reference
 * does not vendor (CMakeLists.txt:14-15 find_package(Ceres) under a git-ignored third_party/,
 * .gitignore:13; version unpinned, `ceres::CUDA` at src/core/slam_core.cpp:1120 implies >= 2.1).
 * Neither Ceres nor Eigen exists in this image and the reference ships no tests or golden vectors,
 * so this file restates Ceres' *published* trust-region Levenberg-Marquardt algorithm and is
 * anchored on the reference's own call sites:
 *   residual functor          src/core/slam_core.cpp:699-733   (ReprojectionError::operator())
 *   autodiff 2x6 / 2x3        src/core/slam_core.cpp:735-738   (AutoDiffCostFunction<...,2,6,3>)
 *   robust loss               src/core/slam_core.cpp:814       (CauchyLoss(1.0); Huber per :1115 comment)
 *   fixed cameras             src/core/slam_core.cpp:831-833
 *   solver options            src/core/slam_core.cpp:842-847   (SPARSE_SCHUR, 30 iterations, 8 threads)
 *   pose-only variant         src/core/slam_core.cpp:1043-1140
 * It is cross-checked against an independent numpy/complex-step implementation (oracle/py_oracle.py) and against
 * OpenCV — the reference's own dependency, importable here as cv2: Rodrigues convention, analytic projectPoints
 * Jacobians, solvePnPRefineLM optimum; see tests/test_oracle.py.  None of that replaces a run of libceres: the LM
 * trajectory itself (accept/reject sequence, radii) remains unpinned.
 *
 * Ceres semantics encoded here (Ceres 2.x, documented behaviour):
 *   cost = 1/2 sum rho(|r|^2); corrector with rho'' <= 0: r~ = sqrt(rho') r, J~ = sqrt(rho') J;
 *   Jacobi column scaling 1/(1+|col|) fixed at iteration 0; LM diagonal clamp [1e-6,1e32],
 *   D = sqrt(diag/radius); exact Schur elimination of the point blocks; step = -y;
 *   model_cost_change = -(J step).(r~ + J step/2); parameter tolerance, function tolerance and
 *   step acceptance tested in that order; radius update radius/max(1/3, 1-(2q-1)^3) on accept,
 *   radius/decrease_factor (factor doubling) on reject; rejected steps count as iterations.
 */
#include "../include/glba.h"

#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// Forward-mode dual numbers with N partials: the arithmetic an AutoDiffCostFunction performs.
// ---------------------------------------------------------------------------------------------
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  explicit Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) {
  Jet<N> h; h.a = s - f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) {
  Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator/(double s, const Jet<N>& g) {
  Jet<N> h; const double gi = 1.0 / g.a; h.a = s * gi; const double m = -s * gi * gi;
  for (int i = 0; i < N; ++i) h.v[i] = m * g.v[i]; return h; }
template <int N> inline Jet<N> jsqrt(const Jet<N>& f) {
  Jet<N> h; const double t = std::sqrt(f.a); h.a = t; const double m = 1.0 / (2.0 * t);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * m; return h; }
template <int N> inline Jet<N> jcos(const Jet<N>& f) {
  Jet<N> h; h.a = std::cos(f.a); const double m = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
template <int N> inline Jet<N> jsin(const Jet<N>& f) {
  Jet<N> h; h.a = std::sin(f.a); const double m = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = m * f.v[i]; return h; }
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jcos(double x) { return std::cos(x); }
inline double jsin(double x) { return std::sin(x); }
inline double scalar_of(double x) { return x; }
template <int N> inline double scalar_of(const Jet<N>& x) { return x.a; }
template <typename T> inline T make_const(double s);
template <> inline double make_const<double>(double s) { return s; }
template <> inline Jet<9> make_const<Jet<9>>(double s) { return Jet<9>(s); }
template <> inline Jet<6> make_const<Jet<6>>(double s) { return Jet<6>(s); }

// ceres::AngleAxisRotatePoint as documented in ceres/rotation.h (called at slam_core.cpp:713):
// Rodrigues' formula for theta^2 > epsilon, first-order Taylor expansion otherwise.
template <typename T>
inline void angle_axis_rotate_point(const T aa[3], const T pt[3], T out[3]) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (scalar_of(theta2) > std::numeric_limits<double>::epsilon()) {
    const T theta = jsqrt(theta2);
    const T costheta = jcos(theta);
    const T sintheta = jsin(theta);
    const T theta_inverse = 1.0 / theta;
    const T w[3] = {aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const T w_cross_pt[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2],
                             w[0] * pt[1] - w[1] * pt[0]};
    const T tmp = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (1.0 - costheta);
    for (int i = 0; i < 3; ++i) out[i] = pt[i] * costheta + w_cross_pt[i] * sintheta + w[i] * tmp;
  } else {
    const T w_cross_pt[3] = {aa[1] * pt[2] - aa[2] * pt[1], aa[2] * pt[0] - aa[0] * pt[2],
                             aa[0] * pt[1] - aa[1] * pt[0]};
    for (int i = 0; i < 3; ++i) out[i] = pt[i] + w_cross_pt[i];
  }
}

// ReprojectionError::operator() (slam_core.cpp:699-733): camera = [angle-axis of R_wc, centre],
// q = X - c, p = R(-w) q, pinhole projection, residual = prediction - observation.
template <typename T>
inline void reprojection_residual(const T cam[6], const T pt[3], double fx, double fy, double cx,
                                  double cy, double u, double v, T res[2]) {
  T p_trans[3] = {pt[0] - cam[3], pt[1] - cam[4], pt[2] - cam[5]};
  T minus_cam[3] = {-cam[0], -cam[1], -cam[2]};
  T p[3];
  angle_axis_rotate_point(minus_cam, p_trans, p);
  const T xp = p[0] / p[2];
  const T yp = p[1] / p[2];
  const T predicted_x = fx * xp + cx;
  const T predicted_y = fy * yp + cy;
  res[0] = predicted_x - u;
  res[1] = predicted_y - v;
}

// ceres::LossFunction::Evaluate for TrivialLoss / HuberLoss(a) / CauchyLoss(a) (slam_core.cpp:814, :1115).
inline void loss_eval(int loss, double a, double s, double rho[3]) {
  if (loss == GLBA_LOSS_CAUCHY) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c;
    const double inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(std::numeric_limits<double>::min(), inv);
    rho[2] = -c * (inv * inv);
  } else if (loss == GLBA_LOSS_HUBER) {
    const double b = a * a;
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
  } else {
    rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
  }
}

struct View {
  int n_cam, n_pt;
  long n_obs;
  const int32_t* obs_cam;
  const int32_t* obs_pt;
  const double* obs_u;
  const double* obs_v;
  double fx, fy, cx, cy;
  const double* pt_info;           // optional per-point information weight (glba_problem::pt_info), nullptr = 1
  std::vector<uint8_t> cam_free;   // not fixed and observed
  std::vector<uint8_t> pt_free;    // not fixed and observed
  std::vector<uint8_t> pt_seen;
  std::vector<int> cam_slot;       // index among free cameras or -1
  int n_free_cam;
  // tracks: observations grouped by point (stable in the caller's order)
  std::vector<long> trk_start;     // n_pt+1
  std::vector<long> trk_obs;       // n_obs
};

int build_view(const glba_problem* p, View& V) {
  if (!p || p->n_cam < 0 || p->n_pt < 0 || p->n_obs < 0) return GLBA_E_INVALID_ARG;
  if (p->n_obs > 0 && (!p->obs_cam || !p->obs_pt || !p->obs_u || !p->obs_v)) return GLBA_E_INVALID_ARG;
  if ((p->n_cam > 0 && !p->cam) || (p->n_pt > 0 && !p->pt)) return GLBA_E_INVALID_ARG;
  V.n_cam = p->n_cam; V.n_pt = p->n_pt; V.n_obs = (long)p->n_obs;
  V.obs_cam = p->obs_cam; V.obs_pt = p->obs_pt; V.obs_u = p->obs_u; V.obs_v = p->obs_v;
  V.fx = p->fx; V.fy = p->fy; V.cx = p->cx; V.cy = p->cy;
  V.pt_info = p->pt_info;
  std::vector<uint8_t> cam_seen(V.n_cam, 0);
  V.pt_seen.assign(V.n_pt, 0);
  V.trk_start.assign(V.n_pt + 1, 0);
  for (long k = 0; k < V.n_obs; ++k) {
    const int c = V.obs_cam[k], j = V.obs_pt[k];
    if (c < 0 || c >= V.n_cam || j < 0 || j >= V.n_pt) return GLBA_E_INVALID_ARG;
    cam_seen[c] = 1; V.pt_seen[j] = 1; V.trk_start[j + 1]++;
  }
  for (int j = 0; j < V.n_pt; ++j) V.trk_start[j + 1] += V.trk_start[j];
  V.trk_obs.resize(V.n_obs);
  { std::vector<long> fill(V.trk_start.begin(), V.trk_start.end() - 1);
    for (long k = 0; k < V.n_obs; ++k) V.trk_obs[fill[V.obs_pt[k]]++] = k; }
  V.cam_free.assign(V.n_cam, 0); V.cam_slot.assign(V.n_cam, -1); V.n_free_cam = 0;
  for (int i = 0; i < V.n_cam; ++i) {
    const bool fixed = p->cam_fixed && p->cam_fixed[i];
    if (!fixed && cam_seen[i]) { V.cam_free[i] = 1; V.cam_slot[i] = V.n_free_cam++; }
  }
  V.pt_free.assign(V.n_pt, 0);
  for (int j = 0; j < V.n_pt; ++j) {
    const bool fixed = p->pt_fixed && p->pt_fixed[j];
    V.pt_free[j] = (!fixed && V.pt_seen[j]) ? 1 : 0;
  }
  return GLBA_OK;
}

struct Lin {              // one evaluation: loss-corrected residuals and Jacobian blocks per observation
  std::vector<double> r;  // 2*n_obs
  std::vector<double> jc; // 12*n_obs (row-major 2x6)
  std::vector<double> jp; // 6*n_obs  (row-major 2x3)
};

// Evaluate cost (and optionally the corrected residuals / Jacobians).  Returns false if any
// residual or Jacobian entry is non-finite (Ceres: evaluation failure).
bool evaluate(const View& V, const double* cam, const double* pt, int loss, double loss_a,
              double* cost_out, Lin* lin) {
  const long n = V.n_obs;
  if (lin) { lin->r.resize(2 * n); lin->jc.resize(12 * n); lin->jp.resize(6 * n); }
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<double> part(nthreads, 0.0);
  std::vector<int> bad(nthreads, 0);
#pragma omp parallel num_threads(nthreads)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double acc = 0.0; int isbad = 0;
#pragma omp for schedule(static)
    for (long k = 0; k < n; ++k) {
      const double* c = cam + 6 * (long)V.obs_cam[k];
      const double* x = pt + 3 * (long)V.obs_pt[k];
      double r[2], s;
      if (lin) {
        typedef Jet<9> J;
        J jc[6], jx[3], jr[2];
        for (int i = 0; i < 6; ++i) jc[i] = J(c[i], i);
        for (int i = 0; i < 3; ++i) jx[i] = J(x[i], 6 + i);
        reprojection_residual<J>(jc, jx, V.fx, V.fy, V.cx, V.cy, V.obs_u[k], V.obs_v[k], jr);
        r[0] = jr[0].a; r[1] = jr[1].a;
        const double info = V.pt_info ? std::max(V.pt_info[V.obs_pt[k]], 0.0) : 1.0;     // r' Omega r, Omega = info * I
        s = info * (r[0] * r[0] + r[1] * r[1]);
        double rho[3]; loss_eval(loss, loss_a, s, rho);
        acc += 0.5 * rho[0];
        const double sq = std::sqrt(info) * std::sqrt(rho[1]);   // corrector, rho'' <= 0 branch for all three losses
        double* Jc = &lin->jc[12 * k]; double* Jp = &lin->jp[6 * k];
        for (int row = 0; row < 2; ++row) {
          for (int i = 0; i < 6; ++i) { Jc[row * 6 + i] = sq * jr[row].v[i]; if (!std::isfinite(Jc[row * 6 + i])) isbad = 1; }
          for (int i = 0; i < 3; ++i) { Jp[row * 3 + i] = sq * jr[row].v[6 + i]; if (!std::isfinite(Jp[row * 3 + i])) isbad = 1; }
        }
        lin->r[2 * k] = sq * r[0]; lin->r[2 * k + 1] = sq * r[1];
      } else {
        reprojection_residual<double>(c, x, V.fx, V.fy, V.cx, V.cy, V.obs_u[k], V.obs_v[k], r);
        s = (V.pt_info ? std::max(V.pt_info[V.obs_pt[k]], 0.0) : 1.0) * (r[0] * r[0] + r[1] * r[1]);
        double rho[3]; loss_eval(loss, loss_a, s, rho);
        acc += 0.5 * rho[0];
      }
      if (!std::isfinite(r[0]) || !std::isfinite(r[1])) isbad = 1;
    }
    part[tid] = acc; bad[tid] = isbad;
  }
  double cost = 0.0; int anybad = 0;
  for (int t = 0; t < nthreads; ++t) { cost += part[t]; anybad |= bad[t]; }
  *cost_out = cost;
  return !anybad && std::isfinite(cost);
}

// 3x3 symmetric positive definite inverse via Cholesky; returns false if not PD.
inline bool inv3_spd(const double C[9], double Ci[9]) {
  const double l00 = std::sqrt(C[0]);
  if (!(C[0] > 0.0)) return false;
  const double l10 = C[3] / l00, l20 = C[6] / l00;
  const double d1 = C[4] - l10 * l10; if (!(d1 > 0.0)) return false;
  const double l11 = std::sqrt(d1);
  const double l21 = (C[7] - l20 * l10) / l11;
  const double d2 = C[8] - l20 * l20 - l21 * l21; if (!(d2 > 0.0)) return false;
  const double l22 = std::sqrt(d2);
  // inverse of L
  const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
  const double i10 = -l10 * i00 * i11;
  const double i21 = -l21 * i11 * i22;
  const double i20 = -(l20 * i00 + l21 * i10) * i22;
  // Ci = Li' Li
  Ci[0] = i00 * i00 + i10 * i10 + i20 * i20;
  Ci[1] = Ci[3] = i10 * i11 + i20 * i21;
  Ci[2] = Ci[6] = i20 * i22;
  Ci[4] = i11 * i11 + i21 * i21;
  Ci[5] = Ci[7] = i21 * i22;
  Ci[8] = i22 * i22;
  return true;
}

// In-place dense Cholesky (lower) of an n x n row-major SPD matrix and solve; false if not PD.
bool cholesky_solve(std::vector<double>& A, int n, std::vector<double>& b) {
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    const double ljj = std::sqrt(d);
    A[(size_t)j * n + j] = ljj;
    const double inv = 1.0 / ljj;
#pragma omp parallel for schedule(static) if (n - j > 256)
    for (int i = j + 1; i < n; ++i) {
      double s = A[(size_t)i * n + j];
      const double* ai = &A[(size_t)i * n];
      const double* aj = &A[(size_t)j * n];
      for (int k = 0; k < j; ++k) s -= ai[k] * aj[k];
      A[(size_t)i * n + j] = s * inv;
    }
  }
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= A[(size_t)i * n + k] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= A[(size_t)k * n + i] * b[k];
    b[i] = s / A[(size_t)i * n + i];
  }
  return true;
}

// 6x6 SPD inverse by Cholesky (for the block-Jacobi preconditioner).
bool inv6_spd(const double* A, double* Ai) {
  double L[36] = {0};
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
    if (!(d > 0.0)) return false;
    L[j * 6 + j] = std::sqrt(d);
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = s / L[j * 6 + j];
    }
  }
  for (int c = 0; c < 6; ++c) {
    double y[6];
    for (int i = 0; i < 6; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int k = 0; k < i; ++k) s -= L[i * 6 + k] * y[k];
      y[i] = s / L[i * 6 + i];
    }
    for (int i = 5; i >= 0; --i) {
      double s = y[i];
      for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * y[k];
      y[i] = s / L[i * 6 + i];
    }
    for (int i = 0; i < 6; ++i) Ai[i * 6 + c] = y[i];
  }
  return true;
}

// The (scaled) linear system of one LM iteration, Schur-eliminated on the point blocks.
struct Schur {
  int nfc = 0;                       // free cameras
  bool dense = true;
  std::vector<double> S;             // dense (6nfc)^2, or empty in PCG mode
  std::vector<double> Sdiag;         // 36*nfc diagonal blocks (always)
  std::vector<double> rhs;           // 6*nfc
  std::vector<double> Cinv;          // 9*n_pt
  std::vector<double> gp;            // 3*n_pt  (J_p' r~, scaled)
  std::vector<double> Bdiag;         // 36*nfc: J_c'J_c + D_c^2 (PCG mode: needed for the product)
};

// Js: Jacobian with columns already scaled.  Dc (6 per camera, indexed by camera), Dp (3 per point).
bool build_schur(const View& V, const Lin& L, const std::vector<double>& Dc, const std::vector<double>& Dp,
                 bool dense, Schur& out) {
  const int nfc = V.n_free_cam; const int n = 6 * nfc;
  out.nfc = nfc; out.dense = dense;
  out.rhs.assign(n, 0.0); out.Sdiag.assign((size_t)36 * nfc, 0.0);
  out.Cinv.assign((size_t)9 * V.n_pt, 0.0); out.gp.assign((size_t)3 * V.n_pt, 0.0);
  if (dense) out.S.assign((size_t)n * n, 0.0); else { out.S.clear(); out.Bdiag.assign((size_t)36 * nfc, 0.0); }
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  // per-thread accumulators, merged in thread order (deterministic for a fixed thread count)
  std::vector<std::vector<double>> tS(nthreads), tD(nthreads), tR(nthreads), tB(nthreads);
  int notpd = 0;
#pragma omp parallel num_threads(nthreads) reduction(| : notpd)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    std::vector<double>& myS = tS[tid]; std::vector<double>& myD = tD[tid];
    std::vector<double>& myR = tR[tid]; std::vector<double>& myB = tB[tid];
    if (dense) myS.assign((size_t)n * n, 0.0);
    myD.assign((size_t)36 * nfc, 0.0); myR.assign(n, 0.0);
    if (!dense) myB.assign((size_t)36 * nfc, 0.0);
    std::vector<double> W;  // 18 per observation in the track
#pragma omp for schedule(static)
    for (int j = 0; j < V.n_pt; ++j) {
      const long b = V.trk_start[j], e = V.trk_start[j + 1];
      if (b == e) continue;
      const bool pfree = V.pt_free[j];
      double C[9] = {0}, g[3] = {0};
      // camera diagonal blocks B_i += Jc'Jc, rhs_i += Jc' r
      for (long t = b; t < e; ++t) {
        const long k = V.trk_obs[t];
        const int slot = V.cam_slot[V.obs_cam[k]];
        const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k]; const double* r = &L.r[2 * k];
        if (slot >= 0) {
          double* Bd = dense ? &myS[0] : &myB[(size_t)36 * slot];
          for (int a = 0; a < 6; ++a) {
            for (int c = 0; c < 6; ++c) {
              const double v = Jc[a] * Jc[c] + Jc[6 + a] * Jc[6 + c];
              if (dense) Bd[(size_t)(6 * slot + a) * n + 6 * slot + c] += v; else Bd[a * 6 + c] += v;
              myD[(size_t)36 * slot + a * 6 + c] += v;
            }
            myR[6 * slot + a] += Jc[a] * r[0] + Jc[6 + a] * r[1];
          }
        }
        if (pfree) {
          for (int a = 0; a < 3; ++a) {
            for (int c = 0; c < 3; ++c) C[a * 3 + c] += Jp[a] * Jp[c] + Jp[3 + a] * Jp[3 + c];
            g[a] += Jp[a] * r[0] + Jp[3 + a] * r[1];
          }
        }
      }
      if (!pfree) continue;
      for (int a = 0; a < 3; ++a) C[a * 3 + a] += Dp[3 * j + a] * Dp[3 * j + a];
      double Ci[9];
      if (!inv3_spd(C, Ci)) { notpd |= 1; continue; }
      std::memcpy(&out.Cinv[(size_t)9 * j], Ci, sizeof(Ci));
      out.gp[3 * j] = g[0]; out.gp[3 * j + 1] = g[1]; out.gp[3 * j + 2] = g[2];
      const double u0[3] = {Ci[0] * g[0] + Ci[1] * g[1] + Ci[2] * g[2], Ci[3] * g[0] + Ci[4] * g[1] + Ci[5] * g[2],
                            Ci[6] * g[0] + Ci[7] * g[1] + Ci[8] * g[2]};
      const int len = (int)(e - b);
      W.resize((size_t)18 * len);
      for (int t = 0; t < len; ++t) {       // W_t = Jc' Jp (6x3)
        const long k = V.trk_obs[b + t];
        const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k];
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c)
          W[(size_t)18 * t + a * 3 + c] = Jc[a] * Jp[c] + Jc[6 + a] * Jp[3 + c];
      }
      for (int t = 0; t < len; ++t) {
        const int si = V.cam_slot[V.obs_cam[V.trk_obs[b + t]]];
        if (si < 0) continue;
        const double* Wt = &W[(size_t)18 * t];
        double WC[18];                       // W_t Cinv (6x3)
        for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c)
          WC[a * 3 + c] = Wt[a * 3] * Ci[c] + Wt[a * 3 + 1] * Ci[3 + c] + Wt[a * 3 + 2] * Ci[6 + c];
        for (int a = 0; a < 6; ++a) myR[6 * si + a] -= Wt[a * 3] * u0[0] + Wt[a * 3 + 1] * u0[1] + Wt[a * 3 + 2] * u0[2];
        for (int t2 = 0; t2 < len; ++t2) {
          const int sk = V.cam_slot[V.obs_cam[V.trk_obs[b + t2]]];
          if (sk < 0) continue;
          if (!dense && sk != si) continue;
          const double* Wk = &W[(size_t)18 * t2];
          for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c) {
            const double v = WC[a * 3] * Wk[c * 3] + WC[a * 3 + 1] * Wk[c * 3 + 1] + WC[a * 3 + 2] * Wk[c * 3 + 2];
            if (dense) myS[(size_t)(6 * si + a) * n + 6 * sk + c] -= v;
            if (sk == si) myD[(size_t)36 * si + a * 6 + c] -= v;
          }
        }
      }
    }
  }
  if (notpd) return false;
  for (int t = 0; t < nthreads; ++t) {
    if (dense) for (size_t i = 0; i < (size_t)n * n; ++i) out.S[i] += tS[t][i];
    else for (size_t i = 0; i < (size_t)36 * nfc; ++i) out.Bdiag[i] += tB[t][i];
    for (size_t i = 0; i < (size_t)36 * nfc; ++i) out.Sdiag[i] += tD[t][i];
    for (int i = 0; i < n; ++i) out.rhs[i] += tR[t][i];
  }
  for (int i = 0; i < V.n_cam; ++i) {
    const int slot = V.cam_slot[i]; if (slot < 0) continue;
    for (int a = 0; a < 6; ++a) {
      const double d2 = Dc[6 * i + a] * Dc[6 * i + a];
      if (dense) out.S[(size_t)(6 * slot + a) * n + 6 * slot + a] += d2; else out.Bdiag[(size_t)36 * slot + a * 7] += d2;
      out.Sdiag[(size_t)36 * slot + a * 7] += d2;
    }
  }
  return true;
}

// y = S x with S implicit: (B + Dc^2) x - sum_j W_j Cinv_j W_j' x   (PCG mode)
void schur_apply(const View& V, const Lin& L, const Schur& sc, const std::vector<double>& x, std::vector<double>& y) {
  const int nfc = sc.nfc; const int n = 6 * nfc;
  y.assign(n, 0.0);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<std::vector<double>> ty(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    std::vector<double>& my = ty[tid]; my.assign(n, 0.0);
#pragma omp for schedule(static)
    for (int j = 0; j < V.n_pt; ++j) {
      if (!V.pt_free[j]) continue;
      const long b = V.trk_start[j], e = V.trk_start[j + 1];
      double t3[3] = {0, 0, 0};
      for (long t = b; t < e; ++t) {
        const long k = V.trk_obs[t]; const int s = V.cam_slot[V.obs_cam[k]]; if (s < 0) continue;
        const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k];
        double a0 = 0, a1 = 0;
        for (int a = 0; a < 6; ++a) { a0 += Jc[a] * x[6 * s + a]; a1 += Jc[6 + a] * x[6 * s + a]; }
        for (int c = 0; c < 3; ++c) t3[c] += Jp[c] * a0 + Jp[3 + c] * a1;
      }
      const double* Ci = &sc.Cinv[(size_t)9 * j];
      const double u[3] = {Ci[0] * t3[0] + Ci[1] * t3[1] + Ci[2] * t3[2], Ci[3] * t3[0] + Ci[4] * t3[1] + Ci[5] * t3[2],
                           Ci[6] * t3[0] + Ci[7] * t3[1] + Ci[8] * t3[2]};
      for (long t = b; t < e; ++t) {
        const long k = V.trk_obs[t]; const int s = V.cam_slot[V.obs_cam[k]]; if (s < 0) continue;
        const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k];
        const double b0 = Jp[0] * u[0] + Jp[1] * u[1] + Jp[2] * u[2];
        const double b1 = Jp[3] * u[0] + Jp[4] * u[1] + Jp[5] * u[2];
        for (int a = 0; a < 6; ++a) my[6 * s + a] -= Jc[a] * b0 + Jc[6 + a] * b1;
      }
    }
  }
  for (int t = 0; t < nthreads; ++t) for (int i = 0; i < n; ++i) y[i] += ty[t][i];
  for (int s = 0; s < nfc; ++s)
    for (int a = 0; a < 6; ++a) {
      double v = 0; for (int c = 0; c < 6; ++c) v += sc.Bdiag[(size_t)36 * s + a * 6 + c] * x[6 * s + c];
      y[6 * s + a] += v;
    }
}

int pcg_solve(const View& V, const Lin& L, const Schur& sc, double tol, int max_it, std::vector<double>& x) {
  const int nfc = sc.nfc; const int n = 6 * nfc;
  std::vector<double> Minv((size_t)36 * nfc);
  for (int s = 0; s < nfc; ++s) if (!inv6_spd(&sc.Sdiag[(size_t)36 * s], &Minv[(size_t)36 * s])) return -1;
  auto precond = [&](const std::vector<double>& r, std::vector<double>& z) {
    for (int s = 0; s < nfc; ++s) for (int a = 0; a < 6; ++a) {
      double v = 0; for (int c = 0; c < 6; ++c) v += Minv[(size_t)36 * s + a * 6 + c] * r[6 * s + c];
      z[6 * s + a] = v; }
  };
  x.assign(n, 0.0);
  std::vector<double> r = sc.rhs, z(n), p(n), q(n);
  precond(r, z); p = z;
  double rz = 0; for (int i = 0; i < n; ++i) rz += r[i] * z[i];
  const double rz0 = rz;
  if (!(rz0 > 0.0)) return 0;
  int it = 0;
  for (; it < max_it; ++it) {
    if (std::sqrt(rz) <= tol * std::sqrt(rz0)) break;
    schur_apply(V, L, sc, p, q);
    double pq = 0; for (int i = 0; i < n; ++i) pq += p[i] * q[i];
    if (!(pq > 0.0)) break;
    const double alpha = rz / pq;
    for (int i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * q[i]; }
    precond(r, z);
    double rz1 = 0; for (int i = 0; i < n; ++i) rz1 += r[i] * z[i];
    const double beta = rz1 / rz; rz = rz1;
    for (int i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  return it;
}


// ---------------------------------------------------------------------------------------------------------------
// g2o formulation (GLBA_MODE_G2O): the archived BA of the reference, Old/mult_img_recoverpose_single_ba:258-314 and
// docs/old_unorganized/4image_pnp_ba.txt:321-430 — VertexSE3Expmap / VertexSBAPointXYZ / EdgeProjectXYZ2UV with
// OptimizationAlgorithmLevenberg.  g2o itself is absent from /root/reference; what follows restates its
// published algorithm: world-to-camera pose T = (R, t), p = R X + t, error = observation - projection, update
// T <- exp([dw, dv]) T with SE3Quat::exp (rotation-first 6-vector), closed-form Jacobians of
// EdgeProjectXYZ2UV::linearizeOplus, robust kernels as rho' reweighting.
// ---------------------------------------------------------------------------------------------------------------
inline void so3_exp(const double w[3], double R[9], double* A_out = nullptr, double* B_out = nullptr) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double A, B;       // R = I + A [w]x + B [w]x^2
  if (th2 < 1e-8) { A = 1.0 - th2 / 6.0 + th2 * th2 / 120.0; B = 0.5 - th2 / 24.0 + th2 * th2 / 720.0; }
  else { const double th = std::sqrt(th2); A = std::sin(th) / th; B = (1.0 - std::cos(th)) / th2; }
  const double x = w[0], y = w[1], z = w[2];
  R[0] = 1.0 - B * (y * y + z * z); R[1] = -A * z + B * x * y;        R[2] = A * y + B * x * z;
  R[3] = A * z + B * x * y;         R[4] = 1.0 - B * (x * x + z * z); R[5] = -A * x + B * y * z;
  R[6] = -A * y + B * x * z;        R[7] = A * x + B * y * z;         R[8] = 1.0 - B * (x * x + y * y);
  if (A_out) *A_out = A;
  if (B_out) *B_out = B;
}
// exp of se(3): rotation exp([w]x), translation V v with V = I + B [w]x + C [w]x^2, C = (th - sin th)/th^3
inline void se3_exp(const double d[6], double R[9], double t[3]) {
  double B;
  so3_exp(d, R, nullptr, &B);
  const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double Cc;
  if (th2 < 1e-8) Cc = 1.0 / 6.0 - th2 / 120.0 + th2 * th2 / 5040.0;
  else { const double th = std::sqrt(th2); Cc = (th - std::sin(th)) / (th2 * th); }
  const double* w = d; const double* v = d + 3;
  const double wxv[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
  const double wxwxv[3] = {w[1] * wxv[2] - w[2] * wxv[1], w[2] * wxv[0] - w[0] * wxv[2], w[0] * wxv[1] - w[1] * wxv[0]};
  for (int i = 0; i < 3; ++i) t[i] = v[i] + B * wxv[i] + Cc * wxwxv[i];
}
// log of a rotation matrix through its unit quaternion (stable for every angle in [0, pi])
inline void so3_log(const double R[9], double w[3]) {
  double q[4];   // w, x, y, z
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) { const double s = std::sqrt(tr + 1.0) * 2.0; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  if (q[0] < 0.0) for (double& v : q) v = -v;
  const double n = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const double k = (n < 1e-10) ? 2.0 / q[0] : 2.0 * std::atan2(n, q[0]) / n;
  w[0] = k * q[1]; w[1] = k * q[2]; w[2] = k * q[3];
}

// poses as [R row-major 9 | t 3] per camera
bool evaluate_g2o(const View& V, const double* Rt, const double* pt, int loss, double loss_a, double* cost_out, Lin* lin) {
  const long n = V.n_obs;
  if (lin) { lin->r.resize(2 * n); lin->jc.resize(12 * n); lin->jp.resize(6 * n); }
  double cost = 0.0; int anybad = 0;
#pragma omp parallel for schedule(static) reduction(+ : cost) reduction(| : anybad)
  for (long k = 0; k < n; ++k) {
    const double* R = Rt + 12 * (long)V.obs_cam[k]; const double* t = R + 9;
    const double* X = pt + 3 * (long)V.obs_pt[k];
    const double x = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0];
    const double y = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1];
    const double z = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2];
    const double iz = 1.0 / z, iz2 = iz * iz;
    const double e0 = V.obs_u[k] - (V.fx * x * iz + V.cx), e1 = V.obs_v[k] - (V.fy * y * iz + V.cy);
    const double info = V.pt_info ? std::max(V.pt_info[V.obs_pt[k]], 0.0) : 1.0;          // edge->setInformation(info * I)
    double rho[3]; loss_eval(loss, loss_a, info * (e0 * e0 + e1 * e1), rho);
    cost += 0.5 * rho[0];
    if (!std::isfinite(e0) || !std::isfinite(e1)) anybad |= 1;
    if (!lin) continue;
    const double sq = std::sqrt(info) * std::sqrt(rho[1]);
    lin->r[2 * k] = sq * e0; lin->r[2 * k + 1] = sq * e1;
    // EdgeProjectXYZ2UV::linearizeOplus: d e / d X = -1/z [[fx,0,-fx x/z],[0,fy,-fy y/z]] R ;  d e / d [dw, dv]
    const double P[6] = {V.fx * iz, 0.0, -V.fx * x * iz2, 0.0, V.fy * iz, -V.fy * y * iz2};
    double* Jp = &lin->jp[6 * k]; double* Jc = &lin->jc[12 * k];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c)
      Jp[r * 3 + c] = -sq * (P[r * 3] * R[c] + P[r * 3 + 1] * R[3 + c] + P[r * 3 + 2] * R[6 + c]);
    Jc[0] = sq * (x * y * iz2 * V.fx);          Jc[1] = sq * (-(1.0 + x * x * iz2) * V.fx); Jc[2] = sq * (y * iz * V.fx);
    Jc[3] = sq * (-iz * V.fx);                  Jc[4] = 0.0;                                Jc[5] = sq * (x * iz2 * V.fx);
    Jc[6] = sq * ((1.0 + y * y * iz2) * V.fy);  Jc[7] = sq * (-x * y * iz2 * V.fy);         Jc[8] = sq * (-x * iz * V.fy);
    Jc[9] = 0.0;                                Jc[10] = sq * (-iz * V.fy);                 Jc[11] = sq * (y * iz2 * V.fy);
    for (int q = 0; q < 12; ++q) if (!std::isfinite(Jc[q])) anybad |= 1;
  }
  *cost_out = cost;
  return !anybad && std::isfinite(cost);
}

struct Timer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

bool use_dense(const glba_options& o, int nfc) {
  if (o.linsolve == GLBA_LINSOLVE_DENSE) return true;
  if (o.linsolve == GLBA_LINSOLVE_PCG) return false;
  return 6 * nfc <= 3000;   // the oracle prefers the exact solve wherever it is affordable
}

// Column scaling + LM diagonal + Schur + solve.  step (scaled space, = -y) for cams (6*n_cam) and points (3*n_pt).
struct StepOut { std::vector<double> sc, sp; int cg_iters = 0; bool ok = true; };

// OptimizationAlgorithmLevenberg::solve, run for o->max_iters iterations (SparseOptimizer::optimize(N) as called at
// docs/old_unorganized/4image_pnp_ba.txt:419-420).  Every TRIAL is recorded as one summary entry.
int solve_g2o(const glba_problem* p, const glba_options* o, glba_summary* sum) {
  Timer ttotal;
  std::memset(sum, 0, sizeof(*sum));
  View V; int st = build_view(p, V); if (st) { sum->status = st; return st; }
  const long n = V.n_obs;
  std::vector<double> Rt((size_t)12 * V.n_cam), pt(p->pt, p->pt + (size_t)3 * V.n_pt);
  for (int i = 0; i < V.n_cam; ++i) { so3_exp(p->cam + 6 * i, &Rt[12 * i]); for (int a = 0; a < 3; ++a) Rt[12 * i + 9 + a] = p->cam[6 * i + 3 + a]; }
  std::vector<double> Rt_c(Rt), pt_c(pt);
  Lin L; double cost;
  if (!evaluate_g2o(V, Rt.data(), pt.data(), o->loss, o->loss_scale, &cost, &L)) {
    sum->status = GLBA_E_NUMERIC; sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_NUMERIC; return GLBA_E_NUMERIC; }
  sum->n_linearizations = 1;
  std::vector<double> gc, gp;          // J' r  (= -b)
  auto gradient = [&]() {
    gc.assign((size_t)6 * V.n_cam, 0.0); gp.assign((size_t)3 * V.n_pt, 0.0);
    for (long k = 0; k < n; ++k) {
      const int i = V.obs_cam[k], j = V.obs_pt[k];
      const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k]; const double* r = &L.r[2 * k];
      if (V.cam_free[i]) for (int a = 0; a < 6; ++a) gc[6 * i + a] += Jc[a] * r[0] + Jc[6 + a] * r[1];
      if (V.pt_free[j]) for (int a = 0; a < 3; ++a) gp[3 * j + a] += Jp[a] * r[0] + Jp[3 + a] * r[1];
    }
    double m = 0; for (double v : gc) m = std::max(m, std::fabs(v)); for (double v : gp) m = std::max(m, std::fabs(v)); return m; };
  double gmax = gradient();
  // computeLambdaInit: tau * max over the non-fixed vertices of the Hessian diagonal
  double hmax = 0.0;
  {
    std::vector<double> dc((size_t)6 * V.n_cam, 0.0), dp((size_t)3 * V.n_pt, 0.0);
    for (long k = 0; k < n; ++k) {
      const int i = V.obs_cam[k], j = V.obs_pt[k];
      for (int a = 0; a < 6; ++a) dc[6 * i + a] += L.jc[12 * k + a] * L.jc[12 * k + a] + L.jc[12 * k + 6 + a] * L.jc[12 * k + 6 + a];
      for (int a = 0; a < 3; ++a) dp[3 * j + a] += L.jp[6 * k + a] * L.jp[6 * k + a] + L.jp[6 * k + 3 + a] * L.jp[6 * k + 3 + a];
    }
    for (int i = 0; i < V.n_cam; ++i) if (V.cam_free[i]) for (int a = 0; a < 6; ++a) hmax = std::max(hmax, dc[6 * i + a]);
    for (int j = 0; j < V.n_pt; ++j) if (V.pt_free[j]) for (int a = 0; a < 3; ++a) hmax = std::max(hmax, dp[3 * j + a]);
  }
  double lambda = o->g2o_tau * hmax, nu = 2.0;
  sum->initial_cost = cost; sum->cost[0] = cost; sum->cost_candidate[0] = cost; sum->radius[0] = lambda > 0 ? 1.0 / lambda : 0.0;
  sum->gradient_max_norm[0] = gmax;
  const bool dense = use_dense(*o, V.n_free_cam);
  int n_free_params = 6 * V.n_free_cam; for (int j = 0; j < V.n_pt; ++j) n_free_params += 3 * V.pt_free[j];
  int it = 0;
  sum->termination = GLBA_TERM_NO_CONVERGENCE; sum->stop_reason = GLBA_STOP_NONE;
  if (n_free_params == 0 || !(lambda > 0.0)) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_GRADIENT_TOL; }
  else for (;;) {
    if (sum->n_successful >= o->max_iters) { sum->stop_reason = GLBA_STOP_MAX_ITERS; break; }
    int n_rej = 0; bool terminate = false; double rho = 0.0;
    do {                                                     // trials of one g2o iteration
      if (it >= GLBA_MAX_ITERS) { terminate = true; break; }
      ++it;
      std::vector<double> Dc((size_t)6 * V.n_cam, std::sqrt(lambda)), Dp((size_t)3 * V.n_pt, std::sqrt(lambda));
      Schur S; bool ok = build_schur(V, L, Dc, Dp, dense, S);
      std::vector<double> yc((size_t)6 * V.n_free_cam, 0.0);
      int cg_it = 0;
      if (ok && V.n_free_cam > 0) {
        if (dense) { yc = S.rhs; ok = cholesky_solve(S.S, 6 * V.n_free_cam, yc); }
        else { int mi = o->cg_max_iters > 0 ? o->cg_max_iters : std::min(4000, 4 * 6 * V.n_free_cam);
               cg_it = pcg_solve(V, L, S, o->cg_rel_tol, mi, yc); if (cg_it < 0) { ok = false; cg_it = 0; } }
      }
      sum->cg_iters[it] = cg_it;
      std::vector<double> dc((size_t)6 * V.n_cam, 0.0), dp((size_t)3 * V.n_pt, 0.0);      // the update  d = -y
      double cand = std::numeric_limits<double>::max(), scale = 0.0, sn = 0.0;
      if (ok) {
        for (int i = 0; i < V.n_cam; ++i) { const int sl = V.cam_slot[i]; if (sl < 0) continue;
          for (int a = 0; a < 6; ++a) dc[6 * i + a] = -yc[6 * sl + a]; }
        for (int j = 0; j < V.n_pt; ++j) {
          if (!V.pt_free[j]) continue;
          double t3[3] = {S.gp[3 * j], S.gp[3 * j + 1], S.gp[3 * j + 2]};
          for (long t = V.trk_start[j]; t < V.trk_start[j + 1]; ++t) {
            const long k = V.trk_obs[t]; const int sl = V.cam_slot[V.obs_cam[k]]; if (sl < 0) continue;
            const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k];
            double a0 = 0, a1 = 0;
            for (int a = 0; a < 6; ++a) { a0 += Jc[a] * yc[6 * sl + a]; a1 += Jc[6 + a] * yc[6 * sl + a]; }
            for (int c = 0; c < 3; ++c) t3[c] -= Jp[c] * a0 + Jp[3 + c] * a1;
          }
          const double* Ci = &S.Cinv[(size_t)9 * j];
          for (int a = 0; a < 3; ++a) dp[3 * j + a] = -(Ci[a * 3] * t3[0] + Ci[a * 3 + 1] * t3[1] + Ci[a * 3 + 2] * t3[2]);
        }
        // computeScale: sum d (lambda d + b), b = -J'r
        for (int i = 0; i < V.n_cam; ++i) if (V.cam_free[i]) for (int a = 0; a < 6; ++a) { const double d = dc[6 * i + a]; scale += d * (lambda * d - gc[6 * i + a]); sn += d * d; }
        for (int j = 0; j < V.n_pt; ++j) if (V.pt_free[j]) for (int a = 0; a < 3; ++a) { const double d = dp[3 * j + a]; scale += d * (lambda * d - gp[3 * j + a]); sn += d * d; }
        // oplus: T <- exp(d) T ; X <- X + d
        for (int i = 0; i < V.n_cam; ++i) {
          if (!V.cam_free[i]) { std::memcpy(&Rt_c[12 * i], &Rt[12 * i], 12 * sizeof(double)); continue; }
          double dR[9], dt[3];
          se3_exp(&dc[6 * i], dR, dt);
          const double* R = &Rt[12 * i]; const double* t = R + 9; double* Rn = &Rt_c[12 * i];
          for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = dR[r * 3] * R[c] + dR[r * 3 + 1] * R[3 + c] + dR[r * 3 + 2] * R[6 + c];
            Rn[9 + r] = dR[r * 3] * t[0] + dR[r * 3 + 1] * t[1] + dR[r * 3 + 2] * t[2] + dt[r];
          }
        }
        for (int j = 0; j < V.n_pt; ++j) for (int a = 0; a < 3; ++a) pt_c[3 * j + a] = pt[3 * j + a] + (V.pt_free[j] ? dp[3 * j + a] : 0.0);
        if (!evaluate_g2o(V, Rt_c.data(), pt_c.data(), o->loss, o->loss_scale, &cand, nullptr)) cand = std::numeric_limits<double>::max();
      }
      // rho = (chi2 - chi2') / (scale + 1e-3), in half-chi2 units
      rho = (cand >= std::numeric_limits<double>::max()) ? -1.0 : (cost - cand) / (0.5 * scale + 0.5e-3);
      sum->cost_candidate[it] = cand; sum->step_norm[it] = std::sqrt(sn); sum->relative_decrease[it] = rho;
      if (rho > 0.0 && std::isfinite(cand)) {
        Rt = Rt_c; pt = pt_c;
        if (!evaluate_g2o(V, Rt.data(), pt.data(), o->loss, o->loss_scale, &cost, &L)) {
          sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_NUMERIC; terminate = true; }
        gmax = gradient();
        sum->n_linearizations++; sum->n_successful++; sum->accepted[it] = 1;
        const double alpha = std::min(1.0 - std::pow(2.0 * rho - 1.0, 3), 2.0 / 3.0);
        lambda *= std::max(1.0 / 3.0, alpha);
        nu = 2.0;
      } else {
        lambda *= nu; nu *= 2.0; ++n_rej;
        sum->accepted[it] = 0;
        if (!std::isfinite(lambda)) terminate = true;
      }
      sum->cost[it] = cost; sum->radius[it] = 1.0 / lambda; sum->gradient_max_norm[it] = gmax;
    } while (rho < 0.0 && n_rej < o->g2o_max_trials && !terminate);
    if (terminate || n_rej >= o->g2o_max_trials || rho == 0.0) {
      if (sum->stop_reason == GLBA_STOP_NONE) sum->stop_reason = GLBA_STOP_TRIALS;
      break;
    }
  }
  sum->n_iters = it; sum->final_cost = cost;
  sum->status = GLBA_OK;
  if (sum->termination != GLBA_TERM_FAILURE) {
    for (int i = 0; i < V.n_cam; ++i) {
      if (!V.cam_free[i]) continue;                   // constant vertices keep their input bits
      so3_log(&Rt[12 * i], p->cam + 6 * i);
      for (int a = 0; a < 3; ++a) p->cam[6 * i + 3 + a] = Rt[12 * i + 9 + a];
    }
    std::memcpy(p->pt, pt.data(), sizeof(double) * pt.size());
  }
  sum->t_total_ms = ttotal.ms();
  return GLBA_OK;
}

}  // namespace

extern "C" {

void glbao_default_options(glba_options* o) {
  std::memset(o, 0, sizeof(*o));
  o->loss = GLBA_LOSS_CAUCHY; o->loss_scale = 1.0; o->max_iters = 30;
  o->function_tol = 1e-6; o->gradient_tol = 1e-10; o->parameter_tol = 1e-8;
  o->initial_radius = 1e4; o->max_radius = 1e16; o->min_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32;
  o->jacobi_scaling = 1; o->max_consecutive_invalid_steps = 5;
  o->linsolve = GLBA_LINSOLVE_AUTO; o->dense_max_dim = 96; o->cg_rel_tol = 1e-13; o->cg_max_iters = 0; o->verbose = 0;
  o->mode = GLBA_MODE_CERES; o->g2o_tau = 1e-5; o->g2o_max_trials = 10;
}

// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline arm asks for the box's cores explicitly
void glbao_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int glbao_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Cost only (what Ceres evaluates at the candidate point).
int glbao_cost(const glba_problem* p, const glba_options* o, double* cost) {
  View V; int st = build_view(p, V); if (st) return st;
  return evaluate(V, p->cam, p->pt, o->loss, o->loss_scale, cost, nullptr) ? GLBA_OK : GLBA_E_NUMERIC;
}

// One linearisation at `radius` in UNSCALED variables, derived from the Ceres-style scaled system:
//   s_k = 1/(1+|col_k|), d_k = clamp(|s_k col_k|^2), Lambda_k = d_k/(radius s_k^2)
//   S = B + Lambda_c - W (C+Lambda_p)^-1 W',  rhs = g_c - W (C+Lambda_p)^-1 g_p
int glbao_linearize(const glba_problem* p, const glba_options* o, double radius, glba_linearization* out) {
  View V; int st = build_view(p, V); if (st) return st;
  Lin L; double cost;
  Timer t0;
  if (!evaluate(V, p->cam, p->pt, o->loss, o->loss_scale, &cost, &L)) return GLBA_E_NUMERIC;
  out->cost = cost;
  const long n = V.n_obs;
  if (out->residuals) std::memcpy(out->residuals, L.r.data(), sizeof(double) * 2 * n);
  if (out->jac_cam) std::memcpy(out->jac_cam, L.jc.data(), sizeof(double) * 12 * n);
  if (out->jac_pt) std::memcpy(out->jac_pt, L.jp.data(), sizeof(double) * 6 * n);
  // unscaled Hessian blocks / gradient
  std::vector<double> B((size_t)36 * V.n_cam, 0.0), C((size_t)9 * V.n_pt, 0.0), gc((size_t)6 * V.n_cam, 0.0), gp((size_t)3 * V.n_pt, 0.0);
  for (long k = 0; k < n; ++k) {
    const int i = V.obs_cam[k], j = V.obs_pt[k];
    const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k]; const double* r = &L.r[2 * k];
    if (V.cam_free[i]) for (int a = 0; a < 6; ++a) {
      for (int c = 0; c < 6; ++c) B[(size_t)36 * i + a * 6 + c] += Jc[a] * Jc[c] + Jc[6 + a] * Jc[6 + c];
      gc[6 * i + a] += Jc[a] * r[0] + Jc[6 + a] * r[1];
    }
    if (V.pt_free[j]) for (int a = 0; a < 3; ++a) {
      for (int c = 0; c < 3; ++c) C[(size_t)9 * j + a * 3 + c] += Jp[a] * Jp[c] + Jp[3 + a] * Jp[3 + c];
      gp[3 * j + a] += Jp[a] * r[0] + Jp[3 + a] * r[1];
    }
  }
  out->t_linearize_ms = t0.ms();
  if (out->grad_cam) std::memcpy(out->grad_cam, gc.data(), sizeof(double) * gc.size());
  if (out->grad_pt) std::memcpy(out->grad_pt, gp.data(), sizeof(double) * gp.size());
  if (out->hess_cam) std::memcpy(out->hess_cam, B.data(), sizeof(double) * B.size());
  if (out->hess_pt) std::memcpy(out->hess_pt, C.data(), sizeof(double) * C.size());
  Timer t1;
  // scaled system exactly as the LM iteration builds it
  std::vector<double> sc((size_t)6 * V.n_cam, 1.0), sp((size_t)3 * V.n_pt, 1.0);
  if (o->jacobi_scaling) {
    for (int i = 0; i < V.n_cam; ++i) for (int a = 0; a < 6; ++a) sc[6 * i + a] = 1.0 / (1.0 + std::sqrt(B[(size_t)36 * i + a * 7]));
    for (int j = 0; j < V.n_pt; ++j) for (int a = 0; a < 3; ++a) sp[3 * j + a] = 1.0 / (1.0 + std::sqrt(C[(size_t)9 * j + a * 4]));
  }
  Lin Ls = L;
  for (long k = 0; k < n; ++k) {
    const int i = V.obs_cam[k], j = V.obs_pt[k];
    for (int row = 0; row < 2; ++row) {
      for (int a = 0; a < 6; ++a) Ls.jc[12 * k + row * 6 + a] *= sc[6 * i + a];
      for (int a = 0; a < 3; ++a) Ls.jp[6 * k + row * 3 + a] *= sp[3 * j + a];
    }
  }
  std::vector<double> dc((size_t)6 * V.n_cam, 0.0), dp((size_t)3 * V.n_pt, 0.0);
  for (long k = 0; k < n; ++k) {
    const int i = V.obs_cam[k], j = V.obs_pt[k];
    for (int a = 0; a < 6; ++a) dc[6 * i + a] += Ls.jc[12 * k + a] * Ls.jc[12 * k + a] + Ls.jc[12 * k + 6 + a] * Ls.jc[12 * k + 6 + a];
    for (int a = 0; a < 3; ++a) dp[3 * j + a] += Ls.jp[6 * k + a] * Ls.jp[6 * k + a] + Ls.jp[6 * k + 3 + a] * Ls.jp[6 * k + 3 + a];
  }
  for (auto& d : dc) d = std::sqrt(std::min(std::max(d, o->min_lm_diagonal), o->max_lm_diagonal) / radius);
  for (auto& d : dp) d = std::sqrt(std::min(std::max(d, o->min_lm_diagonal), o->max_lm_diagonal) / radius);
  Schur S;
  if (!build_schur(V, Ls, dc, dp, false, S)) return GLBA_E_NUMERIC;
  out->t_schur_ms = t1.ms();
  if (out->schur_diag) {
    std::memset(out->schur_diag, 0, sizeof(double) * 36 * V.n_cam);
    for (int i = 0; i < V.n_cam; ++i) { const int s = V.cam_slot[i]; if (s < 0) continue;
      for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c)
        out->schur_diag[(size_t)36 * i + a * 6 + c] = S.Sdiag[(size_t)36 * s + a * 6 + c] / (sc[6 * i + a] * sc[6 * i + c]); }
  }
  if (out->schur_rhs) {
    std::memset(out->schur_rhs, 0, sizeof(double) * 6 * V.n_cam);
    for (int i = 0; i < V.n_cam; ++i) { const int s = V.cam_slot[i]; if (s < 0) continue;
      for (int a = 0; a < 6; ++a) out->schur_rhs[6 * i + a] = S.rhs[6 * s + a] / sc[6 * i + a]; }
  }
  return GLBA_OK;
}

// One linearise + Schur pass exactly as an LM iteration performs it, fully threaded, no outputs but the
// cost: evaluate (autodiff) -> Jacobi column scaling -> LM diagonal -> Schur elimination (diagonal blocks of S
// for block-Jacobi, reduced rhs).  This is what bench.py times as the CPU baseline of the headline metric.
int glbao_step(const glba_problem* p, const glba_options* o, double radius, double* cost_out, double* t_eval_ms, double* t_schur_ms) {
  View V; int st = build_view(p, V); if (st) return st;
  Lin L; double cost;
  Timer t0;
  if (!evaluate(V, p->cam, p->pt, o->loss, o->loss_scale, &cost, &L)) return GLBA_E_NUMERIC;
  if (t_eval_ms) *t_eval_ms = t0.ms();
  Timer t1;
  const long n = V.n_obs;
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  std::vector<double> dp((size_t)3 * V.n_pt, 0.0), dc((size_t)6 * V.n_cam, 0.0);
  std::vector<std::vector<double>> tdc(nthreads);
#pragma omp parallel num_threads(nthreads)
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    std::vector<double>& my = tdc[tid]; my.assign((size_t)6 * V.n_cam, 0.0);
#pragma omp for schedule(static)
    for (int j = 0; j < V.n_pt; ++j)
      for (long t = V.trk_start[j]; t < V.trk_start[j + 1]; ++t) {
        const long k = V.trk_obs[t]; const int i = V.obs_cam[k];
        for (int a = 0; a < 6; ++a) my[6 * i + a] += L.jc[12 * k + a] * L.jc[12 * k + a] + L.jc[12 * k + 6 + a] * L.jc[12 * k + 6 + a];
        for (int a = 0; a < 3; ++a) dp[3 * j + a] += L.jp[6 * k + a] * L.jp[6 * k + a] + L.jp[6 * k + 3 + a] * L.jp[6 * k + 3 + a];
      }
  }
  for (int t = 0; t < nthreads; ++t) for (size_t i = 0; i < dc.size(); ++i) dc[i] += tdc[t][i];
  std::vector<double> sc(dc.size(), 1.0), sp(dp.size(), 1.0);
  if (o->jacobi_scaling) {
    for (size_t i = 0; i < dc.size(); ++i) sc[i] = 1.0 / (1.0 + std::sqrt(dc[i]));
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)dp.size(); ++i) sp[i] = 1.0 / (1.0 + std::sqrt(dp[i]));
  }
#pragma omp parallel for schedule(static)
  for (long k = 0; k < n; ++k) {
    const int i = V.obs_cam[k], j = V.obs_pt[k];
    for (int row = 0; row < 2; ++row) {
      for (int a = 0; a < 6; ++a) L.jc[12 * k + row * 6 + a] *= sc[6 * i + a];
      for (int a = 0; a < 3; ++a) L.jp[6 * k + row * 3 + a] *= sp[3 * j + a];
    }
  }
  for (size_t i = 0; i < dc.size(); ++i) dc[i] = std::sqrt(std::min(std::max(dc[i] * sc[i] * sc[i], o->min_lm_diagonal), o->max_lm_diagonal) / radius);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < (long)dp.size(); ++i) dp[i] = std::sqrt(std::min(std::max(dp[i] * sp[i] * sp[i], o->min_lm_diagonal), o->max_lm_diagonal) / radius);
  Schur S;
  if (!build_schur(V, L, dc, dp, false, S)) return GLBA_E_NUMERIC;
  if (t_schur_ms) *t_schur_ms = t1.ms();
  if (cost_out) *cost_out = cost;
  return GLBA_OK;
}

// Ceres-semantics LM (replaces ceres::Solve at slam_core.cpp:849).
int glbao_solve(const glba_problem* p, const glba_options* o, glba_summary* sum) {
  if (o->mode == GLBA_MODE_G2O) return solve_g2o(p, o, sum);
  Timer ttotal;
  std::memset(sum, 0, sizeof(*sum));
  View V; int st = build_view(p, V); if (st) { sum->status = st; return st; }
  if (o->max_iters > GLBA_MAX_ITERS) { sum->status = GLBA_E_INVALID_ARG; return GLBA_E_INVALID_ARG; }
  std::vector<double> cam(p->cam, p->cam + (size_t)6 * V.n_cam), pt(p->pt, p->pt + (size_t)3 * V.n_pt);
  std::vector<double> cam_c(cam), pt_c(pt);
  Lin L; double cost;
  Timer tl;
  if (!evaluate(V, cam.data(), pt.data(), o->loss, o->loss_scale, &cost, &L)) {
    sum->status = GLBA_E_NUMERIC; sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_NUMERIC; return GLBA_E_NUMERIC; }
  sum->t_linearize_ms += tl.ms();
  sum->n_linearizations = 1;
  const long n = V.n_obs;
  auto norm_x = [&](const std::vector<double>& c, const std::vector<double>& x) {
    double s = 0;
    for (int i = 0; i < V.n_cam; ++i) if (V.cam_free[i]) for (int a = 0; a < 6; ++a) s += c[6 * i + a] * c[6 * i + a];
    for (int j = 0; j < V.n_pt; ++j) if (V.pt_free[j]) for (int a = 0; a < 3; ++a) s += x[3 * j + a] * x[3 * j + a];
    return std::sqrt(s); };
  // gradient (unscaled J) and its max-norm
  std::vector<double> gc, gp;
  auto gradient_max = [&]() {
    gc.assign((size_t)6 * V.n_cam, 0.0); gp.assign((size_t)3 * V.n_pt, 0.0);
    for (long k = 0; k < n; ++k) {
      const int i = V.obs_cam[k], j = V.obs_pt[k];
      const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k]; const double* r = &L.r[2 * k];
      if (V.cam_free[i]) for (int a = 0; a < 6; ++a) gc[6 * i + a] += Jc[a] * r[0] + Jc[6 + a] * r[1];
      if (V.pt_free[j]) for (int a = 0; a < 3; ++a) gp[3 * j + a] += Jp[a] * r[0] + Jp[3 + a] * r[1];
    }
    double m = 0; for (double v : gc) m = std::max(m, std::fabs(v)); for (double v : gp) m = std::max(m, std::fabs(v)); return m; };
  double gmax = gradient_max();
  // Jacobi scaling, fixed at iteration 0
  std::vector<double> sc((size_t)6 * V.n_cam, 1.0), sp((size_t)3 * V.n_pt, 1.0);
  auto col_sqnorm = [&](std::vector<double>& dc, std::vector<double>& dp) {
    dc.assign((size_t)6 * V.n_cam, 0.0); dp.assign((size_t)3 * V.n_pt, 0.0);
    for (long k = 0; k < n; ++k) {
      const int i = V.obs_cam[k], j = V.obs_pt[k];
      for (int a = 0; a < 6; ++a) dc[6 * i + a] += L.jc[12 * k + a] * L.jc[12 * k + a] + L.jc[12 * k + 6 + a] * L.jc[12 * k + 6 + a];
      for (int a = 0; a < 3; ++a) dp[3 * j + a] += L.jp[6 * k + a] * L.jp[6 * k + a] + L.jp[6 * k + 3 + a] * L.jp[6 * k + 3 + a];
    } };
  if (o->jacobi_scaling) {
    std::vector<double> dc, dp; col_sqnorm(dc, dp);
    for (size_t i = 0; i < dc.size(); ++i) sc[i] = 1.0 / (1.0 + std::sqrt(dc[i]));
    for (size_t i = 0; i < dp.size(); ++i) sp[i] = 1.0 / (1.0 + std::sqrt(dp[i]));
  }
  auto scale_columns = [&]() {
    for (long k = 0; k < n; ++k) {
      const int i = V.obs_cam[k], j = V.obs_pt[k];
      for (int row = 0; row < 2; ++row) {
        for (int a = 0; a < 6; ++a) L.jc[12 * k + row * 6 + a] *= sc[6 * i + a];
        for (int a = 0; a < 3; ++a) L.jp[6 * k + row * 3 + a] *= sp[3 * j + a];
      } } };
  scale_columns();
  double x_norm = norm_x(cam, pt);
  double radius = o->initial_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false; int n_invalid = 0;
  std::vector<double> diag_c, diag_p, Dc, Dp;
  sum->initial_cost = cost; sum->cost[0] = cost; sum->cost_candidate[0] = cost; sum->radius[0] = radius;
  sum->gradient_max_norm[0] = gmax; sum->accepted[0] = 0;
  const bool dense = use_dense(*o, V.n_free_cam);
  int n_free_params = 6 * V.n_free_cam; for (int j = 0; j < V.n_pt; ++j) n_free_params += 3 * V.pt_free[j];
  int it = 0;
  sum->termination = GLBA_TERM_NO_CONVERGENCE; sum->stop_reason = GLBA_STOP_NONE;
  if (n_free_params == 0) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_GRADIENT_TOL; }
  else for (;;) {
    if (it >= o->max_iters) { sum->termination = GLBA_TERM_NO_CONVERGENCE; sum->stop_reason = GLBA_STOP_MAX_ITERS; break; }
    if (gmax <= o->gradient_tol) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_GRADIENT_TOL; break; }
    if (radius <= o->min_radius) { sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_MIN_RADIUS; break; }
    ++it;
    // --- ComputeTrustRegionStep ---
    Timer ts;
    if (!reuse_diagonal) {
      col_sqnorm(diag_c, diag_p);
      for (auto& d : diag_c) d = std::min(std::max(d, o->min_lm_diagonal), o->max_lm_diagonal);
      for (auto& d : diag_p) d = std::min(std::max(d, o->min_lm_diagonal), o->max_lm_diagonal);
    }
    Dc.resize(diag_c.size()); Dp.resize(diag_p.size());
    for (size_t i = 0; i < Dc.size(); ++i) Dc[i] = std::sqrt(diag_c[i] / radius);
    for (size_t i = 0; i < Dp.size(); ++i) Dp[i] = std::sqrt(diag_p[i] / radius);
    reuse_diagonal = true;
    Schur S; bool ok = build_schur(V, L, Dc, Dp, dense, S);
    sum->t_schur_ms += ts.ms();
    Timer tv;
    std::vector<double> yc((size_t)6 * V.n_free_cam, 0.0);
    int cg_it = 0;
    if (ok && V.n_free_cam > 0) {
      if (dense) { yc = S.rhs; ok = cholesky_solve(S.S, 6 * V.n_free_cam, yc); }
      else { int mi = o->cg_max_iters > 0 ? o->cg_max_iters : std::min(4000, 4 * 6 * V.n_free_cam);
             cg_it = pcg_solve(V, L, S, o->cg_rel_tol, mi, yc); if (cg_it < 0) { ok = false; cg_it = 0; } }
    }
    // back-substitution  y_p = Cinv (g_p - W' y_c);  step = -y
    std::vector<double> step_c((size_t)6 * V.n_cam, 0.0), step_p((size_t)3 * V.n_pt, 0.0);
    if (ok) {
      for (int i = 0; i < V.n_cam; ++i) { const int s = V.cam_slot[i]; if (s < 0) continue;
        for (int a = 0; a < 6; ++a) step_c[6 * i + a] = -yc[6 * s + a]; }
#pragma omp parallel for schedule(static)
      for (int j = 0; j < V.n_pt; ++j) {
        if (!V.pt_free[j]) continue;
        double t3[3] = {S.gp[3 * j], S.gp[3 * j + 1], S.gp[3 * j + 2]};
        for (long t = V.trk_start[j]; t < V.trk_start[j + 1]; ++t) {
          const long k = V.trk_obs[t]; const int s = V.cam_slot[V.obs_cam[k]]; if (s < 0) continue;
          const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k];
          double a0 = 0, a1 = 0;
          for (int a = 0; a < 6; ++a) { a0 += Jc[a] * yc[6 * s + a]; a1 += Jc[6 + a] * yc[6 * s + a]; }
          for (int c = 0; c < 3; ++c) t3[c] -= Jp[c] * a0 + Jp[3 + c] * a1;
        }
        const double* Ci = &S.Cinv[(size_t)9 * j];
        for (int a = 0; a < 3; ++a) step_p[3 * j + a] = -(Ci[a * 3] * t3[0] + Ci[a * 3 + 1] * t3[1] + Ci[a * 3 + 2] * t3[2]);
      }
    }
    sum->t_solve_ms += tv.ms();
    sum->cg_iters[it] = cg_it;
    // model_cost_change = -(J step).(r + J step / 2)
    double model_cost_change = 0.0;
    if (ok) {
      double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
      for (long k = 0; k < n; ++k) {
        const int i = V.obs_cam[k], j = V.obs_pt[k];
        const double* Jc = &L.jc[12 * k]; const double* Jp = &L.jp[6 * k]; const double* r = &L.r[2 * k];
        double m0 = 0, m1 = 0;
        if (V.cam_free[i]) for (int a = 0; a < 6; ++a) { m0 += Jc[a] * step_c[6 * i + a]; m1 += Jc[6 + a] * step_c[6 * i + a]; }
        if (V.pt_free[j]) for (int a = 0; a < 3; ++a) { m0 += Jp[a] * step_p[3 * j + a]; m1 += Jp[3 + a] * step_p[3 * j + a]; }
        acc += m0 * (r[0] + m0 / 2.0) + m1 * (r[1] + m1 / 2.0);
      }
      model_cost_change = -acc;
    }
    const bool valid = ok && (model_cost_change > 0.0);
    if (!valid) {
      ++n_invalid;
      sum->cost[it] = cost; sum->cost_candidate[it] = cost; sum->step_norm[it] = 0; sum->relative_decrease[it] = 0;
      sum->gradient_max_norm[it] = gmax; sum->accepted[it] = 0;
      if (n_invalid >= o->max_consecutive_invalid_steps) {
        sum->radius[it] = radius; sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_INVALID_STEPS; break; }
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      sum->radius[it] = radius;
      continue;
    }
    n_invalid = 0;
    Timer tu;
    for (int i = 0; i < V.n_cam; ++i) for (int a = 0; a < 6; ++a)
      cam_c[6 * i + a] = V.cam_free[i] ? cam[6 * i + a] + step_c[6 * i + a] * sc[6 * i + a] : cam[6 * i + a];
    for (int j = 0; j < V.n_pt; ++j) for (int a = 0; a < 3; ++a)
      pt_c[3 * j + a] = V.pt_free[j] ? pt[3 * j + a] + step_p[3 * j + a] * sp[3 * j + a] : pt[3 * j + a];
    double cand_cost;
    if (!evaluate(V, cam_c.data(), pt_c.data(), o->loss, o->loss_scale, &cand_cost, nullptr)) cand_cost = std::numeric_limits<double>::max();
    double sn = 0;
    for (int i = 0; i < V.n_cam; ++i) if (V.cam_free[i]) for (int a = 0; a < 6; ++a) { const double d = cam[6 * i + a] - cam_c[6 * i + a]; sn += d * d; }
    for (int j = 0; j < V.n_pt; ++j) if (V.pt_free[j]) for (int a = 0; a < 3; ++a) { const double d = pt[3 * j + a] - pt_c[3 * j + a]; sn += d * d; }
    const double step_norm = std::sqrt(sn);
    sum->t_update_ms += tu.ms();
    sum->cost_candidate[it] = cand_cost; sum->step_norm[it] = step_norm;
    sum->cost[it] = cost; sum->radius[it] = radius; sum->gradient_max_norm[it] = gmax;
    if (step_norm <= o->parameter_tol * (x_norm + o->parameter_tol)) {
      sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_PARAMETER_TOL; break; }
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= o->function_tol * cost) {
      sum->termination = GLBA_TERM_CONVERGENCE; sum->stop_reason = GLBA_STOP_FUNCTION_TOL; break; }
    const double rel = (cand_cost >= std::numeric_limits<double>::max()) ? std::numeric_limits<double>::lowest()
                                                                         : cost_change / model_cost_change;
    sum->relative_decrease[it] = rel;
    if (rel > o->min_relative_decrease) {
      cam = cam_c; pt = pt_c; x_norm = norm_x(cam, pt);
      Timer tl2;
      if (!evaluate(V, cam.data(), pt.data(), o->loss, o->loss_scale, &cost, &L)) {
        sum->termination = GLBA_TERM_FAILURE; sum->stop_reason = GLBA_STOP_NUMERIC; break; }
      gmax = gradient_max();
      scale_columns();
      sum->t_linearize_ms += tl2.ms();
      sum->n_linearizations++;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
      radius = std::min(o->max_radius, radius);
      decrease_factor = 2.0; reuse_diagonal = false;
      sum->n_successful++; sum->accepted[it] = 1;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      sum->accepted[it] = 0;
    }
    sum->cost[it] = cost; sum->radius[it] = radius; sum->gradient_max_norm[it] = gmax;
  }
  sum->n_iters = it; sum->final_cost = cost;
  sum->status = GLBA_OK;
  if (sum->termination != GLBA_TERM_FAILURE) {   // Ceres writes back the best (= last accepted) state
    std::memcpy(p->cam, cam.data(), sizeof(double) * cam.size());
    std::memcpy(p->pt, pt.data(), sizeof(double) * pt.size());
  }
  sum->t_total_ms = ttotal.ms();
  return GLBA_OK;
}

// pose_only_ba (slam_core.cpp:1092-1140): one free camera, all points constant.
int glbao_pose_only(double* cam, int32_t n, const double* X, const double* uv, double fx, double fy, double cx,
                    double cy, const glba_options* o, glba_summary* sum) {
  if (n <= 0 || !cam || !X || !uv) return GLBA_E_INVALID_ARG;
  std::vector<int32_t> oc(n, 0), op(n); std::vector<double> u(n), v(n); std::vector<uint8_t> pf(n, 1);
  std::vector<double> pts(X, X + (size_t)3 * n);
  for (int k = 0; k < n; ++k) { op[k] = k; u[k] = uv[2 * k]; v[k] = uv[2 * k + 1]; }
  glba_problem p; std::memset(&p, 0, sizeof(p));
  p.n_cam = 1; p.n_pt = n; p.n_obs = n; p.cam = cam; p.pt = pts.data();
  p.obs_cam = oc.data(); p.obs_pt = op.data(); p.obs_u = u.data(); p.obs_v = v.data();
  p.cam_fixed = nullptr; p.pt_fixed = pf.data(); p.fx = fx; p.fy = fy; p.cx = cx; p.cy = cy;
  return glbao_solve(&p, o, sum);
}

// post_ba_map_point_culling arithmetic (slam_core.cpp:993-1035): per point over all its observations,
// depth <= 0 -> bad; mean pixel error > max_mean_err or n_obs < min_obs -> bad.  Uses R = Rodrigues(w).
int glbao_cull_points(const glba_problem* p, int32_t min_obs, double max_mean_err, uint8_t* bad, double* mean_err) {
  View V; int st = build_view(p, V); if (st) return st;
  for (int j = 0; j < V.n_pt; ++j) {
    double tot = 0; int cnt = 0; bool isbad = false;
    for (long t = V.trk_start[j]; t < V.trk_start[j + 1]; ++t) {
      const long k = V.trk_obs[t];
      const double* c = p->cam + 6 * (long)V.obs_cam[k]; const double* x = p->pt + 3 * (long)j;
      double q[3] = {x[0] - c[3], x[1] - c[4], x[2] - c[5]}, mw[3] = {-c[0], -c[1], -c[2]}, pc[3];
      angle_axis_rotate_point<double>(mw, q, pc);
      if (pc[2] <= 0) { isbad = true; break; }
      const double du = V.fx * pc[0] / pc[2] + V.cx - V.obs_u[k], dv = V.fy * pc[1] / pc[2] + V.cy - V.obs_v[k];
      tot += std::sqrt(du * du + dv * dv); cnt++;
    }
    double avg = cnt ? tot / cnt : 0.0;
    if (!isbad && (cnt < min_obs || avg > max_mean_err)) isbad = true;
    bad[j] = isbad ? 1 : 0; if (mean_err) mean_err[j] = avg;
  }
  return GLBA_OK;
}

}  // extern "C"
