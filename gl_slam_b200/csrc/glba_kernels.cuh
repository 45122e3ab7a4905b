// glba_kernels.cuh — sm_100a device code of the bundle-adjustment backend (FP64, HBM-bound).
//
// Replaces, on the GPU, the arithmetic Ceres performs for slam_core::full_ba (GL-SLAM
// src/core/slam_core.cpp:799-849): per-observation residual / robust weight / Jacobians
// (ReprojectionError, :699-733, CauchyLoss :814), Hessian block assembly, Schur complement,
// PCG on the reduced camera system, back-substitution and candidate cost.
//
// Design: the Jacobian is never materialised.  Each observation keeps one 32-byte
// record (xh, yh, 1/z, w): normalised image coordinates, inverse depth and sqrt(rho').  Every
// Jacobian block is rebuilt from it and two per-camera 3x3 matrices that live in L1/L2:
//     J~_p = w P R,   J~_c = J^ T,   J^ = w [Q | -P] (2x6),   T = blockdiag(G, R)
//     P = 1/z [[fx,0,-fx xh],[0,fy,-fy yh]],  Q = [[fx,0,-fx xh],[0,fy,-fy yh]] [m]x
//     R = R(-w) (world->camera),  G = J_r(w) (right Jacobian of SO(3); additive angle-axis),
//     m = (xh,yh,1) + sv x (xh,yh,1)  with sv = w only on Ceres' small-angle branch, else 0.
// Records exist in two orders: point-major (track-contiguous) for everything reduced per point,
// camera-major for everything reduced per camera.  Both reductions are fixed-order (no atomics),
// so results are bit-reproducible run to run.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

namespace glba {

constexpr int CAMTAB = 24;  // doubles per camera: R[9] c[3] | G[9] sv[3]  (192 B; R,c = the first three 32-byte sectors)
constexpr int CT_C = 9, CT_G = 12, CT_SV = 21;
constexpr int PBLK = 12;    // doubles per point block: Cinv[6] (00,01,02,11,12,22) u0[3] pad[3]  (96 B = 3 sectors)
constexpr int NT_PM = 128;  // threads per CTA, point-major kernels (one thread per point)
constexpr int NT_CM = 256;  // threads per CTA, k_spmv_cm (one CTA per chunk of one camera)
#ifndef GLBA_NT_HCM
#define GLBA_NT_HCM 128
#endif
constexpr int NT_HCM = GLBA_NT_HCM;  // threads per CTA of the two 27-accumulator kernels (k_linearize_cm, k_schur_cm)
constexpr int NT_CAM = 1024;  // threads of the single-CTA camera kernels

// scalar slots (device array `scal`), written by fixed-order reductions
constexpr int MAX_WORLD = 16;
enum Scal {
  // [0..3] point-derived sums of a linearisation; they sit right behind the per-camera partial sums so ONE
  // all-reduce carries both when the map is sharded
  S_COST = 0,     // 1/2 sum rho at the linearisation point
  S_XN2_P,        // |x|^2 over free points
  S_BAD,          // non-finite residuals at the linearisation point (count)
  S_NOTPD_P,      // point blocks that failed LDL'
  S_GSLOT0,       // [4..19] one slot per rank: max |g| over that rank's free points (sum-reduce, then max locally)
  S_COST_C = S_GSLOT0 + MAX_WORLD,   // [20..24] candidate-step sums, one small all-reduce per LM iteration
  S_YN2_P,        // |y_p|^2
  S_YG_P,         // y_p . g_p
  S_YLY_P,        // y_p' Lambda_p y_p
  S_BAD_C,        // non-finite residuals at the candidate
  S_GMAX_P,       // max |g| over free points (all ranks)
  // camera-derived, replicated on every rank
  S_XN2_C, S_GMAX_C, S_YN2_C, S_YG_C, S_YLY_C, S_NOTPD_C,
  S_COUNT
};
constexpr int NSCAL = 40;

// Programmatic dependent launch (sm_90+): every kernel of this library starts with pdl_grid_sync().  launch_dependents lets the
// NEXT kernel of the stream be scheduled as soon as all CTAs of this one have started (its launch latency hides behind this
// kernel); wait blocks until every prerequisite grid has completed and its memory is visible, so data dependencies are those of
// plain stream order.  Every thread executes both before anything else (also before any early exit: a kernel that finished
// without waiting would release ITS dependents too early).  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct Intr { double fx, fy, cx, cy; };

struct LossP { int kind; double a; };

struct CgState {
  double g[2], a[2];   // gamma = r.u and alpha of the previous iteration, as launch li reads them: g[li & 1], a[li & 1]
  double gamma0, tol;
  double dot[2];       // single GPU: gamma' = r.u and delta = u.S u of the current iteration (k_cg_w -> k_cg_update)
  int done_at;         // kernels of PCG launch li exit when done_at <= li (published by launch done_at-1)
  int reason;          // 1 converged, 2 breakdown (p'Sp <= 0), 3 iteration cap
  int iters, max_iters;
};
// Device-resident trust-region control (glba_lm.cuh): the decision kernel writes it, every kernel of an LM iteration that the
// host enqueues WITHOUT knowing the decision reads it.  A null pointer means "host-driven": run, and use the by-value radius.
struct LmCtl {
  double inv_radius;       // 1 / trust-region radius of the step being computed
  int done;                // the loop has terminated: everything enqueued behind exits at once
  int accepted;            // last step accepted: the candidate becomes current and is re-linearised
  int need_redamp;         // last step rejected / invalid: point blocks are re-damped for the new radius
  // trust-region state (k_lm_decide / k_lm_absorb only)
  double radius, decrease_factor, cost, gmax, x_norm;
  int it, n_invalid, n_rejected, pad;
};
enum LmGate { GATE_ALWAYS = 0, GATE_ACCEPTED = 1, GATE_REDAMP = 2 };
__device__ __forceinline__ bool ctl_skip(const LmCtl* c, const int gate) {
  return c != nullptr && (c->done != 0 || (gate == GATE_ACCEPTED && c->accepted == 0) || (gate == GATE_REDAMP && c->need_redamp == 0));
}
__device__ __forceinline__ double ctl_inv_radius(const LmCtl* c, const double by_value) { return c != nullptr ? c->inv_radius : by_value; }

constexpr int XTAB = 16;   // per-camera gather row of the point passes (128 B): xg[6] = T x | R[9] | small-angle flag

// ---------------------------------------------------------------------------------------------
// sm_100 256-bit global accesses (SASS LDG.E.ENL2.256 / STG.E.ENL2.256): one instruction, one L1 wavefront
// per 32-byte record.  Every record / table row here is 32-byte aligned.
__device__ __forceinline__ double4 ldg4(const double4* p) {           // read-only (non-coherent) path
  double4 r;      // not volatile, no memory clobber: read-only data, the compiler may hoist / batch these freely
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double4 ld4(const double4* p) {            // coherent (data written by this kernel)
  double4 r;
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st4(double4* p, const double4 v) {
  // volatile (has no outputs) but no memory clobber: records are write-only here, later loads may pass the store
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w));
}
// first 96 bytes of a camera-table row: R (row-major) and the centre, as three 256-bit gathers
__device__ __forceinline__ void load_Rc(const double* __restrict__ ct, double* R /*9*/, double* c /*3*/) {
  const double4* q = reinterpret_cast<const double4*>(ct);
  const double4 a = ldg4(q), b = ldg4(q + 1), d = ldg4(q + 2);
  R[0] = a.x; R[1] = a.y; R[2] = a.z; R[3] = a.w; R[4] = b.x; R[5] = b.y; R[6] = b.z; R[7] = b.w; R[8] = d.x;
  c[0] = d.y; c[1] = d.z; c[2] = d.w;
}
// Damped point block, one 96-byte row per point in ONE plane: [Cinv(6) u0x u0y | u0z pad(3)] = three sectors, three 256-bit
// accesses.  (A split into a 48-byte Cinv plane + a u0 plane was measured: the gather of k_schur_cm then needs four L1
// requests per observation instead of three and ran 17 % slower.)  The PCG product reads the first two sectors only.
// `u0p` is kept in the signatures as an alias of the same plane (rows of three double4).
__device__ __forceinline__ void load_cinv(const double* __restrict__ cinv, const int j, double* Ci /*6*/) {
  const double4* q = reinterpret_cast<const double4*>(cinv + (size_t)PBLK * j);
  const double4 a = ldg4(q), b = ldg4(q + 1);
  Ci[0] = a.x; Ci[1] = a.y; Ci[2] = a.z; Ci[3] = a.w; Ci[4] = b.x; Ci[5] = b.y;
}
__device__ __forceinline__ void load_pblk(const double* __restrict__ cinv, const double4* __restrict__ /*u0p*/, const int j, double* Ci /*6*/,
                                          double* u0 /*3*/) {
  const double4* q = reinterpret_cast<const double4*>(cinv + (size_t)PBLK * j);
  const double4 a = ldg4(q), b = ldg4(q + 1), d = ldg4(q + 2);
  Ci[0] = a.x; Ci[1] = a.y; Ci[2] = a.z; Ci[3] = a.w; Ci[4] = b.x; Ci[5] = b.y; u0[0] = b.z; u0[1] = b.w; u0[2] = d.x;
}
__device__ __forceinline__ void store_pblk(double* __restrict__ cinv, double4* __restrict__ /*u0p*/, const int j, const double* blk /* PBLK */) {
  double4* q = reinterpret_cast<double4*>(cinv + (size_t)PBLK * j);
  st4(q, make_double4(blk[0], blk[1], blk[2], blk[3]));
  st4(q + 1, make_double4(blk[4], blk[5], blk[6], blk[7]));
  st4(q + 2, make_double4(blk[8], blk[9], blk[10], blk[11]));
}

// rho(s): returns 1/2-free rho and sqrt(rho') (Ceres corrector with rho'' <= 0: plain IRLS scaling)
__device__ __forceinline__ void loss_eval(const LossP L, const double s, double& rho, double& w) {
  if (L.kind == 2) {  // Cauchy
    const double b = L.a * L.a;
    const double sum = 1.0 + s / b;
    rho = b * log(sum);
    w = rsqrt(sum);
  } else if (L.kind == 1) {  // Huber
    const double b = L.a * L.a;
    if (s > b) {
      const double r = sqrt(s);
      rho = 2.0 * L.a * r - b;
      w = sqrt(L.a / r);
    } else { rho = s; w = 1.0; }
  } else { rho = s; w = 1.0; }
}

// Residual at (camera table row, point).  Returns the record and the un-weighted residual.
__device__ __forceinline__ void project_obs(const double* __restrict__ ct, const double X, const double Y,
                                            const double Z, const Intr K, const double u, const double v,
                                            double& xh, double& yh, double& iz, double& rx, double& ry) {
  // plain loads: ct may live in shared memory (pose-only kernel); global callers pass __restrict__ const
  const double qx = X - ct[CT_C], qy = Y - ct[CT_C + 1], qz = Z - ct[CT_C + 2];
  const double px = ct[0] * qx + ct[1] * qy + ct[2] * qz;
  const double py = ct[3] * qx + ct[4] * qy + ct[5] * qz;
  const double pz = ct[6] * qx + ct[7] * qy + ct[8] * qz;
  iz = 1.0 / pz;
  xh = px * iz; yh = py * iz;
  rx = K.fx * xh + K.cx - u;
  ry = K.fy * yh + K.cy - v;
}

// Rows of J^ = w [Q | -P] (2x6) from a record; sv = small-angle vector of the camera (usually 0).
__device__ __forceinline__ void jhat_rows(const double4 rec, const double sv0, const double sv1, const double sv2,
                                          const Intr K, double a[6], double b[6]) {
  const double xh = rec.x, yh = rec.y, iz = rec.z, w = rec.w;
  const double m0 = xh + (sv1 - sv2 * yh);
  const double m1 = yh + (sv2 * xh - sv0);
  const double m2 = 1.0 + (sv0 * yh - sv1 * xh);
  const double wfx = w * K.fx, wfy = w * K.fy;
  a[0] = wfx * (xh * m1); a[1] = -wfx * (m2 + xh * m0); a[2] = wfx * m1;
  b[0] = wfy * (m2 + yh * m1); b[1] = -wfy * (yh * m0); b[2] = -wfy * m0;
  const double pfx = wfx * iz, pfy = wfy * iz;
  a[3] = -pfx; a[4] = 0.0; a[5] = pfx * xh;
  b[3] = 0.0; b[4] = -pfy; b[5] = pfy * yh;
}

// J^ has two structural zeros (a[4] = b[3] = 0: a pixel row does not see the other axis' translation).  The camera-major
// kernels are FP64-pipe co-limited (ncu: 46 % pipe at 37 % DRAM), so the symmetric updates below skip those terms:
//   acc(r,c) += x_r a_c + y_r b_c   with (x,y) = (a,b) for J^'J^  or (E a-ish, E b-ish) for J^' E J^
// 31/30 FMAs instead of 42.  Upper-triangular packing order as tri(r,c).
// Two-term entries are written as two chained FMAs: `acc += x a + y b` compiles to DMUL + DFMA + DADD, three instructions of
// the FP64 pipe these kernels are co-limited by, instead of two.
__device__ __forceinline__ double fma2(const double x, const double a, const double y, const double b, const double acc) {
  return fma(x, a, fma(y, b, acc));
}
__device__ __forceinline__ void acc_sym_sparse(double* acc, const double* x, const double* y, const double* a, const double* b) {
  acc[0] = fma2(x[0], a[0], y[0], b[0], acc[0]);  acc[1] = fma2(x[0], a[1], y[0], b[1], acc[1]);  acc[2] = fma2(x[0], a[2], y[0], b[2], acc[2]);
  acc[3] = fma(x[0], a[3], acc[3]);               acc[4] = fma(y[0], b[4], acc[4]);               acc[5] = fma2(x[0], a[5], y[0], b[5], acc[5]);
  acc[6] = fma2(x[1], a[1], y[1], b[1], acc[6]);  acc[7] = fma2(x[1], a[2], y[1], b[2], acc[7]);  acc[8] = fma(x[1], a[3], acc[8]);
  acc[9] = fma(y[1], b[4], acc[9]);               acc[10] = fma2(x[1], a[5], y[1], b[5], acc[10]);
  acc[11] = fma2(x[2], a[2], y[2], b[2], acc[11]); acc[12] = fma(x[2], a[3], acc[12]);            acc[13] = fma(y[2], b[4], acc[13]);
  acc[14] = fma2(x[2], a[5], y[2], b[5], acc[14]);
  acc[15] = fma(x[3], a[3], acc[15]);             acc[16] = fma(y[3], b[4], acc[16]);             acc[17] = fma(x[5], a[3], acc[17]);   // (3,5) taken as (5,3)
  acc[18] = fma(y[4], b[4], acc[18]);             acc[19] = fma(y[5], b[4], acc[19]);                                                   // (4,5) taken as (5,4)
  acc[20] = fma2(x[5], a[5], y[5], b[5], acc[20]);
}
__device__ __forceinline__ void acc_vec_sparse(double* acc, const double* a, const double* b, const double f0, const double f1) {
  acc[0] = fma2(a[0], f0, b[0], f1, acc[0]); acc[1] = fma2(a[1], f0, b[1], f1, acc[1]); acc[2] = fma2(a[2], f0, b[2], f1, acc[2]);
  acc[3] = fma(a[3], f0, acc[3]);            acc[4] = fma(b[4], f1, acc[4]);            acc[5] = fma2(a[5], f0, b[5], f1, acc[5]);
}

// Rows of J~_p = w P R (2x3) from a record and R (row-major 3x3).
__device__ __forceinline__ void jp_rows(const double4 rec, const double* R, const Intr K, double ap[3], double bp[3]) {
  const double pfx = rec.w * rec.z * K.fx, pfy = rec.w * rec.z * K.fy;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    ap[c] = pfx * (R[c] - rec.x * R[6 + c]);
    bp[c] = pfy * (R[3 + c] - rec.y * R[6 + c]);
  }
}

// Symmetric 3x3 (00,01,02,11,12,22) positive definite inverse by LDL'.  Returns false if not PD.
__device__ __forceinline__ bool inv3_sym(const double* C, double* Ci) {
  const double d0 = C[0];
  if (!(d0 > 0.0)) return false;
  const double i0 = 1.0 / d0;
  const double l10 = C[1] * i0, l20 = C[2] * i0;
  const double d1 = C[3] - l10 * C[1];
  if (!(d1 > 0.0)) return false;
  const double i1 = 1.0 / d1;
  const double t21 = C[4] - l20 * C[1];
  const double l21 = t21 * i1;
  const double d2 = C[5] - l20 * C[2] - l21 * t21;
  if (!(d2 > 0.0)) return false;
  const double i2 = 1.0 / d2;
  // L^-1 = [[1,0,0],[-l10,1,0],[l10 l21 - l20, -l21, 1]];  Ci = L^-T D^-1 L^-1
  const double m20 = l10 * l21 - l20;
  Ci[0] = i0 + l10 * l10 * i1 + m20 * m20 * i2;
  Ci[1] = -l10 * i1 - m20 * l21 * i2;
  Ci[2] = m20 * i2;
  Ci[3] = i1 + l21 * l21 * i2;
  Ci[4] = -l21 * i2;
  Ci[5] = i2;
  return true;
}

// Dense 6x6 SPD inverse by Cholesky (row-major).  Returns false if not PD.
__device__ inline bool inv6_spd(const double* A, double* Ai) {
  double L[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
    for (int k = 0; k < j; ++k) d -= L[j * 6 + k] * L[j * 6 + k];
    if (!(d > 0.0)) return false;
    const double ljj = sqrt(d);
    L[j * 6 + j] = ljj;
    const double inv = 1.0 / ljj;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
      for (int k = 0; k < j; ++k) s -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = s * inv;
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

// Fixed-order CTA reduction: every value goes through the same xor-butterfly and the same
// warp-by-warp serial sum, so the result does not depend on scheduling.
template <int NV, int NT, bool IS_MAX = false>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* sm /* NV * NT/32 */, double* out /* NV */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double y = __shfl_xor_sync(0xffffffffu, x, o);
      x = IS_MAX ? fmax(x, y) : x + y;
    }
    if (lane == 0) sm[wid * NV + i] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = sm[threadIdx.x];
    for (int w = 1; w < NT / 32; ++w) x = IS_MAX ? fmax(x, sm[w * NV + threadIdx.x]) : x + sm[w * NV + threadIdx.x];
    out[threadIdx.x] = x;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Per-camera table: R = R(-w) exactly as ceres::AngleAxisRotatePoint rotates (Rodrigues for
// theta^2 > DBL_EPSILON, I + [v]x otherwise), G = J_r(w), centre, small-angle vector.
// ---------------------------------------------------------------------------------------------
__device__ inline void cam_table_row(const double* cam, double* ct) {
  const double w0 = cam[0], w1 = cam[1], w2 = cam[2];
  const double v0 = -w0, v1 = -w1, v2 = -w2;
  const double th2 = v0 * v0 + v1 * v1 + v2 * v2;
  double R[9], G[9], sv[3];
  if (th2 > DBL_EPSILON) {
    const double th = sqrt(th2);
    double s, c;
    sincos(th, &s, &c);
    const double ith = 1.0 / th;
    const double k0 = v0 * ith, k1 = v1 * ith, k2 = v2 * ith;
    const double oc = 1.0 - c;
    R[0] = c + oc * k0 * k0;      R[1] = oc * k0 * k1 - s * k2; R[2] = oc * k0 * k2 + s * k1;
    R[3] = oc * k1 * k0 + s * k2; R[4] = c + oc * k1 * k1;      R[5] = oc * k1 * k2 - s * k0;
    R[6] = oc * k2 * k0 - s * k1; R[7] = oc * k2 * k1 + s * k0; R[8] = c + oc * k2 * k2;
    double a, b;
    if (th < 1e-2) {
      a = 0.5 - th2 * (1.0 / 24.0) + th2 * th2 * (1.0 / 720.0);
      b = 1.0 / 6.0 - th2 * (1.0 / 120.0) + th2 * th2 * (1.0 / 5040.0);
    } else {
      a = oc / th2;
      b = (th - s) / (th2 * th);
    }
    // G = I - a [w]x + b [w]x^2,  [w]x^2 = w w' - |w|^2 I
    G[0] = 1.0 + b * (w0 * w0 - th2); G[1] = a * w2 + b * w0 * w1;      G[2] = -a * w1 + b * w0 * w2;
    G[3] = -a * w2 + b * w1 * w0;     G[4] = 1.0 + b * (w1 * w1 - th2); G[5] = a * w0 + b * w1 * w2;
    G[6] = a * w1 + b * w2 * w0;      G[7] = -a * w0 + b * w2 * w1;     G[8] = 1.0 + b * (w2 * w2 - th2);
    sv[0] = sv[1] = sv[2] = 0.0;
  } else {
    R[0] = 1.0; R[1] = -v2; R[2] = v1;
    R[3] = v2;  R[4] = 1.0; R[5] = -v0;
    R[6] = -v1; R[7] = v0;  R[8] = 1.0;
    G[0] = 1.0; G[1] = 0.0; G[2] = 0.0; G[3] = 0.0; G[4] = 1.0; G[5] = 0.0; G[6] = 0.0; G[7] = 0.0; G[8] = 1.0;
    sv[0] = w0; sv[1] = w1; sv[2] = w2;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) { ct[i] = R[i]; ct[CT_G + i] = G[i]; }
  ct[CT_C] = cam[3]; ct[CT_C + 1] = cam[4]; ct[CT_C + 2] = cam[5];
  ct[CT_SV] = sv[0]; ct[CT_SV + 1] = sv[1]; ct[CT_SV + 2] = sv[2];
}

// ---- GLBA_MODE_G2O (world-to-camera SE(3), T <- exp([dw, dv]) T) --------------------------------------------------
// The per-observation kernels only ever see the table row [R | c | G | sv] with p = R (X - c) and J~c = J^ blockdiag(G, R).
// For g2o's left perturbation, d(obs - proj)/d[dw, dv] = J^ exactly, and (H + lambda I) is invariant under the orthogonal
// change of variables  dw = x_w,  dv = -R x_t:  so the g2o system is solved in x with G = -I and the existing
// translation block R, and the update maps x back (k_cam_step2).  R = exp([w_cw]x), c = -R' t.
__device__ inline void so3_exp(const double* w, double* R, double* B_out) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double A, B;
  if (th2 < 1e-8) { A = 1.0 - th2 * (1.0 / 6.0) + th2 * th2 * (1.0 / 120.0); B = 0.5 - th2 * (1.0 / 24.0) + th2 * th2 * (1.0 / 720.0); }
  else { const double th = sqrt(th2); double sn, cs; sincos(th, &sn, &cs); A = sn / th; B = (1.0 - cs) / th2; }
  const double x = w[0], y = w[1], z = w[2];
  R[0] = 1.0 - B * (y * y + z * z); R[1] = -A * z + B * x * y;        R[2] = A * y + B * x * z;
  R[3] = A * z + B * x * y;         R[4] = 1.0 - B * (x * x + z * z); R[5] = -A * x + B * y * z;
  R[6] = -A * y + B * x * z;        R[7] = A * x + B * y * z;         R[8] = 1.0 - B * (x * x + y * y);
  if (B_out) *B_out = B;
}
__device__ inline void so3_log(const double* R, double* w) {
  double q0, q1, q2, q3;
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0.0) { const double s = sqrt(tr + 1.0) * 2.0; q0 = 0.25 * s; q1 = (R[7] - R[5]) / s; q2 = (R[2] - R[6]) / s; q3 = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2.0; q0 = (R[7] - R[5]) / s; q1 = 0.25 * s; q2 = (R[1] + R[3]) / s; q3 = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2.0; q0 = (R[2] - R[6]) / s; q1 = (R[1] + R[3]) / s; q2 = 0.25 * s; q3 = (R[5] + R[7]) / s; }
  else { const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2.0; q0 = (R[3] - R[1]) / s; q1 = (R[2] + R[6]) / s; q2 = (R[5] + R[7]) / s; q3 = 0.25 * s; }
  if (q0 < 0.0) { q0 = -q0; q1 = -q1; q2 = -q2; q3 = -q3; }
  const double n = sqrt(q1 * q1 + q2 * q2 + q3 * q3);
  const double k = (n < 1e-10) ? 2.0 / q0 : 2.0 * atan2(n, q0) / n;
  w[0] = k * q1; w[1] = k * q2; w[2] = k * q3;
}
__device__ inline void cam_table_row_Rt(const double* R, const double* t, double* ct) {
#pragma unroll
  for (int i = 0; i < 9; ++i) { ct[i] = R[i]; ct[CT_G + i] = (i % 4 == 0) ? -1.0 : 0.0; }
#pragma unroll
  for (int r = 0; r < 3; ++r) ct[CT_C + r] = -(R[r] * t[0] + R[3 + r] * t[1] + R[6 + r] * t[2]);
  ct[CT_SV] = ct[CT_SV + 1] = ct[CT_SV + 2] = 0.0;
}
__device__ inline void cam_table_row_g2o(const double* cam, double* ct) {
  double R[9];
  so3_exp(cam, R, nullptr);
  cam_table_row_Rt(R, cam + 3, ct);
}

__global__ void k_cam_prep(const int n_cam, const double* __restrict__ cam, double* __restrict__ camtab, const int mode) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cam) return;
  if (mode) cam_table_row_g2o(cam + 6 * i, camtab + (size_t)CAMTAB * i);
  else cam_table_row(cam + 6 * i, camtab + (size_t)CAMTAB * i);
}

// GLBA_MODE_G2O, computeLambdaInit: max over the free vertices of the Hessian diagonal in g2o's own coordinates
// (points: diag C_j; cameras: rotation part diag B_ww, translation part diag(R B_tt R')).  Non-negative doubles order
// like their bit patterns, so an integer atomicMax is exact and order-independent.
__global__ void k_hmax(const int n_pt, const uint8_t* __restrict__ pt_free, const double* __restrict__ Craw, const int n_cam,
                       const uint8_t* __restrict__ cam_free, const double* __restrict__ Bc, const double* __restrict__ camtab,
                       unsigned long long* __restrict__ out) {
  pdl_grid_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  double m = 0.0;
  if (t < n_pt) {
    if (pt_free[t]) m = fmax(Craw[t], fmax(Craw[(size_t)3 * n_pt + t], Craw[(size_t)5 * n_pt + t]));
  } else if (t < n_pt + n_cam) {
    const int i = t - n_pt;
    if (cam_free[i]) {
      const double* B = Bc + (size_t)36 * i;
      const double* R = camtab + (size_t)CAMTAB * i;
      m = fmax(B[0], fmax(B[7], B[14]));
      for (int r = 0; r < 3; ++r) {
        double d = 0.0;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) d += R[r * 3 + a] * B[(3 + a) * 6 + 3 + b] * R[r * 3 + b];
        m = fmax(m, d);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// ---------------------------------------------------------------------------------------------
// Damped point block: Cinv = (C + lam/radius)^-1, u0 = Cinv g.  lam = clamp(s^2 h)/s^2 per axis.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool point_block(const double* C6, const double* g3, const double* lam3, const double inv_radius,
                                            double* blk /* PBLK */) {
  double Cd[6] = {C6[0] + lam3[0] * inv_radius, C6[1], C6[2], C6[3] + lam3[1] * inv_radius, C6[4],
                  C6[5] + lam3[2] * inv_radius};
  double Ci[6];
  const bool ok = inv3_sym(Cd, Ci);
  if (!ok) {
#pragma unroll
    for (int i = 0; i < 6; ++i) Ci[i] = 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) blk[i] = Ci[i];
  blk[6] = Ci[0] * g3[0] + Ci[1] * g3[1] + Ci[2] * g3[2];
  blk[7] = Ci[1] * g3[0] + Ci[3] * g3[1] + Ci[4] * g3[2];
  blk[8] = Ci[2] * g3[0] + Ci[4] * g3[1] + Ci[5] * g3[2];
  blk[9] = blk[10] = blk[11] = 0.0;
  return ok;
}

struct PmArgs {
  int n_pt;
  const int* pt_start;          // n_pt+1 (point-major CSR)
  const int* pm_cam;            // camera of observation k
  const double2* pm_uv;         // measurement
  const int* pm2cm;             // position of observation k in camera-major order
  const uint8_t* pt_free;
  Intr K;
  LossP loss;
};

// K_A + point half of K_B.  One thread per point walks its track: residual, robust weight, record
// (written in both orders), C_j = sum J~p'J~p, g_j = sum J~p' r~, then the damped inverse.
// first != 0: this is iteration 0, also fix the Jacobi scaling s = 1/(1+sqrt(h)).
__global__ void __launch_bounds__(NT_PM)
k_linearize_pm(const PmArgs A, const double4* __restrict__ pt, const double* __restrict__ camtab,
               double4* __restrict__ rec_pm, double4* __restrict__ rec_cm, double* __restrict__ Craw /* 9 SoA */,
               double4* __restrict__ sp4, double4* __restrict__ lam4, double* __restrict__ cinv, double4* __restrict__ u0p, const int first,
               const int jacobi, const double min_diag, const double max_diag, const double inv_radius,
               double* __restrict__ part /* [grid][5] */) {
  pdl_grid_sync();
  __shared__ double sm[5 * NT_PM / 32];
  __shared__ double smo[5];
  const int j = blockIdx.x * NT_PM + threadIdx.x;
  double cost = 0.0, xn2 = 0.0, gmax = 0.0, bad = 0.0, notpd = 0.0;
  if (j < A.n_pt) {
    const int b = __ldg(A.pt_start + j), e = __ldg(A.pt_start + j + 1);
    const double4 X = ldg4(pt + j);
    double C[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (int k = b; k < e; ++k) {
      const int i = __ldg(A.pm_cam + k);
      const double2 uv = __ldg(A.pm_uv + k);
      const double* ct = camtab + (size_t)CAMTAB * i;
      double xh, yh, iz, rx, ry;
      project_obs(ct, X.x, X.y, X.z, A.K, uv.x, uv.y, xh, yh, iz, rx, ry);
      double rho, w;
      loss_eval(A.loss, (X.w * X.w) * (rx * rx + ry * ry), rho, w);
      w *= X.w;
      if (!isfinite(rx) || !isfinite(ry)) bad += 1.0;
      cost += 0.5 * rho;
      const double4 rec = make_double4(xh, yh, iz, w);
      st4(rec_pm + k, rec);
      st4(rec_cm + __ldg(A.pm2cm + k), rec);
      double R[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) R[q] = __ldg(ct + q);
      double ap[3], bp[3];
      jp_rows(rec, R, A.K, ap, bp);
      const double r0 = w * rx, r1 = w * ry;
      C[0] += ap[0] * ap[0] + bp[0] * bp[0]; C[1] += ap[0] * ap[1] + bp[0] * bp[1]; C[2] += ap[0] * ap[2] + bp[0] * bp[2];
      C[3] += ap[1] * ap[1] + bp[1] * bp[1]; C[4] += ap[1] * ap[2] + bp[1] * bp[2]; C[5] += ap[2] * ap[2] + bp[2] * bp[2];
      g[0] += ap[0] * r0 + bp[0] * r1; g[1] += ap[1] * r0 + bp[1] * r1; g[2] += ap[2] * r0 + bp[2] * r1;
    }
    const bool free_pt = A.pt_free[j] != 0;
#pragma unroll
    for (int q = 0; q < 6; ++q) Craw[(size_t)q * A.n_pt + j] = C[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) Craw[(size_t)(6 + q) * A.n_pt + j] = g[q];
    double blk[PBLK];
    if (free_pt) {
      const double h[3] = {C[0], C[3], C[5]};
      double s[3], lam[3];
      if (first) {
#pragma unroll
        for (int q = 0; q < 3; ++q) s[q] = jacobi ? 1.0 / (1.0 + sqrt(h[q])) : 1.0;
        st4(sp4 + j, make_double4(s[0], s[1], s[2], 0.0));
      } else {
        const double4 s4 = ldg4(sp4 + j);
        s[0] = s4.x; s[1] = s4.y; s[2] = s4.z;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double s2 = s[q] * s[q];
        lam[q] = fmin(fmax(s2 * h[q], min_diag), max_diag) / s2;
      }
      st4(lam4 + j, make_double4(lam[0], lam[1], lam[2], 0.0));
      if (!point_block(C, g, lam, inv_radius, blk)) notpd += 1.0;
      xn2 = X.x * X.x + X.y * X.y + X.z * X.z;
      gmax = fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2])));
    } else {
#pragma unroll
      for (int q = 0; q < PBLK; ++q) blk[q] = 0.0;
      if (first) st4(sp4 + j, make_double4(1.0, 1.0, 1.0, 0.0));
      st4(lam4 + j, make_double4(0.0, 0.0, 0.0, 0.0));
    }
    store_pblk(cinv, u0p, j, blk);
  }
  double v[4] = {cost, xn2, bad, notpd};
  block_reduce<4, NT_PM>(v, sm, smo);
  double m[1] = {gmax};
  __shared__ double smm[NT_PM / 32];
  __shared__ double smmo[1];
  block_reduce<1, NT_PM, true>(m, smm, smmo);
  if (threadIdx.x == 0) {
    double* p = part + (size_t)5 * blockIdx.x;
    p[0] = smo[0]; p[1] = smo[1]; p[2] = smo[2]; p[3] = smo[3]; p[4] = smmo[0];
  }
}

// Re-damp the point blocks after the radius changed (rejected step): same C, g, lam.
__global__ void __launch_bounds__(NT_PM)
k_point_damp(const int n_pt, const uint8_t* __restrict__ pt_free, const double* __restrict__ Craw,
             const double4* __restrict__ lam4, double* __restrict__ cinv, double4* __restrict__ u0p, const double inv_radius_arg,
             double* __restrict__ part /* [grid][1] notpd */, const LmCtl* __restrict__ ctl = nullptr) {
  pdl_grid_sync();
  __shared__ double sm[NT_PM / 32];
  __shared__ double smo[1];
  if (ctl_skip(ctl, GATE_REDAMP)) return;
  const double inv_radius = ctl_inv_radius(ctl, inv_radius_arg);
  const int j = blockIdx.x * NT_PM + threadIdx.x;
  double notpd = 0.0;
  if (j < n_pt && pt_free[j]) {
    double C[6], g[3];
#pragma unroll
    for (int q = 0; q < 6; ++q) C[q] = Craw[(size_t)q * n_pt + j];
#pragma unroll
    for (int q = 0; q < 3; ++q) g[q] = Craw[(size_t)(6 + q) * n_pt + j];
    const double4 l4 = ldg4(lam4 + j);
    const double lam[3] = {l4.x, l4.y, l4.z};
    double blk[PBLK];
    if (!point_block(C, g, lam, inv_radius, blk)) notpd = 1.0;
    store_pblk(cinv, u0p, j, blk);
  }
  double v[1] = {notpd};
  block_reduce<1, NT_PM>(v, sm, smo);
  if (threadIdx.x == 0) part[blockIdx.x] = smo[0];
}

// ---------------------------------------------------------------------------------------------
// Camera-major passes.  One CTA per chunk = (camera, [begin,end)) of the camera-major order.
// ---------------------------------------------------------------------------------------------
struct CmArgs {
  const int* chunk_cam;
  const int* chunk_begin;
  const int* chunk_end;
  const int* cm_pt;
  const double2* cm_uv;
  const uint8_t* cam_free;
  Intr K;
};

// Camera half of K_B: A_i = sum J^'J^ (21 upper entries), ghat_i = sum J^' r~ (6), per chunk.
// 5 CTAs/SM (96 registers, a few spilled accumulators) beat 4 (126 registers): measured 0.067 vs 0.076 ms on C4
__global__ void __launch_bounds__(NT_HCM, 5)
k_linearize_cm(const CmArgs A, const double4* __restrict__ rec_cm, const double* __restrict__ camtab,
               double* __restrict__ part /* [n_chunks][27] */, const LmCtl* __restrict__ ctl = nullptr) {
  pdl_grid_sync();
  __shared__ double sm[27 * NT_HCM / 32];
  __shared__ double smo[27];
  if (ctl_skip(ctl, GATE_ACCEPTED)) return;
  const int ch = blockIdx.x;
  const int cam = A.chunk_cam[ch];
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0.0;
  if (A.cam_free[cam]) {   // uniform per CTA
    const double* ct = camtab + (size_t)CAMTAB * cam;
    const double sv0 = ct[21], sv1 = ct[22], sv2 = ct[23];
    const int b = A.chunk_begin[ch], e = A.chunk_end[ch];
    // two observations per trip: both records / measurements are requested before either is consumed
    for (int k = b + threadIdx.x; k < e; k += 2 * NT_HCM) {
      const int k2 = k + NT_HCM;
      const bool has2 = k2 < e;
      const double4 rec = ldg4(rec_cm + k);
      const double2 uv = __ldg(A.cm_uv + k);
      const double4 recb = ldg4(rec_cm + (has2 ? k2 : k));
      const double2 uvb = __ldg(A.cm_uv + (has2 ? k2 : k));
      {
        double a[6], bb[6];
        jhat_rows(rec, sv0, sv1, sv2, A.K, a, bb);
        const double r0 = rec.w * (A.K.fx * rec.x + A.K.cx - uv.x);
        const double r1 = rec.w * (A.K.fy * rec.y + A.K.cy - uv.y);
        acc_sym_sparse(acc, a, bb, a, bb);
        acc_vec_sparse(acc + 21, a, bb, r0, r1);
      }
      if (has2) {
        double a[6], bb[6];
        jhat_rows(recb, sv0, sv1, sv2, A.K, a, bb);
        const double r0 = recb.w * (A.K.fx * recb.x + A.K.cx - uvb.x);
        const double r1 = recb.w * (A.K.fy * recb.y + A.K.cy - uvb.y);
        acc_sym_sparse(acc, a, bb, a, bb);
        acc_vec_sparse(acc + 21, a, bb, r0, r1);
      }
    }
  }
  block_reduce<27, NT_HCM>(acc, sm, smo);
  if (threadIdx.x < 27) part[(size_t)27 * ch + threadIdx.x] = smo[threadIdx.x];
}

// Schur half: for every observation of the camera, E = J~p Cinv J~p' (2x2) and f = J~p u0 (2):
//   Mhat_i = sum J^' E J^ (21),  rhat_i = sum J^' f (6).
// here the spills cost more than the occupancy gains: 4 CTAs/SM
__global__ void __launch_bounds__(NT_HCM, 4)
k_schur_cm(const CmArgs A, const double4* __restrict__ rec_cm, const double* __restrict__ camtab,
           const double* __restrict__ cinv, const double4* __restrict__ u0p, double* __restrict__ part /* [n_chunks][27] */) {
  pdl_grid_sync();
  __shared__ double sm[27 * NT_HCM / 32];
  __shared__ double smo[27];
  const int ch = blockIdx.x;
  const int cam = A.chunk_cam[ch];
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0.0;
  if (A.cam_free[cam]) {
    const double* ct = camtab + (size_t)CAMTAB * cam;
    double R[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = ct[q];
    const double sv0 = ct[21], sv1 = ct[22], sv2 = ct[23];
    const int b = A.chunk_begin[ch], e = A.chunk_end[ch];
    for (int k = b + threadIdx.x; k < e; k += NT_HCM) {
      const double4 rec = ldg4(rec_cm + k);
      const int j = __ldg(A.cm_pt + k);
      double Ci[6], u0[3];
      load_pblk(cinv, u0p, j, Ci, u0);
      double ap[3], bp[3];
      jp_rows(rec, R, A.K, ap, bp);
      // t = Cinv ap, s = Cinv bp   (Cinv packed 00,01,02,11,12,22)
      const double ta0 = Ci[0] * ap[0] + Ci[1] * ap[1] + Ci[2] * ap[2];
      const double ta1 = Ci[1] * ap[0] + Ci[3] * ap[1] + Ci[4] * ap[2];
      const double ta2 = Ci[2] * ap[0] + Ci[4] * ap[1] + Ci[5] * ap[2];
      const double tb0 = Ci[0] * bp[0] + Ci[1] * bp[1] + Ci[2] * bp[2];
      const double tb1 = Ci[1] * bp[0] + Ci[3] * bp[1] + Ci[4] * bp[2];
      const double tb2 = Ci[2] * bp[0] + Ci[4] * bp[1] + Ci[5] * bp[2];
      const double E00 = ap[0] * ta0 + ap[1] * ta1 + ap[2] * ta2;
      const double E01 = ap[0] * tb0 + ap[1] * tb1 + ap[2] * tb2;
      const double E11 = bp[0] * tb0 + bp[1] * tb1 + bp[2] * tb2;
      const double f0 = ap[0] * u0[0] + ap[1] * u0[1] + ap[2] * u0[2];
      const double f1 = bp[0] * u0[0] + bp[1] * u0[1] + bp[2] * u0[2];
      double a[6], bb[6];
      jhat_rows(rec, sv0, sv1, sv2, A.K, a, bb);
      // J^' E J^ = (E00 a + E01 b) a' + (E01 a + E11 b) b'
      double ea[6], eb[6];
#pragma unroll
      for (int r = 0; r < 3; ++r) { ea[r] = E00 * a[r] + E01 * bb[r]; eb[r] = E01 * a[r] + E11 * bb[r]; }
      ea[3] = E00 * a[3]; eb[3] = E01 * a[3];
      ea[4] = E01 * bb[4]; eb[4] = E11 * bb[4];
      ea[5] = E00 * a[5] + E01 * bb[5]; eb[5] = E01 * a[5] + E11 * bb[5];
      acc_sym_sparse(acc, ea, eb, a, bb);
      acc_vec_sparse(acc + 21, a, bb, f0, f1);
    }
  }
  block_reduce<27, NT_HCM>(acc, sm, smo);
  if (threadIdx.x < 27) part[(size_t)27 * ch + threadIdx.x] = smo[threadIdx.x];
}

// Second half of the implicit product: yhat_i = sum_j J^' (J~p u_j), u_j from the point-major pass.
__global__ void __launch_bounds__(NT_CM)
k_spmv_cm(const CmArgs A, const double4* __restrict__ rec_cm, const double* __restrict__ camtab,
          const double4* __restrict__ u4, const CgState* __restrict__ cg, const int li, double* __restrict__ part /* [n_chunks][6] */) {
  pdl_grid_sync();
  __shared__ double sm[6 * NT_CM / 32];
  __shared__ double smo[6];
  if (cg && cg->done_at <= li) return;
  const int ch = blockIdx.x;
  const int cam = A.chunk_cam[ch];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  if (A.cam_free[cam]) {
    const double* ct = camtab + (size_t)CAMTAB * cam;
    double R[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = ct[q];
    const double sv0 = ct[21], sv1 = ct[22], sv2 = ct[23];
    const int b = A.chunk_begin[ch], e = A.chunk_end[ch];
    for (int k = b + threadIdx.x; k < e; k += 2 * NT_CM) {
      const int k2 = k + NT_CM;
      const bool has2 = k2 < e;
      const double4 rec = ldg4(rec_cm + k);
      const int j = __ldg(A.cm_pt + k);
      const double4 recb = ldg4(rec_cm + (has2 ? k2 : k));
      const int jb = __ldg(A.cm_pt + (has2 ? k2 : k));
      const double4 u = ldg4(u4 + j);
      const double4 ub = ldg4(u4 + jb);
      {
        double ap[3], bp[3], a[6], bb[6];
        jp_rows(rec, R, A.K, ap, bp);
        const double f0 = ap[0] * u.x + ap[1] * u.y + ap[2] * u.z;
        const double f1 = bp[0] * u.x + bp[1] * u.y + bp[2] * u.z;
        jhat_rows(rec, sv0, sv1, sv2, A.K, a, bb);
        acc_vec_sparse(acc, a, bb, f0, f1);
      }
      if (has2) {
        double ap[3], bp[3], a[6], bb[6];
        jp_rows(recb, R, A.K, ap, bp);
        const double f0 = ap[0] * ub.x + ap[1] * ub.y + ap[2] * ub.z;
        const double f1 = bp[0] * ub.x + bp[1] * ub.y + bp[2] * ub.z;
        jhat_rows(recb, sv0, sv1, sv2, A.K, a, bb);
        acc_vec_sparse(acc, a, bb, f0, f1);
      }
    }
  }
  block_reduce<6, NT_CM>(acc, sm, smo);
  if (threadIdx.x < 6) part[(size_t)6 * ch + threadIdx.x] = smo[threadIdx.x];
}

// Sum the chunk partials of every camera in chunk order (fixed): acc[cam][NV].
template <int NV>
__global__ void k_chunk_sum(const int n_cam, const int* __restrict__ cam_chunk_start, const double* __restrict__ part,
                            double* __restrict__ acc, const CgState* __restrict__ cg, const int li) {
  pdl_grid_sync();
  if (cg && cg->done_at <= li) return;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cam * NV) return;
  const int cam = t / NV, q = t - cam * NV;
  double s = 0.0;
  for (int ch = cam_chunk_start[cam]; ch < cam_chunk_start[cam + 1]; ++ch) s += part[(size_t)NV * ch + q];
  acc[t] = s;
}

// ---------------------------------------------------------------------------------------------
// Point-major half of the implicit product / back-substitution.
//   t_j = sum_i J~p' (J~c x_i) = sum_i R_i' (w P)' (J^ . xg_i),   xg_i = T_i x_i  (per-camera table)
//   MODE 0 (SpMV):    u_j = Cinv_j t_j
//   MODE 1 (backsub): y_j = u0_j - Cinv_j t_j;  X+ = X - y_j;  then candidate cost over the track.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(NT_PM)
k_point_pass(const PmArgs A, const double4* __restrict__ rec_pm, const double* __restrict__ camtab,
             const double* __restrict__ xtab /* [n_cam][XTAB] */, const double* __restrict__ cinv, const double4* __restrict__ u0p,
             double4* __restrict__ u4, const CgState* __restrict__ cg, const int li,
             // MODE 1 only:
             const double4* __restrict__ pt, double4* __restrict__ pt_c, const double* __restrict__ camtab_c,
             const double* __restrict__ Craw, const double4* __restrict__ lam4, const double inv_radius,
             double* __restrict__ part /* [grid][5] */) {
  pdl_grid_sync();
  if (MODE == 0 && cg && cg->done_at <= li) return;
  const int j = blockIdx.x * NT_PM + threadIdx.x;
  double cost_c = 0.0, yn2 = 0.0, yg = 0.0, yly = 0.0, bad = 0.0;
  if (j < A.n_pt) {
    const int b = __ldg(A.pt_start + j), e = __ldg(A.pt_start + j + 1);
    const bool free_pt = A.pt_free[j] != 0;
    double t0 = 0, t1 = 0, t2 = 0;
    if (free_pt) {
      for (int k = b; k < e; ++k) {
        const int i = __ldg(A.pm_cam + k);
        const double4 rec = ldg4(rec_pm + k);
        const double* ct = camtab + (size_t)CAMTAB * i;
        const double* x = xtab + (size_t)XTAB * i;
        double a[6], bb[6];
        jhat_rows(rec, __ldg(ct + 21), __ldg(ct + 22), __ldg(ct + 23), A.K, a, bb);
        double al0 = 0, al1 = 0;
#pragma unroll
        for (int r = 0; r < 6; ++r) { const double xv = __ldg(x + r); al0 += a[r] * xv; al1 += bb[r] * xv; }
        double R[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) R[q] = __ldg(ct + q);
        double ap[3], bp[3];
        jp_rows(rec, R, A.K, ap, bp);
        t0 += ap[0] * al0 + bp[0] * al1; t1 += ap[1] * al0 + bp[1] * al1; t2 += ap[2] * al0 + bp[2] * al1;
      }
    }
    double Ci[6];
    load_cinv(cinv, j, Ci);
    const double v0 = Ci[0] * t0 + Ci[1] * t1 + Ci[2] * t2;
    const double v1 = Ci[1] * t0 + Ci[3] * t1 + Ci[4] * t2;
    const double v2 = Ci[2] * t0 + Ci[4] * t1 + Ci[5] * t2;
    if (MODE == 0) {
      st4(u4 + j, make_double4(v0, v1, v2, 0.0));
    } else {
      double Cj[6], u0[3];
      load_pblk(cinv, u0p, j, Cj, u0);
      const double y0 = u0[0] - v0, y1 = u0[1] - v1, y2 = u0[2] - v2;
      const double4 X = ldg4(pt + j);
      const double4 Xc = make_double4(X.x - y0, X.y - y1, X.z - y2, X.w);
      st4(pt_c + j, Xc);
      if (free_pt) {
        const double4 l4 = ldg4(lam4 + j);
        yn2 = y0 * y0 + y1 * y1 + y2 * y2;
        yg = y0 * Craw[(size_t)6 * A.n_pt + j] + y1 * Craw[(size_t)7 * A.n_pt + j] + y2 * Craw[(size_t)8 * A.n_pt + j];
        yly = (l4.x * y0 * y0 + l4.y * y1 * y1 + l4.z * y2 * y2) * inv_radius;
      }
      for (int k = b; k < e; ++k) {
        const int i = __ldg(A.pm_cam + k);
        const double2 uv = __ldg(A.pm_uv + k);
        double xh, yh, iz, rx, ry;
        project_obs(camtab_c + (size_t)CAMTAB * i, Xc.x, Xc.y, Xc.z, A.K, uv.x, uv.y, xh, yh, iz, rx, ry);
        double rho, w;
        loss_eval(A.loss, (Xc.w * Xc.w) * (rx * rx + ry * ry), rho, w);
        if (!isfinite(rx) || !isfinite(ry)) bad += 1.0;
        cost_c += 0.5 * rho;
      }
    }
  }
  if (MODE == 1) {
    __shared__ double sm[5 * NT_PM / 32];
    __shared__ double smo[5];
    double v[5] = {cost_c, yn2, yg, yly, bad};
    block_reduce<5, NT_PM>(v, sm, smo);
    if (threadIdx.x < 5) part[(size_t)5 * blockIdx.x + threadIdx.x] = smo[threadIdx.x];
  }
}

// Cost only at (camtab, pt): used for fixed-cost checks and by glba_cull (no records written).
// ---------------------------------------------------------------------------------------------
// Final fixed-order reduction of per-CTA partials: out[slot[q]] = sum/max over rows of part[:, q].
// ---------------------------------------------------------------------------------------------
struct ReduceMap { int n; int slot[8]; int is_max[8]; };

__global__ void __launch_bounds__(NT_CAM)
k_reduce_partials(const int rows, const int nv, const double* __restrict__ part, const ReduceMap M, double* __restrict__ scal,
                  const LmCtl* __restrict__ ctl = nullptr, const int gate = GATE_ALWAYS) {
  pdl_grid_sync();
  __shared__ double sm[NT_CAM / 32];
  __shared__ double smo[1];
  if (ctl_skip(ctl, gate)) return;
  for (int q = 0; q < nv; ++q) {
    const bool is_max = M.is_max[q] != 0;
    double acc = 0.0;   // all reduced quantities are >= 0, so 0 is neutral for max as well
    for (int r = threadIdx.x; r < rows; r += NT_CAM) {
      const double x = part[(size_t)nv * r + q];
      acc = is_max ? fmax(acc, x) : acc + x;
    }
    double v[1] = {acc};
    if (is_max) block_reduce<1, NT_CAM, true>(v, sm, smo); else block_reduce<1, NT_CAM, false>(v, sm, smo);
    if (threadIdx.x == 0) scal[M.slot[q]] = smo[0];
    __syncthreads();
  }
}

// Sharded runs, one launch in front of the fused all-reduce of a linearisation: per-camera sums of the chunk partials of
// k_linearize_cm (partA -> accA) and k_schur_cm (partB -> accB, may be null), and this rank's max |g_point| into its
// own slot of the payload (a sum over ranks of one non-zero per slot; the host takes the max over the slots).
__global__ void k_chunk_sum_lin(const int n_cam, const int* __restrict__ cam_chunk_start, const double* __restrict__ partA,
                                const double* __restrict__ partB, double* __restrict__ accA, double* __restrict__ accB,
                                const int rank, double* __restrict__ scal) {
  pdl_grid_sync();
  if (blockIdx.x == 0 && threadIdx.x < MAX_WORLD) scal[S_GSLOT0 + threadIdx.x] = ((int)threadIdx.x == rank) ? scal[S_GMAX_P] : 0.0;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_cam * 27;
  const double* part = partA;
  double* acc = accA;
  if (t >= n) { if (!partB || t >= 2 * n) return; t -= n; part = partB; acc = accB; }
  const int cam = t / 27, q = t - cam * 27;
  double s = 0.0;
  for (int ch = cam_chunk_start[cam]; ch < cam_chunk_start[cam + 1]; ++ch) s += part[(size_t)27 * ch + q];
  acc[t] = s;
}

// ---------------------------------------------------------------------------------------------
// Sharded runs, "owner-computes" layout (glba.cu, load_problem): a rank's problem holds only the cameras its own tracks observe.
// mask[i] = OR over ranks of (1 << rank) for the ranks that observe camera i (built by a sum all-reduce of disjoint bits).
// ---------------------------------------------------------------------------------------------
__global__ void k_act_mask(const long n, const int* __restrict__ obs_cam, const int bit, int* __restrict__ mask) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) mask[obs_cam[k]] = bit;           // same value from every observation of the camera
}
__global__ void k_own_flags(const int n_cam, const int* __restrict__ mask, const int rank, int* __restrict__ f_act, int* __restrict__ f_sh) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_cam) return;
  const int m = (i < n_cam) ? mask[i] : 0;
  f_act[i] = (m >> rank) & 1;
  f_sh[i] = (__popc(m) >= 2) ? 1 : 0;
}
// local camera a <-> global camera i; owner = lowest observing rank; shared slot (global order) or -1
__global__ void k_own_build(const int n_cam, const int* __restrict__ mask, const int rank, const int* __restrict__ pos_act,
                            const int* __restrict__ pos_sh, const uint8_t* __restrict__ fixed, int* __restrict__ l2g, uint8_t* __restrict__ owned,
                            int* __restrict__ shared, int* __restrict__ n_free_owned) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cam) return;
  const int m = mask[i];
  if (!((m >> rank) & 1)) return;
  const int a = pos_act[i];
  l2g[a] = i;
  const bool own = (__ffs(m) - 1) == rank;
  owned[a] = own ? 1 : 0;
  shared[a] = (__popc(m) >= 2) ? pos_sh[i] : -1;
  if (own && !(fixed && fixed[i])) atomicAdd(n_free_owned, 1);
}
__global__ void k_relabel_cam(const long n, const int* __restrict__ obs_cam, const int* __restrict__ pos_act, int* __restrict__ out) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = pos_act[obs_cam[k]];
}
__global__ void k_gather_cam(const int n_act, const int* __restrict__ l2g, const double* __restrict__ cam_g, const uint8_t* __restrict__ fix_g,
                             double* __restrict__ cam_l, uint8_t* __restrict__ fix_l) {
  pdl_grid_sync();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_act) return;
  const size_t i = l2g[a];
#pragma unroll
  for (int q = 0; q < 6; ++q) cam_l[6 * (size_t)a + q] = cam_g[6 * i + q];
  if (fix_l) fix_l[a] = fix_g ? fix_g[i] : 0;
}
// rows of `width` doubles per local camera -> rows of the global-camera-sized array `out` (stride out_stride); other rows untouched
__global__ void k_scatter_rows(const int n_act, const int* __restrict__ l2g, const uint8_t* __restrict__ owned, const int only_owned,
                               const double* __restrict__ in, const int width, double* __restrict__ out) {
  pdl_grid_sync();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)n_act * width) return;
  const int a = (int)(t / width), q = (int)(t - (long)a * width);
  if (only_owned && !owned[a]) return;
  out[(size_t)l2g[a] * width + q] = in[t];
}
// exchange buffer of a linearisation: row s of a shared camera = its 27 (+27) partial sums; the tail carries the point scalars
__global__ void k_xch_pack(const int n_act, const int* __restrict__ shared, const double* __restrict__ accA, const double* __restrict__ accB,
                           const int n_shared, double* __restrict__ xsend, const double* __restrict__ scal, const int n_tail) {
  pdl_grid_sync();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tail) xsend[(size_t)54 * n_shared + t] = scal[t];
  if (t >= (long)n_act * 54) return;
  const int a = (int)(t / 54), q = (int)(t - (long)a * 54);
  const int sidx = shared[a];
  if (sidx < 0) return;
  xsend[(size_t)54 * sidx + q] = (q < 27) ? accA[(size_t)27 * a + q] : (accB ? accB[(size_t)27 * a + q - 27] : 0.0);
}
__global__ void k_xch_unpack(const int n_act, const int* __restrict__ shared, const double* __restrict__ xrecv, const int n_shared,
                             double* __restrict__ accA, double* __restrict__ accB, double* __restrict__ scal, const int n_tail) {
  pdl_grid_sync();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_tail) scal[t] = xrecv[(size_t)54 * n_shared + t];
  if (t >= (long)n_act * 54) return;
  const int a = (int)(t / 54), q = (int)(t - (long)a * 54);
  const int sidx = shared[a];
  if (sidx < 0) return;
  const double v = xrecv[(size_t)54 * sidx + q];
  if (q < 27) { if (accA) accA[(size_t)27 * a + q] = v; } else if (accB) accB[(size_t)27 * a + q - 27] = v;
}
// ---------------------------------------------------------------------------------------------
// One-shot all-reduce of the compact exchange buffer over NVLink peer memory (glba.cu, "peer exchange").  The payload is
// latency-bound (C4 on 8 GPUs: 148 KB), so instead of a library collective every rank packs into a buffer its peers can
// read (CUDA IPC mapping), publishes a sequence number, and the unpack kernel of every rank waits for all sequence numbers,
// reads the W copies straight from the peers' memory and adds them in rank order (identical bits on every rank).
// Two slots alternate by sequence parity: a rank can only reach exchange q+2 after its unpack of q+1 has seen every peer's
// pack of q+1, which follows that peer's unpack of q in stream order, so nobody still reads the slot being rewritten.
// Waits give up after two minutes and raise *err (the host turns it into GLBA_E_NCCL): a lost peer cannot hang the device.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_relaxed_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// PUSH form: NVLink stores are posted, loads are round trips, so the packing kernel writes this rank's payload INTO every
// peer's buffer (area [slot][source rank]) and then its sequence number into every peer's flag [slot][source rank]; the
// unpack kernel polls and reads local memory only.  (The pull form — peers read the packer's buffer — measured 31 us per
// exchange on 2 GPUs against 15 us for ncclAllReduce.)
constexpr int P2P_MAX_WORLD = 16;
struct PeerTable { double* area[P2P_MAX_WORLD]; unsigned* flag[P2P_MAX_WORLD]; };      // per destination rank: where THIS rank's copy / flag goes
__global__ void k_xch_inverse(const int n_act, const int* __restrict__ shared, int* __restrict__ sh2loc) {
  pdl_grid_sync();
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a < n_act && shared[a] >= 0) sh2loc[shared[a]] = a;
}
// dense payload: row s of a shared camera this rank observes = its 27 (+27) partial sums, zero otherwise; then n_tail scalars
__global__ void k_xch_push(const int* __restrict__ sh2loc, const double* __restrict__ accA, const double* __restrict__ accB, const int n_shared,
                           const double* __restrict__ scal, const int n_tail, const PeerTable P, const int world, unsigned* counter, const unsigned seq) {
  pdl_grid_sync();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n_rows = (long)54 * n_shared;
  if (t < n_rows + n_tail) {
    double v = 0.0;
    if (t < n_rows) {
      const int sidx = (int)(t / 54), q = (int)(t - (long)sidx * 54);
      const int a = sh2loc[sidx];
      if (a >= 0) v = (q < 27) ? accA[(size_t)27 * a + q] : (accB ? accB[(size_t)27 * a + q - 27] : 0.0);
    } else v = scal[t - n_rows];
    for (int r = 0; r < world; ++r) P.area[r][t] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicInc(counter, gridDim.x - 1);
    if (done == gridDim.x - 1) {
      __threadfence_system();
      for (int r = 0; r < world; ++r) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(P.flag[r]), "r"(seq) : "memory");
    }
  }
}
// wait until every rank's copy has landed in THIS rank's buffer, add the W copies in rank order, unpack
__global__ void k_xch_pop(const int* __restrict__ sh2loc, const double* local_slot /* [world][stride] */, const unsigned* local_flag /* [world] */,
                          const size_t stride, const int world, const unsigned seq, const int n_shared, double* __restrict__ accA,
                          double* __restrict__ accB, double* __restrict__ scal, const int n_tail, int* err) {
  pdl_grid_sync();
  __shared__ int s_ok;
  if (threadIdx.x < 32) {
    bool ok = true;
    if ((int)threadIdx.x < world) {
      const unsigned long long t0 = global_ns();
      unsigned spins = 0;
      while ((int)(ld_relaxed_sys_u32(local_flag + threadIdx.x) - seq) < 0) {
        if ((++spins & 1023u) == 0 && global_ns() - t0 > 120000000000ull) { ok = false; break; }
      }
    }
    ok = __all_sync(0xffffffffu, ok);
    __threadfence_system();
    if (threadIdx.x == 0) { s_ok = ok ? 1 : 0; if (!ok) *err = 1; }
  }
  __syncthreads();
  if (!s_ok) return;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n_rows = (long)54 * n_shared;
  if (t >= n_rows + n_tail) return;
  int a = 0, q = 0;
  if (t < n_rows) {
    const int sidx = (int)(t / 54);
    q = (int)(t - (long)sidx * 54);
    a = sh2loc[sidx];
    if (a < 0) return;                 // a camera this rank does not observe
    if ((q < 27 && !accA) || (q >= 27 && !accB)) return;
  }
  double v = 0.0;
  for (int r = 0; r < world; ++r) v += __ldcg(local_slot + (size_t)r * stride + t);
  if (t >= n_rows) scal[t - n_rows] = v;
  else if (q < 27) accA[(size_t)27 * a + q] = v;
  else accB[(size_t)27 * a + q - 27] = v;
}
// scalars every rank needs before the host reads them: [0..4] the candidate-step sums over this rank's points,
// [5..9] the camera sums counted on owned cameras, [10..10+MAX_WORLD) max |g_camera| in this rank's slot
constexpr int NLATE = 10 + MAX_WORLD;
__global__ void k_late_pack(const double* __restrict__ scal, const int rank, double* __restrict__ out) {
  pdl_grid_sync();
  const int t = threadIdx.x;
  if (blockIdx.x != 0 || t >= NLATE) return;
  const int src[10] = {S_COST_C, S_YN2_P, S_YG_P, S_YLY_P, S_BAD_C, S_XN2_C, S_YN2_C, S_YG_C, S_YLY_C, S_NOTPD_C};
  out[t] = (t < 10) ? scal[src[t]] : ((t - 10 == rank) ? scal[S_GMAX_C] : 0.0);
}

// ---------------------------------------------------------------------------------------------
// Single-CTA camera kernels (camera-sized vectors: <= 10^4 x 6 doubles).
// ---------------------------------------------------------------------------------------------
// upper-triangular index of (r,c), r<=c, in the 21-entry packing used above
__device__ __forceinline__ int tri(int r, int c) { return r * 6 - (r * (r - 1)) / 2 + (c - r); }

// xg_i = T_i x_i = (G x_w, R x_t): the per-camera 6-vector the point passes gather.
__device__ __forceinline__ void apply_T(const double* ct, const double* x, double* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    out[r] = ct[CT_G + r * 3] * x[0] + ct[CT_G + r * 3 + 1] * x[1] + ct[CT_G + r * 3 + 2] * x[2];
    out[3 + r] = ct[r * 3] * x[3] + ct[r * 3 + 1] * x[4] + ct[r * 3 + 2] * x[5];
  }
}
__device__ __forceinline__ void apply_Tt(const double* ct, const double* y, double* out) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    out[r] = ct[CT_G + r] * y[0] + ct[CT_G + 3 + r] * y[1] + ct[CT_G + 6 + r] * y[2];
    out[3 + r] = ct[r] * y[3] + ct[3 + r] * y[4] + ct[6 + r] * y[5];
  }
}

// ---------------------------------------------------------------------------------------------
// Inspection: expand records into explicit loss-corrected residual / Jacobian blocks (glba_linearize).
// ---------------------------------------------------------------------------------------------
__global__ void k_expand(const long n_obs, const int* __restrict__ pm_cam, const double2* __restrict__ pm_uv,
                         const int* __restrict__ pm2orig, const double4* __restrict__ rec_pm,
                         const double* __restrict__ camtab, const Intr K, double* __restrict__ res,
                         double* __restrict__ jc, double* __restrict__ jp) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_obs) return;
  const int i = pm_cam[k];
  const long o = pm2orig ? pm2orig[k] : k;
  const double4 rec = rec_pm[k];
  const double2 uv = pm_uv[k];
  const double* ct = camtab + (size_t)CAMTAB * i;
  double a[6], b[6], ap[3], bp[3], R[9], ta[6], tb[6];
  for (int q = 0; q < 9; ++q) R[q] = ct[q];
  jhat_rows(rec, ct[21], ct[22], ct[23], K, a, b);
  jp_rows(rec, R, K, ap, bp);
  // J~c rows = a' T, b' T  (T = blockdiag(G,R))  ->  T' a
  apply_Tt(ct, a, ta);
  apply_Tt(ct, b, tb);
  if (res) { res[2 * o] = rec.w * (K.fx * rec.x + K.cx - uv.x); res[2 * o + 1] = rec.w * (K.fy * rec.y + K.cy - uv.y); }
  if (jc) for (int q = 0; q < 6; ++q) { jc[12 * o + q] = ta[q]; jc[12 * o + 6 + q] = tb[q]; }
  if (jp) for (int q = 0; q < 3; ++q) { jp[6 * o + q] = ap[q]; jp[6 * o + 3 + q] = bp[q]; }
}

// ---------------------------------------------------------------------------------------------
// Index construction helpers
// ---------------------------------------------------------------------------------------------
__global__ void k_iota(const long n, int* out) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = (int)k;
}
// keys of the camera -> (tile, slot) list: the camera of every (tile, slot) entry, n_cam for unused slots
__global__ void k_tp_keys(const long n, const int* __restrict__ cams, const int n_cam, int* __restrict__ key, int* __restrict__ val) {
  pdl_grid_sync();
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int c = cams[e];
  key[e] = (c >= 0) ? c : n_cam;
  val[e] = (int)e;
}
// camera of every overflow observation (keys of the camera -> overflow-row list)
__global__ void k_ovf_keys(const int n, const int* __restrict__ ovf_k, const int* __restrict__ pm_cam, int* __restrict__ key, int* __restrict__ val) {
  pdl_grid_sync();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  key[e] = pm_cam[ovf_k[e]];
  val[e] = e;
}
__global__ void k_check_sorted(const long n, const int* __restrict__ key, const int n_key, int* flags /* [0]=unsorted, [1]=out of range */) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int v = key[k];
  if (v < 0 || v >= n_key) flags[1] = 1;
  if (k > 0 && key[k - 1] > v) flags[0] = 1;
}
__global__ void k_check_range(const long n, const int* __restrict__ key, const int n_key, int* flag) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && (key[k] < 0 || key[k] >= n_key)) *flag = 1;
}
// start[s] = first k with key[k] >= s (key sorted ascending); start[n_seg] = n
__global__ void k_segment_starts(const long n, const int* __restrict__ key, const int n_seg, int* __restrict__ start) {
  pdl_grid_sync();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n_seg) return;
  long lo = 0, hi = n;
  while (lo < hi) { const long mid = (lo + hi) >> 1; if (key[mid] < s) lo = mid + 1; else hi = mid; }
  start[s] = (int)lo;
}
__global__ void k_gather_obs(const long n, const int* __restrict__ perm, const int* __restrict__ cam_in, const int* __restrict__ pt_in,
                             const double* __restrict__ u_in, const double* __restrict__ v_in, int* __restrict__ cam_out,
                             int* __restrict__ pt_out, double2* __restrict__ uv_out) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long o = perm ? perm[k] : k;
  cam_out[k] = cam_in[o];
  if (pt_out) pt_out[k] = pt_in[o];
  if (uv_out) uv_out[k] = make_double2(u_in[o], v_in[o]);
}
// measurements in both orders, once they have arrived (the upload of u, v overlaps the index construction)
__global__ void k_gather_uv(const long n, const int* __restrict__ pm2orig, const int* __restrict__ cm2pm, const double* __restrict__ u_in,
                            const double* __restrict__ v_in, double2* __restrict__ pm_uv, double2* __restrict__ cm_uv) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const long o = pm2orig ? pm2orig[k] : k;
  pm_uv[k] = make_double2(u_in[o], v_in[o]);
  const long q = cm2pm[k];
  const long oq = pm2orig ? pm2orig[q] : q;
  cm_uv[k] = make_double2(u_in[oq], v_in[oq]);
}
__global__ void k_build_cm(const long n, const int* __restrict__ cm2pm, const int* __restrict__ pm_pt, const double2* __restrict__ pm_uv,
                           int* __restrict__ cm_pt, double2* __restrict__ cm_uv, int* __restrict__ pm2cm) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int o = cm2pm[k];
  cm_pt[k] = pm_pt[o];
  if (cm_uv) cm_uv[k] = pm_uv[o];
  pm2cm[o] = (int)k;
}
__global__ void k_free_flags(const int n, const int* __restrict__ start, const uint8_t* __restrict__ fixed, const int* __restrict__ new2old,
                             uint8_t* __restrict__ free_out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool seen = start[i + 1] > start[i];
  free_out[i] = (seen && !(fixed && fixed[new2old ? new2old[i] : i])) ? 1 : 0;
}
// cnt[i] = local observations of camera i; cnt[n_cam] = local duplicate flag (all-reduced by the host when sharded)
__global__ void k_counts(const int n_cam, const int* __restrict__ cam_start, const int* __restrict__ dup_flag, const int empty,
                         int* __restrict__ cnt) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_cam) cnt[i] = cam_start[i + 1] - cam_start[i];
  else if (i == n_cam) cnt[i] = (*dup_flag != 0) ? 1 : 0;
  else if (i == n_cam + 1) cnt[i] = empty;
}
__global__ void k_free_flags_cnt(const int n, const int* __restrict__ cnt, const uint8_t* __restrict__ fixed, uint8_t* __restrict__ free_out) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) free_out[i] = (cnt[i] > 0 && !(fixed && fixed[i])) ? 1 : 0;
}
// Internal point j corresponds to the caller's point new2old[j] (nullptr = identity): see the locality relabelling
// in load_problem.
// .w carries sqrt(information weight) of the point's observations (1 when the caller gives none)
__global__ void k_pack_pt(const int n, const double* __restrict__ pt3, const double* __restrict__ info, const int* __restrict__ new2old,
                          double4* __restrict__ pt4) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t o = new2old ? new2old[j] : j;
  pt4[j] = make_double4(pt3[3 * o], pt3[3 * o + 1], pt3[3 * o + 2], info ? sqrt(fmax(info[o], 0.0)) : 1.0);
}
__global__ void k_unpack_pt(const int n, const double4* __restrict__ pt4, const int* __restrict__ new2old, double* __restrict__ pt3) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t o = new2old ? new2old[j] : j;
  const double4 v = pt4[j]; pt3[3 * o] = v.x; pt3[3 * o + 1] = v.y; pt3[3 * o + 2] = v.z;
}
// first (smallest) camera index observing each point, from the caller's (possibly unsorted) observation list
__global__ void k_first_cam(const long n, const int* __restrict__ obs_cam, const int* __restrict__ obs_pt, int* __restrict__ first_cam) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) atomicMin(first_cam + obs_pt[k], obs_cam[k]);
}
__global__ void k_fill_int(const int n, int* __restrict__ a, const int v) {
  pdl_grid_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
// number of points whose first camera is smaller than the previous observed point's (0 for creation-ordered ids)
__global__ void k_count_descents(const int n, const int* __restrict__ first_cam, int* __restrict__ count) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > 0 && j < n && first_cam[j] < first_cam[j - 1]) atomicAdd(count, 1);
}
__global__ void k_invert_perm(const int n, const int* __restrict__ new2old, int* __restrict__ old2new) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) old2new[new2old[j]] = j;
}
__global__ void k_relabel(const long n, const int* __restrict__ obs_pt, const int* __restrict__ old2new, int* __restrict__ out) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = old2new[obs_pt[k]];
}
__global__ void k_scatter_u8(const int n, const uint8_t* __restrict__ in, const int* __restrict__ new2old, uint8_t* __restrict__ out) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[new2old ? new2old[j] : j] = in[j];
}
__global__ void k_scatter_f64(const int n, const double* __restrict__ in, const int* __restrict__ new2old, double* __restrict__ out) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) out[new2old ? new2old[j] : j] = in[j];
}
// Craw SoA (6 C + 3 g) -> explicit 3x3 / gradient in AoS for glba_linearize
__global__ void k_unpack_pointblocks(const int n, const uint8_t* __restrict__ pt_free, const double* __restrict__ Craw,
                                     const int* __restrict__ new2old, double* __restrict__ hess /* 9 */, double* __restrict__ grad /* 3 */) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t o = new2old ? new2old[j] : j;
  const bool f = pt_free[j] != 0;
  double C[6], g[3];
  for (int q = 0; q < 6; ++q) C[q] = f ? Craw[(size_t)q * n + j] : 0.0;
  for (int q = 0; q < 3; ++q) g[q] = f ? Craw[(size_t)(6 + q) * n + j] : 0.0;
  if (hess) { double* H = hess + 9 * o; H[0] = C[0]; H[1] = C[1]; H[2] = C[2]; H[3] = C[1]; H[4] = C[3]; H[5] = C[4]; H[6] = C[2]; H[7] = C[4]; H[8] = C[5]; }
  if (grad) { grad[3 * o] = g[0]; grad[3 * o + 1] = g[1]; grad[3 * o + 2] = g[2]; }
}

// post_ba_map_point_culling arithmetic (slam_core.cpp:993-1035), one thread per point.
__global__ void __launch_bounds__(NT_PM)
k_cull(const PmArgs A, const double4* __restrict__ pt, const double* __restrict__ camtab, const int min_obs,
       const double max_mean_err, uint8_t* __restrict__ bad, double* __restrict__ mean_err) {
  pdl_grid_sync();
  const int j = blockIdx.x * NT_PM + threadIdx.x;
  if (j >= A.n_pt) return;
  const int b = A.pt_start[j], e = A.pt_start[j + 1];
  const double4 X = ldg4(pt + j);
  double tot = 0.0; int cnt = 0; bool isbad = false;
  for (int k = b; k < e; ++k) {
    const double* ct = camtab + (size_t)CAMTAB * A.pm_cam[k];
    const double2 uv = A.pm_uv[k];
    const double qx = X.x - ct[CT_C], qy = X.y - ct[CT_C + 1], qz = X.z - ct[CT_C + 2];
    const double px = ct[0] * qx + ct[1] * qy + ct[2] * qz;
    const double py = ct[3] * qx + ct[4] * qy + ct[5] * qz;
    const double pz = ct[6] * qx + ct[7] * qy + ct[8] * qz;
    if (pz <= 0.0) { isbad = true; break; }
    const double du = A.K.fx * px / pz + A.K.cx - uv.x, dv = A.K.fy * py / pz + A.K.cy - uv.y;
    tot += sqrt(du * du + dv * dv); ++cnt;
  }
  const double avg = cnt ? tot / cnt : 0.0;
  if (!isbad && (cnt < min_obs || avg > max_mean_err)) isbad = true;
  bad[j] = isbad ? 1 : 0;
  if (mean_err) mean_err[j] = avg;
}

}  // namespace glba
