// glba_dense.cuh — exact reduced-camera solve for small windows (<= DN_MAXCAM cameras), the regime
// GL-SLAM's live BA runs in (7+3 keyframes: src/core/slam_types.cpp:8-9, thread_pool.cpp:319-323;
// Ceres SPARSE_SCHUR = exact factorisation, slam_core.cpp:843).
//
//   S = blockdiag(B_i + Lambda_i) - sum_j V_j V_j',   V_ij = W_ij chol(Cinv_j),  W_ij = J~c' J~p
//   rhs_i = g_i - sum_j W_ij u0_j
// k_dense_schur: every CTA owns a fixed, contiguous range of points, stages their V blocks in shared
// memory tile by tile and accumulates its private copy of the (block-upper) S in registers — a
// block-sparse SYRK with no atomics.  k_dense_reduce sums the CTA copies in CTA order.
// k_dense_solve: one CTA, S in shared memory, Cholesky + two triangular solves.
// Everything is fixed-order, so the step is bit-reproducible.
#pragma once
#include "glba_kernels.cuh"

namespace glba {

constexpr int DN_MAXCAM = 16;            // reduced dimension <= 96
constexpr int DN_NT = 256;
constexpr int DN_TP = 40;                // points staged per tile (40 x n_cam x 192 B of shared memory: 123 KB at 16 cameras, beside 39 KB of accumulators)
constexpr int DN_SLOTS = 4;              // threads that walk one point's track while staging
constexpr int DN_PAIRS_PER_PASS = DN_NT / 36;                                                   // 7
constexpr int DN_MAXQ = (DN_MAXCAM * (DN_MAXCAM + 1) / 2 + DN_PAIRS_PER_PASS - 1) / DN_PAIRS_PER_PASS;  // 20

// lower Cholesky factor of a symmetric 3x3 (00,01,02,11,12,22); zeros if not PD
__device__ __forceinline__ void chol3(const double* C, double* L /* l00 l10 l11 l20 l21 l22 */) {
  // reciprocal square roots: one special-function chain per pivot instead of a square root and a divide
  const double i00 = C[0] > 0.0 ? rsqrt(C[0]) : 0.0;
  const double l00 = C[0] > 0.0 ? C[0] * i00 : 0.0;
  const double l10 = C[1] * i00, l20 = C[2] * i00;
  const double d1 = C[3] - l10 * l10;
  const double i11 = d1 > 0.0 ? rsqrt(d1) : 0.0;
  const double l11 = d1 > 0.0 ? d1 * i11 : 0.0;
  const double l21 = (C[4] - l20 * l10) * i11;
  const double d2 = C[5] - l20 * l20 - l21 * l21;
  const double l22 = d2 > 0.0 ? d2 * rsqrt(d2) : 0.0;
  L[0] = l00; L[1] = l10; L[2] = l11; L[3] = l20; L[4] = l21; L[5] = l22;
}

__global__ void __launch_bounds__(DN_NT)
k_dense_schur(const PmArgs A, const int n_cam, const uint8_t* __restrict__ cam_free, const double4* __restrict__ rec_pm,
              const double* __restrict__ camtab, const double* __restrict__ cinv, const double4* __restrict__ u0p, const int pts_per_cta,
              double* __restrict__ part /* [grid][n_pairs*36 + 6*n_cam] */, const LmCtl* __restrict__ ctl = nullptr) {
  pdl_grid_sync();
  extern __shared__ double dsm[];
  if (ctl_skip(ctl, GATE_ALWAYS)) return;
  const int n_pairs = n_cam * (n_cam + 1) / 2;
  double* V = dsm;                                   // [DN_TP][n_cam][18]
  double* Wu = V + (size_t)DN_TP * n_cam * 18;       // [DN_TP][n_cam][6]
  unsigned* pres = reinterpret_cast<unsigned*>(Wu + (size_t)DN_TP * n_cam * 6);   // [DN_TP]
  double* Sacc = reinterpret_cast<double*>(pres + DN_TP + (DN_TP & 1));     // [n_pairs][36]: this CTA's copy of the block-upper S
  __shared__ uint8_t plist[DN_TP][DN_MAXCAM];       // the cameras observing each staged point, ascending
  __shared__ uint8_t pair_a[DN_MAXCAM * (DN_MAXCAM + 1) / 2], pair_b[DN_MAXCAM * (DN_MAXCAM + 1) / 2];
  const int tid = threadIdx.x;
  for (int t = tid; t < n_pairs * 36; t += DN_NT) Sacc[t] = 0.0;
  // pair q of a point's camera list, a <= b, ordered by b then a: the first t(t+1)/2 entries are the pairs of a t-camera list
  if (tid < DN_MAXCAM * (DN_MAXCAM + 1) / 2) {
    int b2 = 0;
    while ((b2 + 1) * (b2 + 2) / 2 <= tid) ++b2;
    pair_a[tid] = (uint8_t)(tid - b2 * (b2 + 1) / 2); pair_b[tid] = (uint8_t)b2;
  }
  double racc = 0.0;                                 // thread t < 6*n_cam: rhs entry t
  const int p_begin = blockIdx.x * pts_per_cta;
  const int p_end = min(A.n_pt, p_begin + pts_per_cta);
  const int pl = tid / DN_SLOTS, slot = tid - pl * DN_SLOTS;
  for (int base = p_begin; base < p_end; base += DN_TP) {
    __syncthreads();
    if (tid < DN_TP) pres[tid] = 0u;
    __syncthreads();
    const int j = base + pl;
    if (pl < DN_TP && j < p_end && A.pt_free[j]) {
      const int b = A.pt_start[j], e = A.pt_start[j + 1];
      for (int k = b + slot; k < e; k += DN_SLOTS) {
        const int i = A.pm_cam[k];
        if (!cam_free[i]) continue;
        double Ci[6], u0[3];
        load_pblk(cinv, u0p, j, Ci, u0);
        double Lc[6];
        chol3(Ci, Lc);
        const double4 rec = rec_pm[k];
        const double* ct = camtab + (size_t)CAMTAB * i;
        double R[9], a[6], bb[6], ap[3], bp[3];
#pragma unroll
        for (int q = 0; q < 9; ++q) R[q] = ct[q];
        jhat_rows(rec, ct[21], ct[22], ct[23], A.K, a, bb);
        jp_rows(rec, R, A.K, ap, bp);
        double* Vo = V + ((size_t)pl * n_cam + i) * 18;
        double* Wo = Wu + ((size_t)pl * n_cam + i) * 6;
        // W = T' (a ap' + b bp')  (6x3);  column d of What, rotated by G' (rows 0-2) and R' (rows 3-5)
        double W[18];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          double wh[6], tw[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) wh[r] = a[r] * ap[d] + bb[r] * bp[d];
          apply_Tt(ct, wh, tw);
#pragma unroll
          for (int r = 0; r < 6; ++r) W[r * 3 + d] = tw[r];
        }
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          // V = W Lc,  Lc = [[l00,0,0],[l10,l11,0],[l20,l21,l22]]
          Vo[r * 3 + 0] = W[r * 3] * Lc[0] + W[r * 3 + 1] * Lc[1] + W[r * 3 + 2] * Lc[3];
          Vo[r * 3 + 1] = W[r * 3 + 1] * Lc[2] + W[r * 3 + 2] * Lc[4];
          Vo[r * 3 + 2] = W[r * 3 + 2] * Lc[5];
          Wo[r] = W[r * 3] * u0[0] + W[r * 3 + 1] * u0[1] + W[r * 3 + 2] * u0[2];
        }
        atomicOr(&pres[pl], 1u << i);
      }
    }
    __syncthreads();
    if (tid < DN_TP) {                                // camera list of every staged point
      unsigned m = pres[tid];
      int q = 0;
      while (m) { const int i = __ffs(m) - 1; plist[tid][q++] = (uint8_t)i; m &= m - 1; }
    }
    __syncthreads();
    // Accumulate point by point.  A point seen by t cameras contributes t(t+1)/2 blocks of 36 entries: those (pair, entry)
    // tasks are spread over the CTA, every task adds into ITS accumulator in shared memory, and a barrier separates the
    // points, so each accumulator receives its contributions in point order (fixed) and no two threads ever touch the same
    // one at a time.  (The first version gave every thread a fixed set of pairs in registers and tested each point's
    // presence mask against them: 82 % of the tests failed on 4-camera tracks and that loop was 80 % of the kernel.)
    const int np = min(DN_TP, p_end - base);
    for (int p = 0; p < np; ++p) {
      const unsigned mask = pres[p];
      if (mask) {                                     // uniform over the CTA
        const int t = __popc(mask);
        const int ntask = (t * (t + 1) / 2) * 36;
        const double* Vp = V + (size_t)p * n_cam * 18;
        for (int task = tid; task < ntask; task += DN_NT) {
          const int pi = task / 36, ent = task - pi * 36;
          const int rr = ent / 6, cc = ent - rr * 6;
          const int i = plist[p][pair_a[pi]], k = plist[p][pair_b[pi]];                  // the a-th and b-th observing cameras, i <= k
          const int pr = i * n_cam - (i * (i - 1)) / 2 + (k - i);
          const double* vi = Vp + i * 18 + rr * 3;
          const double* vk = Vp + k * 18 + cc * 3;
          Sacc[pr * 36 + ent] += vi[0] * vk[0] + vi[1] * vk[1] + vi[2] * vk[2];
        }
        if (tid < 6 * n_cam) {
          const int i = tid / 6;
          if ((mask >> i) & 1u) racc += Wu[((size_t)p * n_cam + i) * 6 + (tid - 6 * i)];
        }
      }
      __syncthreads();
    }
  }
  double* out = part + (size_t)blockIdx.x * ((size_t)n_pairs * 36 + 6 * n_cam);
  for (int t = tid; t < n_pairs * 36; t += DN_NT) out[t] = Sacc[t];
  if (tid < 6 * n_cam) out[(size_t)n_pairs * 36 + tid] = racc;
}

// Sums the per-CTA copies in CTA order and emits the ASSEMBLED reduced system: S (n x n, row-major, symmetric, with
// B + Lambda on the diagonal blocks and identity rows for non-free cameras) followed by rhs (n).
__global__ void k_dense_reduce(const int n_parts, const int n_cam, const double* __restrict__ part, const uint8_t* __restrict__ cam_free,
                               const double* __restrict__ Bc, const double* __restrict__ gc, const double* __restrict__ lamc,
                               const double inv_radius_arg, double* __restrict__ Sred /* pair sums (kept for the sharded all-reduce) */,
                               double* __restrict__ Sfull /* n*n + n */, const int assemble, const LmCtl* __restrict__ ctl = nullptr) {
  pdl_grid_sync();
  if (ctl_skip(ctl, GATE_ALWAYS)) return;
  const double inv_radius = ctl_inv_radius(ctl, inv_radius_arg);
  // 8 lanes per element: lane `sub` sums copies sub, sub+8, ... (8x shorter dependent load chains), the 8 partial sums
  // are combined by a fixed xor butterfly, so the result is still order-deterministic
  const int n_pairs = n_cam * (n_cam + 1) / 2;
  const int len = n_pairs * 36 + 6 * n_cam;
  const int sub = threadIdx.x & 7;
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = t < len;
  double s = 0.0;
  if (n_parts > 0) {
    if (valid) {
      // same summation order as one load at a time, four loads in flight
      int g = sub;
      for (; g + 24 < n_parts; g += 32) {
        const double v0 = part[(size_t)g * len + t], v1 = part[(size_t)(g + 8) * len + t], v2 = part[(size_t)(g + 16) * len + t],
                     v3 = part[(size_t)(g + 24) * len + t];
        s += v0; s += v1; s += v2; s += v3;
      }
      for (; g < n_parts; g += 8) s += part[(size_t)g * len + t];
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (valid && sub == 0) Sred[t] = s;
  } else if (valid) {
    s = Sred[t];          // already reduced (and all-reduced across ranks)
  }
  if (!valid || sub != 0) return;
  if (!assemble) return;
  const int n = 6 * n_cam;
  if (t < n_pairs * 36) {
    const int pr = t / 36, ent = t - pr * 36;
    int i = 0, rem = pr;
    while (rem >= n_cam - i) { rem -= n_cam - i; ++i; }
    const int k = i + rem;
    const int rr = ent / 6, cc = ent - rr * 6;
    double v = -s;
    const bool fi = cam_free[i] != 0, fk = cam_free[k] != 0;
    if (i == k) {
      if (fi) { v += Bc[(size_t)36 * i + rr * 6 + cc]; if (rr == cc) v += lamc[6 * i + rr] * inv_radius; }
      else v = (rr == cc) ? 1.0 : 0.0;
    } else if (!fi || !fk) v = 0.0;
    Sfull[(size_t)(6 * i + rr) * n + 6 * k + cc] = v;
    if (i != k) Sfull[(size_t)(6 * k + cc) * n + 6 * i + rr] = v;
  } else {
    const int r = t - n_pairs * 36;
    Sfull[(size_t)n * n + r] = cam_free[r / 6] ? gc[r] - s : 0.0;
  }
}

// One CTA: assemble S (n = 6 n_cam <= 96) in shared memory, Cholesky, solve S y = rhs.  Non-free cameras become
// identity rows with zero rhs.  Thread r owns ROW r: left-looking factorisation where every thread recomputes the
// pivot of column j itself, so each column (and each substitution step) costs exactly one barrier.
// Row stride is odd (no shared-memory bank conflicts between rows).  Outputs y (cg_x layout), diagonal blocks, rhs.
constexpr int DN_NS = 128;   // threads of the solve kernel (>= 6*DN_MAXCAM)
__global__ void __launch_bounds__(DN_NS)
k_dense_solve(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ Sfull /* n*n + n, assembled */,
              double* __restrict__ y, double* __restrict__ Md, double* __restrict__ rhs_out, double* __restrict__ scal,
              const LmCtl* __restrict__ ctl = nullptr) {
  pdl_grid_sync();
  extern __shared__ double dsm[];
  if (ctl_skip(ctl, GATE_ALWAYS)) return;
  const int n = 6 * n_cam;
  const int ld = n | 1;
  double* S = dsm;                       // n x ld, lower triangle is what the factorisation reads
  double* ys = S + (size_t)n * ld;       // n
  double* dg = ys + n;                   // n: 1/L[r][r]
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int t = tid; t < n * n; t += DN_NS) { const int rr = t / n, cc = t - rr * n; S[(size_t)rr * ld + cc] = Sfull[t]; }
  const int r = tid;
  double b = 0.0;
  if (r < n) { b = Sfull[(size_t)n * n + r]; rhs_out[r] = b; }
  __syncthreads();
  for (int t = tid; t < n_cam * 36; t += DN_NS) {
    const int i = t / 36, ent = t - i * 36, rr = ent / 6, cc = ent - rr * 6;
    Md[t] = cam_free[i] ? S[(size_t)(6 * i + rr) * ld + 6 * i + cc] : 0.0;
  }
  // ---- blocked (6x6 = one camera) left-looking Cholesky: per block column two barriers; every thread carries six
  //      independent FMA chains, and factors the 6x6 pivot block redundantly in registers (no serial owner thread)
  for (int J = 0; J < n_cam; ++J) {
    const int c0 = 6 * J;
    const bool act = (r >= c0 && r < n);
    double t[6];
    if (act) {
      const double* Sr = S + (size_t)r * ld;
#pragma unroll
      for (int q = 0; q < 6; ++q) t[q] = Sr[c0 + q];
      for (int k = 0; k < c0; ++k) {
        const double lrk = Sr[k];
#pragma unroll
        for (int q = 0; q < 6; ++q) t[q] -= lrk * S[(size_t)(c0 + q) * ld + k];
      }
      double* Sw = S + (size_t)r * ld;
#pragma unroll
      for (int q = 0; q < 6; ++q) Sw[c0 + q] = t[q];
    }
    __syncthreads();
    if (act) {
      double Lb[21], id[6];     // pivot block factor (lower, packed a(a+1)/2+b) and inverse diagonal
#pragma unroll
      for (int a2 = 0; a2 < 6; ++a2)
#pragma unroll
        for (int b2 = 0; b2 <= a2; ++b2) Lb[a2 * (a2 + 1) / 2 + b2] = S[(size_t)(c0 + a2) * ld + c0 + b2];
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        double d = Lb[j * (j + 1) / 2 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= Lb[j * (j + 1) / 2 + k] * Lb[j * (j + 1) / 2 + k];
        if (!(d > 0.0)) { bad = true; d = 1.0; }
        const double inv = rsqrt(d);          // one reciprocal square root instead of sqrt + divide: the pivot chain is
        const double ljj = d * inv;           // the serial part of this kernel (6 dependent pivots per block column)
        Lb[j * (j + 1) / 2 + j] = ljj; id[j] = inv;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
          double v = Lb[i * (i + 1) / 2 + j];
#pragma unroll
          for (int k = 0; k < j; ++k) v -= Lb[i * (i + 1) / 2 + k] * Lb[j * (j + 1) / 2 + k];
          Lb[i * (i + 1) / 2 + j] = v * inv;
        }
      }
      if (bad && r == c0) s_bad = 1;
      double* Sw = S + (size_t)r * ld;
      if (r < c0 + 6) {
        const int a2 = r - c0;   // static indexing only (keeps Lb / id in registers)
#pragma unroll
        for (int a3 = 0; a3 < 6; ++a3)
          if (a3 == a2) {
#pragma unroll
            for (int b2 = 0; b2 <= a3; ++b2) Sw[c0 + b2] = Lb[a3 * (a3 + 1) / 2 + b2];
            dg[r] = id[a3];     // dg now holds 1/L[r][r]
          }
      } else {
        // x Lb' = t  (row of the block column below the pivot block)
        double x[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          double v = t[q];
#pragma unroll
          for (int p2 = 0; p2 < q; ++p2) v -= x[p2] * Lb[q * (q + 1) / 2 + p2];
          x[q] = v * id[q];
          Sw[c0 + q] = x[q];
        }
      }
    }
    __syncthreads();
  }
  // ---- L z = b, blocked: thread r carries b_r; the six pivot rows publish their entries, everyone solves the 6x6
  for (int J = 0; J < n_cam; ++J) {
    const int c0 = 6 * J;
    if (r >= c0 && r < c0 + 6) ys[r] = b;
    __syncthreads();
    if (r >= c0 && r < n) {
      double z[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) {
        double v = ys[c0 + q];
#pragma unroll
        for (int p2 = 0; p2 < q; ++p2) v -= S[(size_t)(c0 + q) * ld + c0 + p2] * z[p2];
        z[q] = v * dg[c0 + q];
      }
      if (r < c0 + 6) {
#pragma unroll
        for (int q = 0; q < 6; ++q) if (q == r - c0) b = z[q];
      } else {
        const double* Sr = S + (size_t)r * ld;
#pragma unroll
        for (int q = 0; q < 6; ++q) b -= Sr[c0 + q] * z[q];
      }
    }
    __syncthreads();      // ys[c0..] is rewritten by nobody, but keeps the reads of block J ahead of block J+1's publishes
  }
  // ---- L' y = z, blocked, bottom-up
  for (int J = n_cam - 1; J >= 0; --J) {
    const int c0 = 6 * J;
    if (r >= c0 && r < c0 + 6) ys[r] = b;
    __syncthreads();
    if (r < c0 + 6 && r < n) {
      double yb[6];
#pragma unroll
      for (int q = 5; q >= 0; --q) {
        double v = ys[c0 + q];
#pragma unroll
        for (int p2 = 5; p2 > q; --p2) v -= S[(size_t)(c0 + p2) * ld + c0 + q] * yb[p2];
        yb[q] = v * dg[c0 + q];
      }
      if (r >= c0) {
#pragma unroll
        for (int q = 0; q < 6; ++q) if (q == r - c0) b = yb[q];
      } else {
#pragma unroll
        for (int q = 0; q < 6; ++q) b -= S[(size_t)(c0 + q) * ld + r] * yb[q];
      }
    }
    __syncthreads();
  }
  if (r < n) ys[r] = b;
  __syncthreads();
  if (r < n) y[r] = cam_free[r / 6] ? ys[r] : 0.0;
  if (tid == 0) scal[S_NOTPD_C] = s_bad ? 1.0 : 0.0;
}

// duplicate (point, camera) observations make two staging threads collide: detect them at load time
__global__ void k_check_dup(const int n_pt, const int* __restrict__ pt_start, const int* __restrict__ pm_cam, int* flag) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pt) return;
  const int b = pt_start[j], e = pt_start[j + 1];
  for (int k = b; k < e; ++k)
    for (int l = k + 1; l < e; ++l)
      if (pm_cam[k] == pm_cam[l]) { *flag = 1; return; }
}

}  // namespace glba
