// glba_cam.cuh — camera-sized kernels (<= 10^4 cameras): Hessian-block finalisation, block-Jacobi
// preconditioner, the vector half of PCG, candidate cameras.  One thread per camera, many small CTAs;
// scalars are reduced per CTA and summed in CTA order by the last CTA to finish (fixed order).
#pragma once
#include <climits>
#include "glba_kernels.cuh"
#include "glba_tiles.cuh"

namespace glba {

constexpr int NT_C = 64;

__device__ __forceinline__ void mm3(const double* A, const double* B, double* C) {        // C = A B
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
__device__ __forceinline__ void mtm3(const double* A, const double* B, double* C) {       // C = A' B
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}

// Bout (6x6 row-major, full) = T' A T with T = blockdiag(G, R), A symmetric given by its 21 upper entries.
__device__ __forceinline__ void congruence_T(const double* __restrict__ a21, const double* G, const double* R, double* Bout) {
  double A11[9], A12[9], A22[9], t[9], o[9];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A11[r * 3 + c] = (r <= c) ? a21[tri(r, c)] : a21[tri(c, r)];
      A12[r * 3 + c] = a21[tri(r, 3 + c)];
      A22[r * 3 + c] = (r <= c) ? a21[tri(3 + r, 3 + c)] : a21[tri(3 + c, 3 + r)];
    }
  mm3(A11, G, t); mtm3(G, t, o);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Bout[r * 6 + c] = o[r * 3 + c];
  mm3(A12, R, t); mtm3(G, t, o);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { Bout[r * 6 + 3 + c] = o[r * 3 + c]; Bout[(3 + c) * 6 + r] = o[r * 3 + c]; }
  mm3(A22, R, t); mtm3(R, t, o);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Bout[(3 + r) * 6 + 3 + c] = o[r * 3 + c];
  // symmetrise exactly (the two triangular halves of a congruence differ in the last bit)
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = r + 1; c < 6; ++c) Bout[c * 6 + r] = Bout[r * 6 + c];
}

// fully unrolled 6x6 SPD inverse (registers only); false if not positive definite
__device__ __forceinline__ bool inv6_spd_reg(const double* A, double* Ai) {
  double L[21];      // lower triangle, row-major packed: L[r(r+1)/2 + c]
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j * 6 + j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
    if (!(d > 0.0)) { ok = false; d = 1.0; }
    const double ljj = sqrt(d);
    L[j * (j + 1) / 2 + j] = ljj;
    const double inv = 1.0 / ljj;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i * 6 + j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
      L[i * (i + 1) / 2 + j] = s * inv;
    }
  }
  // Linv (lower), then Ai = Linv' Linv
  double Li[21];
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    Li[c * (c + 1) / 2 + c] = 1.0 / L[c * (c + 1) / 2 + c];
#pragma unroll
    for (int r = c + 1; r < 6; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = c; k < r; ++k) s -= L[r * (r + 1) / 2 + k] * Li[k * (k + 1) / 2 + c];
      Li[r * (r + 1) / 2 + c] = s / L[r * (r + 1) / 2 + r];
    }
  }
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = r; c < 6; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = c; k < 6; ++k) s += Li[k * (k + 1) / 2 + r] * Li[k * (k + 1) / 2 + c];
      Ai[r * 6 + c] = s; Ai[c * 6 + r] = s;
    }
  return ok;
}

// CTA partials -> scalars, summed in CTA order by whichever CTA finishes last.
template <int NV>
__device__ __forceinline__ bool finish_scalars(const double (&v)[NV], const bool (&is_max)[NV], const int (&slot)[NV], double* part,
                                               unsigned* counter, double* scal, double* sm /* NV*NT_C/32 */, double* smo /* NV */) {
  // block reduce (sum or max per entry)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double y = __shfl_xor_sync(0xffffffffu, x, o); x = is_max[i] ? fmax(x, y) : x + y; }
    if (lane == 0) sm[wid * NV + i] = x;
  }
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = sm[i];
      for (int w = 1; w < NT_C / 32; ++w) x = is_max[i] ? fmax(x, sm[w * NV + i]) : x + sm[w * NV + i];
      part[(size_t)blockIdx.x * NV + i] = x;
    }
    __threadfence();
    const unsigned t = atomicInc(counter, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = __ldcg(part + i);
      for (unsigned b = 1; b < gridDim.x; ++b) { const double y = __ldcg(part + (size_t)b * NV + i); x = is_max[i] ? fmax(x, y) : x + y; }
      smo[i] = x;
      if (slot[i] >= 0) scal[slot[i]] = x;
    }
  }
  return s_last;     // true in every thread of the last CTA; its thread 0 has the totals in smo[]
}

// per-camera 27 partial sums: either already summed (acc27, sharded runs: all-reduced) or summed here from the chunk
// partials in chunk order (single GPU: one launch less)
template <bool FUSE>
__device__ __forceinline__ void load_acc27(const int i, const int n_cam, const double* __restrict__ acc27, const int* __restrict__ cam_chunk_start,
                                           const double* __restrict__ part27, double* a, double (*stage)[27]) {
  if (FUSE && n_cam > 256) {
    // many cameras: one thread per camera keeps every lane busy already (and measured faster than staging: 25 vs 43 us
    // for both finalisers at 1 800 cameras)
    if (i < n_cam) {
#pragma unroll
      for (int q = 0; q < 27; ++q) a[q] = 0.0;
      for (int ch = cam_chunk_start[i]; ch < cam_chunk_start[i + 1]; ++ch)
#pragma unroll
        for (int q = 0; q < 27; ++q) a[q] += part27[(size_t)27 * ch + q];
    }
  } else if (FUSE) {
    // few cameras: the CTA's (camera, value) pairs are spread over all its threads, so a window of 10 cameras uses 64
    // lanes rather than 10; every sum still runs over the camera's chunks in chunk order (fixed)
    const int cam0 = blockIdx.x * NT_C;
    const int ncl = min(NT_C, n_cam - cam0);
    for (int e = threadIdx.x; e < ncl * 27; e += NT_C) {
      const int cl = e / 27, q = e - cl * 27;
      // same summation order as a plain loop, but four loads in flight at a time (the dependent L2 round trips of the
      // one-at-a-time loop were most of this kernel on a 10-camera window)
      double sacc = 0.0;
      int ch = cam_chunk_start[cam0 + cl];
      const int ce = cam_chunk_start[cam0 + cl + 1];
      for (; ch + 4 <= ce; ch += 4) {
        const double v0 = part27[(size_t)27 * ch + q], v1 = part27[(size_t)27 * (ch + 1) + q], v2 = part27[(size_t)27 * (ch + 2) + q],
                     v3 = part27[(size_t)27 * (ch + 3) + q];
        sacc += v0; sacc += v1; sacc += v2; sacc += v3;
      }
      for (; ch < ce; ++ch) sacc += part27[(size_t)27 * ch + q];
      stage[cl][q] = sacc;
    }
    __syncthreads();
    if (i < n_cam) {
#pragma unroll
      for (int q = 0; q < 27; ++q) a[q] = stage[threadIdx.x][q];
    }
  } else if (i < n_cam) {
#pragma unroll
    for (int q = 0; q < 27; ++q) a[q] = acc27[(size_t)27 * i + q];
  }
}

// B_i = T' A T, g_i = T' ghat; Jacobi scale (iteration 0); lam_c = clamp(s^2 h)/s^2; |x_c|^2, max |g_c|.
template <bool FUSE>
__device__ __forceinline__ void
cam_lin_fin_body(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ cam, const double* __restrict__ camtab,
              const double* __restrict__ acc27, const int* __restrict__ cam_chunk_start, const double* __restrict__ part27, double* __restrict__ Bc, double* __restrict__ gc, double* __restrict__ sc,
              double* __restrict__ lamc, const int first, const int jacobi, const double min_diag, const double max_diag,
              double* part, unsigned* counter, double* scal, const LmCtl* __restrict__ ctl, const LmHook hook,
              const uint8_t* __restrict__ owned /* sharded: scalars count a camera on its owner rank only */) {
  __shared__ double sm[2 * NT_C / 32];
  if (ctl_skip(ctl, GATE_ACCEPTED)) return;
  __shared__ double smo[2];
  __shared__ double stage[FUSE ? NT_C : 1][27];
  const int i = blockIdx.x * NT_C + threadIdx.x;
  double xn2 = 0.0, gmax = 0.0;
  double a[27];
  load_acc27<FUSE>(i, n_cam, acc27, cam_chunk_start, part27, a, stage);
  if (i < n_cam) {
    double* B = Bc + (size_t)36 * i;
    if (!cam_free[i]) {
#pragma unroll
      for (int q = 0; q < 36; ++q) B[q] = 0.0;
#pragma unroll
      for (int q = 0; q < 6; ++q) { gc[6 * i + q] = 0.0; lamc[6 * i + q] = 0.0; if (first) sc[6 * i + q] = 1.0; }
    } else {
      const double* ct = camtab + (size_t)CAMTAB * i;
      double G[9], R[9], Bl[36], gh[6], g6[6];
#pragma unroll
      for (int q = 0; q < 9; ++q) { R[q] = ct[q]; G[q] = ct[CT_G + q]; }
      congruence_T(a, G, R, Bl);
#pragma unroll
      for (int q = 0; q < 36; ++q) B[q] = Bl[q];
#pragma unroll
      for (int q = 0; q < 6; ++q) gh[q] = a[21 + q];
      apply_Tt(ct, gh, g6);
      const bool mine = owned == nullptr || owned[i] != 0;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        gc[6 * i + r] = g6[r];
        if (mine) gmax = fmax(gmax, fabs(g6[r]));
        const double h = Bl[r * 7];
        double sv;
        if (first) { sv = jacobi ? 1.0 / (1.0 + sqrt(h)) : 1.0; sc[6 * i + r] = sv; } else sv = sc[6 * i + r];
        const double s2 = sv * sv;
        lamc[6 * i + r] = fmin(fmax(s2 * h, min_diag), max_diag) / s2;
        if (mine) xn2 += cam[6 * i + r] * cam[6 * i + r];
      }
    }
  }
  const double v[2] = {xn2, gmax};
  const bool mx[2] = {false, true};
  const int slot[2] = {S_XN2_C, S_GMAX_C};
  const bool last = finish_scalars<2>(v, mx, slot, part, counter, scal, sm, smo);
  // device-resident LM loop: this was the last kernel of the re-linearisation after an accepted step
  if (last && threadIdx.x == 0 && hook.ctl != nullptr) { __threadfence(); lm_absorb(hook.ctl, hook.P, scal, hook.sum); }
}
template <bool FUSE>
__global__ void __launch_bounds__(NT_C)
k_cam_lin_fin(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ cam, const double* __restrict__ camtab,
              const double* __restrict__ acc27, const int* __restrict__ cam_chunk_start, const double* __restrict__ part27, double* __restrict__ Bc, double* __restrict__ gc, double* __restrict__ sc,
              double* __restrict__ lamc, const int first, const int jacobi, const double min_diag, const double max_diag,
              double* part, unsigned* counter, double* scal, const LmCtl* __restrict__ ctl = nullptr, const LmHook hook = LmHook{},
              const uint8_t* __restrict__ owned = nullptr) {
  pdl_grid_sync();
  cam_lin_fin_body<FUSE>(n_cam, cam_free, cam, camtab, acc27, cam_chunk_start, part27, Bc, gc, sc, lamc, first, jacobi, min_diag, max_diag, part, counter, scal, ctl, hook, owned);
}

// M_i = B_i + lam/radius - T' Mhat T (diagonal block of S), rhs_i = g_i - T' rhat, Minv_i.
template <bool FUSE>
__device__ __forceinline__ void
cam_schur_fin_body(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ camtab, const double* __restrict__ acc27,
                const int* __restrict__ cam_chunk_start, const double* __restrict__ part27,
                const double* Bc, const double* gc, const double* lamc, const double inv_radius,
                double* __restrict__ Md, double* __restrict__ Minv, double* __restrict__ rhs, double* part, unsigned* counter, double* scal,
                const uint8_t* __restrict__ owned) {
  __shared__ double sm[NT_C / 32];
  __shared__ double smo[1];
  __shared__ double stage[FUSE ? NT_C : 1][27];
  const int i = blockIdx.x * NT_C + threadIdx.x;
  double notpd = 0.0;
  double a[27];
  load_acc27<FUSE>(i, n_cam, acc27, cam_chunk_start, part27, a, stage);
  if (i < n_cam) {
    double* M = Md + (size_t)36 * i; double* Mi = Minv + (size_t)36 * i;
    if (!cam_free[i]) {
#pragma unroll
      for (int q = 0; q < 36; ++q) { M[q] = 0.0; Mi[q] = 0.0; }
#pragma unroll
      for (int q = 0; q < 6; ++q) rhs[6 * i + q] = 0.0;
    } else {
      const double* ct = camtab + (size_t)CAMTAB * i;
      double G[9], R[9], Ml[36], Il[36], rh[6], r6[6];
#pragma unroll
      for (int q = 0; q < 9; ++q) { R[q] = ct[q]; G[q] = ct[CT_G + q]; }
      congruence_T(a, G, R, Ml);
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double v = Bc[(size_t)36 * i + r * 6 + c] - Ml[r * 6 + c];
          if (r == c) v += lamc[6 * i + r] * inv_radius;
          Ml[r * 6 + c] = v;
        }
#pragma unroll
      for (int q = 0; q < 36; ++q) M[q] = Ml[q];
      if (!inv6_spd_reg(Ml, Il)) {
        if (owned == nullptr || owned[i] != 0) notpd = 1.0;
#pragma unroll
        for (int q = 0; q < 36; ++q) Il[q] = 0.0;
      }
#pragma unroll
      for (int q = 0; q < 36; ++q) Mi[q] = Il[q];
#pragma unroll
      for (int q = 0; q < 6; ++q) rh[q] = a[21 + q];
      apply_Tt(ct, rh, r6);
#pragma unroll
      for (int r = 0; r < 6; ++r) rhs[6 * i + r] = gc[6 * i + r] - r6[r];
    }
  }
  const double v[1] = {notpd};
  const bool mx[1] = {false};
  const int slot[1] = {S_NOTPD_C};
  finish_scalars<1>(v, mx, slot, part, counter, scal, sm, smo);
}
template <bool FUSE>
__global__ void __launch_bounds__(NT_C)
k_cam_schur_fin(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ camtab, const double* __restrict__ acc27,
                const int* __restrict__ cam_chunk_start, const double* __restrict__ part27,
                const double* __restrict__ Bc, const double* __restrict__ gc, const double* __restrict__ lamc, const double inv_radius,
                double* __restrict__ Md, double* __restrict__ Minv, double* __restrict__ rhs, double* part, unsigned* counter, double* scal,
                const uint8_t* __restrict__ owned = nullptr) {
  pdl_grid_sync();
  cam_schur_fin_body<FUSE>(n_cam, cam_free, camtab, acc27, cam_chunk_start, part27, Bc, gc, lamc, inv_radius, Md, Minv, rhs, part, counter, scal, owned);
}
// Both finalisations of a linearisation with Schur pieces in one launch (thread i finishes camera i twice: the Schur half reads
// the B_i, g_i and LM diagonal the same thread has just written).  Used behind k_cam_pipe, where both sets of sums arrive together.
template <bool FUSE>
__global__ void __launch_bounds__(NT_C)
k_cam_fin_both(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ cam, const double* __restrict__ camtab,
               const double* __restrict__ accA, const double* __restrict__ accB, const int* __restrict__ cam_chunk_start, const double* __restrict__ partA,
               const double* __restrict__ partB, double* Bc, double* gc, double* __restrict__ sc, double* lamc, const int first, const int jacobi,
               const double min_diag, const double max_diag, const double inv_radius, double* __restrict__ Md, double* __restrict__ Minv,
               double* __restrict__ rhs, double* part, unsigned* counter_lin, unsigned* counter_schur, double* scal, const uint8_t* __restrict__ owned) {
  pdl_grid_sync();
  cam_lin_fin_body<FUSE>(n_cam, cam_free, cam, camtab, accA, cam_chunk_start, partA, Bc, gc, sc, lamc, first, jacobi, min_diag, max_diag, part, counter_lin, scal,
                         (const LmCtl*)nullptr, LmHook{}, owned);
  __syncthreads();
  cam_schur_fin_body<FUSE>(n_cam, cam_free, camtab, accB, cam_chunk_start, partB, Bc, gc, lamc, inv_radius, Md, Minv, rhs, part + 4 * gridDim.x, counter_schur, scal, owned);
}

__device__ __forceinline__ void write_xtab(double* __restrict__ xr, const double* ct, const double* x6) {
  double t[6];
  apply_T(ct, x6, t);
#pragma unroll
  for (int q = 0; q < 6; ++q) xr[q] = t[q];
#pragma unroll
  for (int q = 0; q < 9; ++q) xr[6 + q] = ct[q];
  xr[15] = (ct[CT_SV] != 0.0 || ct[CT_SV + 1] != 0.0 || ct[CT_SV + 2] != 0.0) ? 1.0 : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Block-Jacobi PCG on the implicit Schur complement, single-reduction (Chronopoulos-Gear) form:
//     u = M^-1 r,  w = S u,  gamma' = r.u,  delta = u.w            <- ONE reduction point per iteration
//     beta = gamma'/gamma,  alpha = gamma' / (delta - beta gamma'/alpha_prev)
//     p = u + beta p,  s = w + beta s  (= S p),  x += alpha p,  r -= alpha s
// Same iterates as textbook PCG; both dot products are available right after the product, so a sharded solve needs one
// collective per iteration (payload: w on the cameras several ranks observe + the two scalars) and a single-GPU solve
// two camera-sized launches (k_cg_w, k_cg_update) next to the two halves of the product.
// Launch li of every kernel exits when cg->done_at <= li.  The stop test of iteration li uses gamma' = r_li . u_li, i.e.
// it is the textbook test of iteration li-1 evaluated one product later (the price of the single reduction).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void minv_apply(const double* __restrict__ Mi, const double* rr, double* z) {
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    double sacc = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) sacc += Mi[a * 6 + c] * rr[c];
    z[a] = sacc;
  }
}

// Fused product (k_pt_pipe<2>): the tile kernel leaves, per tile and camera slot, the six sums of that camera's observations
// in the tile.  One warp per camera adds its entries (camera -> (tile, slot) list built at load time, tile order): lane-strided
// partial sums + a fixed xor butterfly, so the result does not depend on scheduling.
__global__ void __launch_bounds__(256)
k_cam_combine(const int n_cam, const int* __restrict__ cam_tp_start, const int* __restrict__ cam_tp, const double* __restrict__ tpart,
              const int* __restrict__ cam_ov_start /* may be null */, const int* __restrict__ cam_ov, const double* __restrict__ ovf_c,
              double* __restrict__ yhat, const CgState* __restrict__ cg, const int li) {
  pdl_grid_sync();
  if (cg && cg->done_at <= li) return;
  const int w = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_cam) return;
  const int e0 = cam_tp_start[w], e1 = cam_tp_start[w + 1];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int e = e0 + lane; e < e1; e += 32) {
    const double2* t = reinterpret_cast<const double2*>(tpart + (size_t)6 * cam_tp[e]);
    const double2 a = t[0], b = t[1], c = t[2];
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y;
  }
  if (cam_ov_start != nullptr)       // observations that found no slot in their tile, in observation order
    for (int e = cam_ov_start[w] + lane; e < cam_ov_start[w + 1]; e += 32) {
      const double* t = ovf_c + (size_t)6 * cam_ov[e];
#pragma unroll
      for (int q = 0; q < 6; ++q) acc[q] += t[q];
    }
#pragma unroll
  for (int q = 0; q < 6; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < 6; ++q) yhat[(size_t)6 * w + q] = acc[q];
  }
}

// x = 0, r = rhs, u = M^-1 r, p = s = 0, gather table <- T u
__global__ void __launch_bounds__(NT_C)
k_cg_start(const int n_cam, const double* __restrict__ camtab, const double* __restrict__ Minv, const double* __restrict__ rhs,
           double* __restrict__ x, double* __restrict__ r, double* __restrict__ u, double* __restrict__ p, double* __restrict__ sv,
           double* __restrict__ xtab, CgState* cg, const double tol, const int max_iters) {
  pdl_grid_sync();
  const int i = blockIdx.x * NT_C + threadIdx.x;
  if (i < n_cam) {
    double rr[6], z[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) { rr[q] = rhs[6 * i + q]; x[6 * i + q] = 0.0; r[6 * i + q] = rr[q]; p[6 * i + q] = 0.0; sv[6 * i + q] = 0.0; }
    minv_apply(Minv + (size_t)36 * i, rr, z);
#pragma unroll
    for (int q = 0; q < 6; ++q) u[6 * i + q] = z[q];
    write_xtab(xtab + (size_t)XTAB * i, camtab + (size_t)CAMTAB * i, z);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    cg->g[0] = 0.0; cg->g[1] = 0.0; cg->a[0] = 0.0; cg->a[1] = 0.0; cg->gamma0 = 0.0; cg->tol = tol;
    cg->dot[0] = 0.0; cg->dot[1] = 0.0;
    cg->iters = 0; cg->max_iters = max_iters; cg->reason = 0; cg->done_at = INT_MAX;
  }
}

// w_i = (B_i + lam/radius) u_i - T_i' yhat_i (yhat_i summed here from the chunk partials of k_spmv_cm, in chunk order);
// partial gamma' = r.u and delta = u.w per CTA, summed in CTA order by the last CTA.
//   single GPU: the totals go to cg->dot.
//   OWNER (sharded, glba.cu "owner-computes"): this rank holds only the cameras its tracks observe; w is PARTIAL on cameras
//   other ranks observe too.  (B + lam) u and r.u are counted on the camera's owner rank only, u.w^rank on every rank
//   (it is linear in the partial products).  The partial rows of shared cameras and the two partial scalars are written
//   into the send buffer of the one all-reduce of the iteration.
template <bool OWNER>
__global__ void __launch_bounds__(NT_C)
k_cg_w(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ camtab, const double* __restrict__ Bc,
       const double* __restrict__ lamc, const double inv_radius, const int* __restrict__ cam_chunk_start, const double* __restrict__ part6,
       const double* __restrict__ r, const double* __restrict__ u, double* __restrict__ w, CgState* cg, const int li,
       const uint8_t* __restrict__ owned, const int* __restrict__ sh_of, double* __restrict__ xsend, const int n_shared,
       double* part, unsigned* counter) {
  pdl_grid_sync();
  __shared__ double sm[2 * NT_C / 32];
  __shared__ double smo[2];
  if (cg->done_at <= li) return;
  const int i = blockIdx.x * NT_C + threadIdx.x;
  double ru = 0.0, uw = 0.0;
  if (i < n_cam) {
    double wi[6] = {0, 0, 0, 0, 0, 0};
    if (cam_free[i]) {
      double yh[6] = {0, 0, 0, 0, 0, 0}, ty[6], ui[6];
      int ch = cam_chunk_start[i];
      const int ce = cam_chunk_start[i + 1];
      for (; ch < ce; ++ch)
#pragma unroll
        for (int a = 0; a < 6; ++a) yh[a] += part6[(size_t)6 * ch + a];
      apply_Tt(camtab + (size_t)CAMTAB * i, yh, ty);
      const bool mine = !OWNER || owned[i] != 0;
      const double* B = Bc + (size_t)36 * i;
#pragma unroll
      for (int a = 0; a < 6; ++a) ui[a] = u[6 * i + a];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        double sacc = 0.0;
        if (mine) {
          sacc = lamc[6 * i + a] * inv_radius * ui[a];
#pragma unroll
          for (int c = 0; c < 6; ++c) sacc += B[a * 6 + c] * ui[c];
          ru += r[6 * i + a] * ui[a];
        }
        sacc -= ty[a];
        wi[a] = sacc; uw += sacc * ui[a];
      }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) w[6 * i + a] = wi[a];
    if (OWNER) {
      const int sidx = sh_of[i];
      if (sidx >= 0) {
#pragma unroll
        for (int a = 0; a < 6; ++a) xsend[(size_t)6 * sidx + a] = wi[a];
      }
    }
  }
  const double v[2] = {ru, uw};
  const bool mx[2] = {false, false};
  const int slot[2] = {-1, -1};
  const bool last = finish_scalars<2>(v, mx, slot, part, counter, nullptr, sm, smo);
  if (last && threadIdx.x == 0) {
    if (OWNER) { xsend[(size_t)6 * n_shared] = smo[0]; xsend[(size_t)6 * n_shared + 1] = smo[1]; }
    else { cg->dot[0] = smo[0]; cg->dot[1] = smo[1]; }
  }
}

// the scalar recurrences and every vector update of the iteration; u = M^-1 r and the gather table for the next product
template <bool OWNER>
__global__ void __launch_bounds__(NT_C)
k_cg_update(const int n_cam, const double* __restrict__ camtab, const double* __restrict__ Minv, double* __restrict__ u, double* __restrict__ w,
            double* __restrict__ p, double* __restrict__ sv, double* __restrict__ x, double* __restrict__ r, double* __restrict__ xtab,
            CgState* cg, const int li, const int* __restrict__ sh_of, const double* __restrict__ xrecv, const int n_shared) {
  pdl_grid_sync();
  if (cg->done_at <= li) return;
  const double gp = OWNER ? xrecv[(size_t)6 * n_shared] : cg->dot[0];          // gamma' = r.u
  const double dl = OWNER ? xrecv[(size_t)6 * n_shared + 1] : cg->dot[1];      // delta  = u.S u
  const double g_old = cg->g[li & 1], a_old = cg->a[li & 1];
  const double g0 = (li == 0) ? gp : cg->gamma0;
  const bool first = (li == 0);
  // every CTA takes the same decisions from the same numbers; CTA 0 publishes them for the launches behind
  const bool zero_rhs = first && !(gp > 0.0);
  const bool converged = !first && sqrt(gp) <= cg->tol * sqrt(g0);
  const double beta = first ? 0.0 : ((g_old > 0.0) ? gp / g_old : 0.0);
  const double denom = first ? dl : dl - beta * gp / a_old;
  const bool breakdown = !(denom > 0.0);
  const bool stop_now = zero_rhs || converged || breakdown;
  const double alpha = stop_now ? 0.0 : gp / denom;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (first) cg->gamma0 = gp;
    cg->g[(li + 1) & 1] = gp; cg->a[(li + 1) & 1] = alpha;
    if (stop_now) { cg->done_at = li + 1; cg->reason = breakdown && !converged && !zero_rhs ? 2 : 1; cg->iters = li; }
    else { cg->iters = li + 1; if (li + 1 >= cg->max_iters) { cg->done_at = li + 1; cg->reason = 3; } }
  }
  if (stop_now) return;
  const int i = blockIdx.x * NT_C + threadIdx.x;
  if (i >= n_cam) return;
  double wi[6], rr[6], z[6];
  const int sidx = OWNER ? sh_of[i] : -1;
#pragma unroll
  for (int a = 0; a < 6; ++a) wi[a] = (sidx >= 0) ? xrecv[(size_t)6 * sidx + a] : w[6 * i + a];     // complete w on shared cameras
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    const double pn = u[6 * i + a] + beta * p[6 * i + a];
    const double sn = wi[a] + beta * sv[6 * i + a];
    p[6 * i + a] = pn; sv[6 * i + a] = sn;
    x[6 * i + a] += alpha * pn;
    rr[a] = r[6 * i + a] - alpha * sn; r[6 * i + a] = rr[a];
  }
  minv_apply(Minv + (size_t)36 * i, rr, z);
#pragma unroll
  for (int a = 0; a < 6; ++a) u[6 * i + a] = z[a];
  double t[6];
  apply_T(camtab + (size_t)CAMTAB * i, z, t);
  double* xr = xtab + (size_t)XTAB * i;
#pragma unroll
  for (int a = 0; a < 6; ++a) xr[a] = t[a];
}

// candidate cameras cam_c = cam - y, their table, gather table [T y | R | sv] for the back-substitution,
// camera parts of |y|^2, y.g, y' Lambda y
__global__ void __launch_bounds__(NT_C)
k_cam_step2(const int n_cam, const uint8_t* __restrict__ cam_free, const double* __restrict__ cam, const double* __restrict__ camtab,
            const double* __restrict__ y, const double* __restrict__ gc, const double* __restrict__ lamc, const double inv_radius_arg,
            double* __restrict__ cam_c, double* __restrict__ camtab_c, double* __restrict__ xtab, double* part, unsigned* counter, double* scal,
            const int mode, const LmCtl* __restrict__ ctl = nullptr, const uint8_t* __restrict__ owned = nullptr) {
  pdl_grid_sync();
  __shared__ double sm[3 * NT_C / 32];
  if (ctl_skip(ctl, GATE_ALWAYS)) return;
  const double inv_radius = ctl_inv_radius(ctl, inv_radius_arg);
  __shared__ double smo[3];
  const int i = blockIdx.x * NT_C + threadIdx.x;
  double yn2 = 0.0, ygd = 0.0, yly = 0.0;
  if (i < n_cam) {
    double yi[6], cc[6];
    const bool f = cam_free[i] != 0;
    const bool mine = owned == nullptr || owned[i] != 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      yi[a] = f ? y[6 * i + a] : 0.0;
      cc[a] = cam[6 * i + a] - yi[a];
      if (mine) { yn2 += yi[a] * yi[a]; ygd += yi[a] * gc[6 * i + a]; yly += lamc[6 * i + a] * inv_radius * yi[a] * yi[a]; }
    }
    const double* ct = camtab + (size_t)CAMTAB * i;
    if (mode == 0) {
      cam_table_row(cc, camtab_c + (size_t)CAMTAB * i);
    } else if (!f) {              // constant vertex: state and table row carried over bit for bit
#pragma unroll
      for (int a = 0; a < 6; ++a) cc[a] = cam[6 * i + a];
      for (int q = 0; q < CAMTAB; ++q) camtab_c[(size_t)CAMTAB * i + q] = ct[q];
    } else {
      // g2o oplus: the step is x = -y with dw = x_w, dv = -R x_t;  T <- exp([dw, dv]) T  (SE3Quat::exp: t' = V dv)
      double dw[3], dv[3], dR[9], Rn[9], tn[3], B;
#pragma unroll
      for (int r = 0; r < 3; ++r) { dw[r] = -yi[r]; dv[r] = ct[r * 3] * yi[3] + ct[r * 3 + 1] * yi[4] + ct[r * 3 + 2] * yi[5]; }
      so3_exp(dw, dR, &B);
      const double th2 = dw[0] * dw[0] + dw[1] * dw[1] + dw[2] * dw[2];
      double Cc;
      if (th2 < 1e-8) Cc = 1.0 / 6.0 - th2 * (1.0 / 120.0) + th2 * th2 * (1.0 / 5040.0);
      else { const double th = sqrt(th2); Cc = (th - sin(th)) / (th2 * th); }
      const double a1[3] = {dw[1] * dv[2] - dw[2] * dv[1], dw[2] * dv[0] - dw[0] * dv[2], dw[0] * dv[1] - dw[1] * dv[0]};
      const double a2[3] = {dw[1] * a1[2] - dw[2] * a1[1], dw[2] * a1[0] - dw[0] * a1[2], dw[0] * a1[1] - dw[1] * a1[0]};
      const double* t = cam + 6 * i + 3;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) Rn[r * 3 + c] = dR[r * 3] * ct[c] + dR[r * 3 + 1] * ct[3 + c] + dR[r * 3 + 2] * ct[6 + c];
        tn[r] = dR[r * 3] * t[0] + dR[r * 3 + 1] * t[1] + dR[r * 3 + 2] * t[2] + dv[r] + B * a1[r] + Cc * a2[r];
      }
      so3_log(Rn, cc);
      cc[3] = tn[0]; cc[4] = tn[1]; cc[5] = tn[2];
      cam_table_row_Rt(Rn, tn, camtab_c + (size_t)CAMTAB * i);
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) cam_c[6 * i + a] = cc[a];
    write_xtab(xtab + (size_t)XTAB * i, ct, yi);
  }
  const double v[3] = {yn2, ygd, yly};
  const bool mx[3] = {false, false, false};
  const int slot[3] = {S_YN2_C, S_YG_C, S_YLY_C};
  finish_scalars<3>(v, mx, slot, part, counter, scal, sm, smo);
}

}  // namespace glba
