// glba_tiles.cuh — point-major kernels, tiled: the "warp-level segmented reduction over point tracks
// with shared-memory staging" of the north star, done per CTA.
//
// A tile is a run of whole point tracks holding < NT_T observations (built at load time from the
// track CSR: tile t owns the points whose first observation index lies in [t*B, (t+1)*B),
// B = NT_T - longest track).  Phase 1: one thread per OBSERVATION — every per-observation array is read
// with unit stride (coalesced), the per-observation contribution is staged in shared memory (SoA).
// Phase 2: one thread per POINT sums its segment in observation order (fixed order => bit-reproducible),
// then does the 3x3 work (damped inverse, Cinv t, back-substitution).
// The thread-per-point kernels in glba_kernels.cuh stay as the fallback for tracks longer than NT_T/2.
#pragma once
#include "glba_kernels.cuh"
#include "glba_lm.cuh"

namespace glba {

#ifndef GLBA_NT_T
#define GLBA_NT_T 128
#endif
constexpr int NT_T = GLBA_NT_T;
// 128 threads per tile CTA: measured on C4 against 64 / 256 / 512 threads at 2..8 observations per thread, 128 x 4 is the
// best shape for all three tile kernels (k_linearize_tile 0.219 vs 0.226 ms, k_point_tile<0> 0.077 vs 0.081, <1> 0.162 vs
// 0.169 for 256 x 4): more, smaller CTAs per SM interleave their barrier-separated phases better.
// Observations per thread in phase 1 (template parameter OPT): 4 for large maps (tile = 512 observations, ~100 points),
// 1 for small ones (tile = 128 observations: more CTAs and a 4x shorter dependent chain per thread — what matters when
// the whole map is a few tiles).
#ifndef GLBA_OPT_LARGE
#define GLBA_OPT_LARGE 4
#endif
constexpr int OPT_LARGE = GLBA_OPT_LARGE, OPT_SMALL = 1;

// The last CTA to finish folds the per-CTA partial rows [rows][5] into the scalar slots, in row order (fixed).
// MAXCOL = column reduced with max (-1: none).  Saves a separate single-CTA reduction launch per pass.
template <int MAXCOL, int NT = GLBA_NT_T>
__device__ __forceinline__ void last_block_reduce5(const double* part, const int rows, const int* slots, unsigned* counter,
                                                   double* scal, double* sm /* 5*NT/32 */, double* smo /* 5 */,
                                                   const LmHook* hook = nullptr) {
  __shared__ bool s_last;
  if (counter == nullptr) return;       // large grids: the host launches k_reduce_rows instead (uniform branch)
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned t = atomicInc(counter, gridDim.x - 1);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  double mx[1] = {0.0};
#pragma unroll 4
  for (int r = threadIdx.x; r < rows; r += NT) {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double x = __ldcg(part + (size_t)5 * r + q);
      if (q == MAXCOL) mx[0] = fmax(mx[0], x); else acc[q] += x;
    }
  }
  block_reduce<5, NT>(acc, sm, smo);
  if (threadIdx.x < 5 && (int)threadIdx.x != MAXCOL) scal[slots[threadIdx.x]] = smo[threadIdx.x];
  if constexpr (MAXCOL >= 0) {
    __syncthreads();
    block_reduce<1, NT, true>(mx, sm, smo);
    if (threadIdx.x == 0) scal[slots[MAXCOL]] = smo[0];
  } else if (hook != nullptr && hook->ctl != nullptr) {
    // back-substitution pass of a device-resident LM loop: its scalars are final, take the trust-region decision
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); lm_decide(hook->ctl, hook->P, scal, hook->sum); }
  }
}
struct RedArgs {
  unsigned* counter; double* scal; int slots[5];
  const LmCtl* ctl; int gate;      // device-resident LM loop: skip / radius (glba_kernels.cuh)
  LmHook hook;                     // ... and the decision to take once the scalars of this pass are final (MAXCOL == -1 passes)
};

// Large grids: 64 CTAs each fold a fixed, contiguous range of the per-tile partial rows, the last one folds the 64.
template <int MAXCOL>
__global__ void __launch_bounds__(NT_T)
k_reduce_rows(const double* __restrict__ part, const int rows, double* __restrict__ part2 /* [grid][5] */, const RedArgs RA) {
  pdl_grid_sync();
  __shared__ double sm[5 * NT_T / 32];
  __shared__ double smo[5];
  if (ctl_skip(RA.ctl, RA.gate)) return;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  double mx[1] = {0.0};
  for (int r = r0 + threadIdx.x; r < r1; r += NT_T) {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double x = __ldg(part + (size_t)5 * r + q);
      if (q == MAXCOL) mx[0] = fmax(mx[0], x); else acc[q] += x;
    }
  }
  block_reduce<5, NT_T>(acc, sm, smo);
  if (threadIdx.x < 5 && (int)threadIdx.x != MAXCOL) part2[(size_t)5 * blockIdx.x + threadIdx.x] = smo[threadIdx.x];
  if constexpr (MAXCOL >= 0) {
    __syncthreads();
    block_reduce<1, NT_T, true>(mx, sm, smo);
    if (threadIdx.x == 0) part2[(size_t)5 * blockIdx.x + MAXCOL] = smo[0];
  }
  __syncthreads();
  last_block_reduce5<MAXCOL>(part2, gridDim.x, RA.slots, RA.counter, RA.scal, sm, smo, &RA.hook);
}

constexpr int CWIN = 32;    // camera rows staged in shared memory per tile: [cmin, cmin + CWIN)
constexpr int XROW = 18;    // padded row stride (doubles) of the staged gather rows: spreads rows over the banks
constexpr int CROW = 14;    // padded stride of staged candidate rows (R, c)
struct TileArgs {
  const int* tile_pt;      // [n_tiles+1] first point of every tile
  const int* pm_pt;        // point of observation k (point-major order)
  const int* tile_cmin;    // [n_tiles] smallest camera index observed in the tile
  int n_cam;
};

// K_A + point half of K_B, tiled.  Same outputs as k_linearize_pm.
template <int OPT>
__global__ void __launch_bounds__(NT_T)
k_linearize_tile(const PmArgs A, const TileArgs T, const double4* __restrict__ pt, const double* __restrict__ camtab,
                 double4* __restrict__ rec_pm, double4* __restrict__ rec_cm, double* __restrict__ Craw, double4* __restrict__ sp4,
                 double4* __restrict__ lam4, double* __restrict__ cinv, double4* __restrict__ u0p, const int first, const int jacobi, const double min_diag,
                 const double max_diag, const double inv_radius_arg, double* __restrict__ part /* [grid][5] */, const RedArgs RA) {
  pdl_grid_sync();
  constexpr int TILE_OBS = NT_T * OPT;
  extern __shared__ double dsm[];
  if (ctl_skip(RA.ctl, RA.gate)) return;
  const double inv_radius = ctl_inv_radius(RA.ctl, inv_radius_arg);
  double (*val)[TILE_OBS] = reinterpret_cast<double (*)[TILE_OBS]>(dsm);     // [8][TILE_OBS]: J~p rows (3+3), r~ (2)
  __shared__ double sm[4 * NT_T / 32];
  __shared__ double smo[4];
  __shared__ double smm[NT_T / 32];
  __shared__ double smmo[1];
  const int tid = threadIdx.x;
  const int j0 = T.tile_pt[blockIdx.x], j1 = T.tile_pt[blockIdx.x + 1];
  const int k0 = A.pt_start[j0], k1 = A.pt_start[j1];
  double cost = 0.0, bad = 0.0;
  // camera rows (R, centre) of the tile's camera window staged once, coalesced (see k_point_tile)
  __shared__ __align__(16) double cs[CWIN * CROW];
  const int cmin = T.tile_cmin[blockIdx.x];
  for (int t = tid; t < CWIN * 6; t += NT_T) {
    const int row = t / 6, part = t - row * 6, cam = cmin + row;
    if (cam < T.n_cam) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(camtab + (size_t)CAMTAB * cam) + part);
      *reinterpret_cast<double2*>(&cs[row * CROW + 2 * part]) = v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < OPT; ++m) {
    const int l = m * NT_T + tid;
    const int k = k0 + l;
    if (k < k1) {
      const int ci = __ldg(A.pm_cam + k);
      const int pj = __ldg(T.pm_pt + k);
      const double2 uv = __ldg(A.pm_uv + k);
      const double4 X = ldg4(pt + pj);
      double R[9], cc[3];
      const unsigned off = (unsigned)(ci - cmin);
      if (off < (unsigned)CWIN) {
        const double2* cr = reinterpret_cast<const double2*>(&cs[off * CROW]);
        const double2 a0 = cr[0], a1 = cr[1], a2 = cr[2], a3 = cr[3], a4 = cr[4], a5 = cr[5];
        R[0] = a0.x; R[1] = a0.y; R[2] = a1.x; R[3] = a1.y; R[4] = a2.x; R[5] = a2.y; R[6] = a3.x; R[7] = a3.y; R[8] = a4.x;
        cc[0] = a4.y; cc[1] = a5.x; cc[2] = a5.y;
      } else {
        load_Rc(camtab + (size_t)CAMTAB * ci, R, cc);
      }
      const double qx = X.x - cc[0], qy = X.y - cc[1], qz = X.z - cc[2];
      const double px = R[0] * qx + R[1] * qy + R[2] * qz;
      const double py = R[3] * qx + R[4] * qy + R[5] * qz;
      const double pz = R[6] * qx + R[7] * qy + R[8] * qz;
      const double iz = 1.0 / pz;
      const double xh = px * iz, yh = py * iz;
      const double rx = A.K.fx * xh + A.K.cx - uv.x, ry = A.K.fy * yh + A.K.cy - uv.y;
      double rho, w;
      loss_eval(A.loss, (X.w * X.w) * (rx * rx + ry * ry), rho, w);      // X.w = sqrt(information) of the point, 1 by default
      w *= X.w;
      if (!isfinite(rx) || !isfinite(ry)) bad += 1.0;
      cost += 0.5 * rho;
      const double4 rec = make_double4(xh, yh, iz, w);
      st4(rec_pm + k, rec);
      st4(rec_cm + __ldg(A.pm2cm + k), rec);
      double ap[3], bp[3];
      jp_rows(rec, R, A.K, ap, bp);
      const double r0 = w * rx, r1 = w * ry;
      // stage the two rows of J~p and r~ (8 values, 64 KB per tile: three CTAs per SM still fit beside the camera window)
      val[0][l] = ap[0]; val[1][l] = ap[1]; val[2][l] = ap[2]; val[3][l] = bp[0]; val[4][l] = bp[1]; val[5][l] = bp[2];
      val[6][l] = r0; val[7][l] = r1;
    }
  }
  __syncthreads();
  double xn2 = 0.0, gmax = 0.0, notpd = 0.0;
  for (int j = j0 + tid; j < j1; j += NT_T) {
    const int b = __ldg(A.pt_start + j) - k0, e = __ldg(A.pt_start + j + 1) - k0;
    // the point's global operands are requested before the shared-memory summation so their latency hides behind it
    const bool free_pt = A.pt_free[j] != 0;
    const double4 Xp = ldg4(pt + j);
    double4 s4 = make_double4(1.0, 1.0, 1.0, 0.0);
    if (!first && free_pt) s4 = ldg4(sp4 + j);
    double C[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
    for (int m = b; m < e; ++m) {
      const double a0 = val[0][m], a1 = val[1][m], a2 = val[2][m], b0 = val[3][m], b1 = val[4][m], b2 = val[5][m];
      const double r0 = val[6][m], r1 = val[7][m];
      C[0] += a0 * a0 + b0 * b0; C[1] += a0 * a1 + b0 * b1; C[2] += a0 * a2 + b0 * b2;
      C[3] += a1 * a1 + b1 * b1; C[4] += a1 * a2 + b1 * b2; C[5] += a2 * a2 + b2 * b2;
      g[0] += a0 * r0 + b0 * r1; g[1] += a1 * r0 + b1 * r1; g[2] += a2 * r0 + b2 * r1;
    }
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
        s[0] = s4.x; s[1] = s4.y; s[2] = s4.z;
      }
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const double s2 = s[q] * s[q];
        lam[q] = fmin(fmax(s2 * h[q], min_diag), max_diag) / s2;
      }
      st4(lam4 + j, make_double4(lam[0], lam[1], lam[2], 0.0));
      if (!point_block(C, g, lam, inv_radius, blk)) notpd += 1.0;
      xn2 += Xp.x * Xp.x + Xp.y * Xp.y + Xp.z * Xp.z;
      gmax = fmax(gmax, fmax(fabs(g[0]), fmax(fabs(g[1]), fabs(g[2]))));
    } else {
#pragma unroll
      for (int q = 0; q < PBLK; ++q) blk[q] = 0.0;
      if (first) st4(sp4 + j, make_double4(1.0, 1.0, 1.0, 0.0));
      st4(lam4 + j, make_double4(0.0, 0.0, 0.0, 0.0));
    }
    store_pblk(cinv, u0p, j, blk);
  }
  double v[4] = {cost, xn2, bad, notpd};
  block_reduce<4, NT_T>(v, sm, smo);
  double m[1] = {gmax};
  block_reduce<1, NT_T, true>(m, smm, smmo);
  if (tid == 0) {
    double* p = part + (size_t)5 * blockIdx.x;
    p[0] = smo[0]; p[1] = smo[1]; p[2] = smo[2]; p[3] = smo[3]; p[4] = smmo[0];
  }
  __shared__ double rsm[5 * NT_T / 32];
  __shared__ double rsmo[5];
  last_block_reduce5<4>(part, gridDim.x, RA.slots, RA.counter, RA.scal, rsm, rsmo);
}

// Point-major half of the implicit product (MODE 0) / back-substitution + candidate cost (MODE 1), tiled.
// xtab row of camera i: [xg(6) = T_i x_i | R_i (9) | sv_i (3)] — one contiguous 144-byte gather.
template <int MODE, int OPT>
__global__ void __launch_bounds__(NT_T)
k_point_tile(const PmArgs A, const TileArgs T, const double4* __restrict__ rec_pm, const double* __restrict__ camtab,
             const double* __restrict__ xtab,
             const double* __restrict__ cinv, const double4* __restrict__ u0p, double4* __restrict__ u4, const CgState* __restrict__ cg, const int li,
             // MODE 1 only:
             const double4* __restrict__ pt, double4* pt_c, const double* __restrict__ camtab_c, const double* __restrict__ Craw,
             const double4* __restrict__ lam4, const double inv_radius_arg, double* __restrict__ part /* [grid][5] */, const RedArgs RA) {
  pdl_grid_sync();
  constexpr int TILE_OBS = NT_T * OPT;
  __shared__ double val[3][TILE_OBS];
  if (MODE == 1 && ctl_skip(RA.ctl, RA.gate)) return;
  const double inv_radius = (MODE == 1) ? ctl_inv_radius(RA.ctl, inv_radius_arg) : inv_radius_arg;
  __shared__ __align__(16) double xs[CWIN * XROW];                     // staged gather rows of cameras [cmin, cmin+CWIN)
  __shared__ __align__(16) double cs[MODE == 1 ? CWIN * CROW : 2];     // staged candidate rows (R, c)
  if (MODE == 0 && cg && cg->done_at <= li) return;
  const int tid = threadIdx.x;
  const int j0 = T.tile_pt[blockIdx.x], j1 = T.tile_pt[blockIdx.x + 1];
  const int k0 = A.pt_start[j0], k1 = A.pt_start[j1];
  const int cmin = T.tile_cmin[blockIdx.x];
  // creation-ordered map points make a tile touch a narrow camera range: stage those rows once (coalesced), so the
  // per-observation gathers below are shared-memory reads instead of L2 round trips; cameras outside the window
  // (loop-closure wrap-around) fall back to the global gather
  for (int t = tid; t < CWIN * 8; t += NT_T) {
    const int row = t >> 3, part = t & 7, cam = cmin + row;
    if (cam < T.n_cam) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(xtab + (size_t)XTAB * cam) + part);
      *reinterpret_cast<double2*>(&xs[row * XROW + 2 * part]) = v;
    }
  }
  if (MODE == 1)
    for (int t = tid; t < CWIN * 6; t += NT_T) {
      const int row = t / 6, part = t - row * 6, cam = cmin + row;
      if (cam < T.n_cam) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(camtab_c + (size_t)CAMTAB * cam) + part);
        *reinterpret_cast<double2*>(&cs[row * CROW + 2 * part]) = v;
      }
    }
  __syncthreads();
#pragma unroll
  for (int m = 0; m < OPT; ++m) {
    const int l = m * NT_T + tid;
    const int k = k0 + l;
    if (k < k1) {
      const int cam_i = __ldg(A.pm_cam + k);
      const double4 rec = ldg4(rec_pm + k);
      double xg[6], R[9], flag;
      const unsigned off = (unsigned)(cam_i - cmin);
      if (off < (unsigned)CWIN) {
        const double2* xr = reinterpret_cast<const double2*>(&xs[off * XROW]);
        const double2 a0 = xr[0], a1 = xr[1], a2 = xr[2], a3 = xr[3], a4 = xr[4], a5 = xr[5], a6 = xr[6], a7 = xr[7];
        xg[0] = a0.x; xg[1] = a0.y; xg[2] = a1.x; xg[3] = a1.y; xg[4] = a2.x; xg[5] = a2.y;
        R[0] = a3.x; R[1] = a3.y; R[2] = a4.x; R[3] = a4.y; R[4] = a5.x; R[5] = a5.y; R[6] = a6.x; R[7] = a6.y; R[8] = a7.x;
        flag = a7.y;
      } else {        // one 128-byte row per camera: four 256-bit gathers
        const double4* xr = reinterpret_cast<const double4*>(xtab + (size_t)XTAB * cam_i);
        const double4 x0 = ldg4(xr), x1 = ldg4(xr + 1), x2 = ldg4(xr + 2), x3 = ldg4(xr + 3);
        xg[0] = x0.x; xg[1] = x0.y; xg[2] = x0.z; xg[3] = x0.w; xg[4] = x1.x; xg[5] = x1.y;
        R[0] = x1.z; R[1] = x1.w; R[2] = x2.x; R[3] = x2.y; R[4] = x2.z; R[5] = x2.w; R[6] = x3.x; R[7] = x3.y; R[8] = x3.z;
        flag = x3.w;
      }
      double sv0 = 0.0, sv1 = 0.0, sv2 = 0.0;
      if (flag != 0.0) {               // Ceres' small-angle branch: only an (almost) exactly-identity keyframe
        const double* ct = camtab + (size_t)CAMTAB * cam_i;
        sv0 = ct[CT_SV]; sv1 = ct[CT_SV + 1]; sv2 = ct[CT_SV + 2];
      }
      double a[6], bb[6];
      jhat_rows(rec, sv0, sv1, sv2, A.K, a, bb);
      double al0 = 0.0, al1 = 0.0;
#pragma unroll
      for (int r = 0; r < 6; ++r) { al0 += a[r] * xg[r]; al1 += bb[r] * xg[r]; }
      double ap[3], bp[3];
      jp_rows(rec, R, A.K, ap, bp);
      val[0][l] = ap[0] * al0 + bp[0] * al1;
      val[1][l] = ap[1] * al0 + bp[1] * al1;
      val[2][l] = ap[2] * al0 + bp[2] * al1;
    }
  }
  __syncthreads();
  double yn2 = 0.0, yg = 0.0, yly = 0.0;
  for (int j = j0 + tid; j < j1; j += NT_T) {
    const int b = __ldg(A.pt_start + j) - k0, e = __ldg(A.pt_start + j + 1) - k0;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    const bool free_pt = A.pt_free[j] != 0;
    if (free_pt)
      for (int m = b; m < e; ++m) { t0 += val[0][m]; t1 += val[1][m]; t2 += val[2][m]; }
    double Ci[6], u0[3] = {0.0, 0.0, 0.0};
    if (MODE == 1) load_pblk(cinv, u0p, j, Ci, u0); else load_cinv(cinv, j, Ci);
    const double v0 = Ci[0] * t0 + Ci[1] * t1 + Ci[2] * t2;
    const double v1 = Ci[1] * t0 + Ci[3] * t1 + Ci[4] * t2;
    const double v2 = Ci[2] * t0 + Ci[4] * t1 + Ci[5] * t2;
    if (MODE == 0) {
      st4(u4 + j, make_double4(v0, v1, v2, 0.0));
    } else {
      const double y0 = u0[0] - v0, y1 = u0[1] - v1, y2 = u0[2] - v2;
      const double4 X = ldg4(pt + j);
      st4(pt_c + j, make_double4(X.x - y0, X.y - y1, X.z - y2, X.w));
      if (free_pt) {
        const double4 l4 = ldg4(lam4 + j);
        yn2 += y0 * y0 + y1 * y1 + y2 * y2;
        yg += y0 * Craw[(size_t)6 * A.n_pt + j] + y1 * Craw[(size_t)7 * A.n_pt + j] + y2 * Craw[(size_t)8 * A.n_pt + j];
        yly += (l4.x * y0 * y0 + l4.y * y1 * y1 + l4.z * y2 * y2) * inv_radius;
      }
    }
  }
  if (MODE == 1) {
    __shared__ double sm[5 * NT_T / 32];
    __shared__ double smo[5];
    __syncthreads();                  // pt_c of this tile is visible to the whole CTA
    double cost_c = 0.0, bad = 0.0;
#pragma unroll
    for (int m = 0; m < OPT; ++m) {
      const int k = k0 + m * NT_T + tid;
      if (k < k1) {
        const int cam_i = __ldg(A.pm_cam + k);
        const int pt_j = __ldg(T.pm_pt + k);
        const double2 uv = __ldg(A.pm_uv + k);
        const double4 xc = ld4(pt_c + pt_j);          // written by this CTA: coherent load
        double Rc[9], cc[3];
        const unsigned off = (unsigned)(cam_i - cmin);
        if (off < (unsigned)CWIN) {
          const double2* cr = reinterpret_cast<const double2*>(&cs[off * CROW]);
          const double2 a0 = cr[0], a1 = cr[1], a2 = cr[2], a3 = cr[3], a4 = cr[4], a5 = cr[5];
          Rc[0] = a0.x; Rc[1] = a0.y; Rc[2] = a1.x; Rc[3] = a1.y; Rc[4] = a2.x; Rc[5] = a2.y; Rc[6] = a3.x; Rc[7] = a3.y; Rc[8] = a4.x;
          cc[0] = a4.y; cc[1] = a5.x; cc[2] = a5.y;
        } else {
          load_Rc(camtab_c + (size_t)CAMTAB * cam_i, Rc, cc);
        }
        const double qx = xc.x - cc[0], qy = xc.y - cc[1], qz = xc.z - cc[2];
        const double px = Rc[0] * qx + Rc[1] * qy + Rc[2] * qz, py = Rc[3] * qx + Rc[4] * qy + Rc[5] * qz, pz = Rc[6] * qx + Rc[7] * qy + Rc[8] * qz;
        const double iz = 1.0 / pz;
        const double rx = A.K.fx * (px * iz) + A.K.cx - uv.x, ry = A.K.fy * (py * iz) + A.K.cy - uv.y;
        double rho, w;
        loss_eval(A.loss, (xc.w * xc.w) * (rx * rx + ry * ry), rho, w);
        if (!isfinite(rx) || !isfinite(ry)) bad += 1.0;
        cost_c += 0.5 * rho;
      }
    }
    double v[5] = {cost_c, yn2, yg, yly, bad};
    block_reduce<5, NT_T>(v, sm, smo);
    if (tid < 5) part[(size_t)5 * blockIdx.x + tid] = smo[tid];
    __syncthreads();
    last_block_reduce5<-1>(part, gridDim.x, RA.slots, RA.counter, RA.scal, sm, smo, &RA.hook);
  }
}

// smallest camera index among a tile's observations
__global__ void __launch_bounds__(NT_T)
k_tile_cmin(const int* __restrict__ tile_pt, const int* __restrict__ pt_start, const int* __restrict__ pm_cam, int* __restrict__ tile_cmin) {
  pdl_grid_sync();
  __shared__ int sm[NT_T / 32];
  const int j0 = tile_pt[blockIdx.x], j1 = tile_pt[blockIdx.x + 1];
  const int k0 = pt_start[j0], k1 = pt_start[j1];
  int m = 0x7fffffff;
  for (int k = k0 + threadIdx.x; k < k1; k += NT_T) m = min(m, pm_cam[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < NT_T / 32; ++w) m = min(m, sm[w]); tile_cmin[blockIdx.x] = (m == 0x7fffffff) ? 0 : m; }
}

// ---- load-time helpers ------------------------------------------------------------------------
__global__ void k_max_track(const int n_pt, const int* __restrict__ pt_start, int* out) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  int len = (j < n_pt) ? pt_start[j + 1] - pt_start[j] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if ((threadIdx.x & 31) == 0 && len > 0) atomicMax(out, len);
}
// tile_pt[t] = first point whose first observation index is >= t*B  (t = 0..n_tiles), clamped to n_pt
__global__ void k_tile_starts(const int n_tiles, const int B, const int n_pt, const int* __restrict__ pt_start, int* __restrict__ tile_pt) {
  pdl_grid_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  if (t == n_tiles) { tile_pt[t] = n_pt; return; }
  const long target = (long)t * B;
  int lo = 0, hi = n_pt;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (pt_start[mid] < target) lo = mid + 1; else hi = mid; }
  tile_pt[t] = lo;
}

}  // namespace glba
