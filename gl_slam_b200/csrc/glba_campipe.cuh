// glba_campipe.cuh — both camera-major passes of a linearisation in one kernel, fed by 1-D TMA bulk copies.
//
// k_linearize_cm (A_i = sum J^'J^, ghat_i = sum J^'r~) and k_schur_cm (Mhat_i = sum J^' E J^, rhat_i = sum J^' f) stream the
// same 32-byte records of a camera's chunk; run back to back they read them twice (2 x 160 MB on C4) and each waits on its
// own loads (ncu: 66 % / 64 % long-scoreboard stalls at 41 % / 31 % of the DRAM peak).  Here one CTA per chunk stages the
// chunk in 512-observation sub-tiles through a two-stage shared-memory ring (cp.async.bulk + mbarrier, one elected thread)
// and two groups of threads consume every staged record ONCE each: group A (3 warps) accumulates the Hessian half, group M
// (5 warps) gathers the damped inverse point block and accumulates the Schur half — the Schur half costs about twice the
// arithmetic per observation and waits on a gather, so with equal groups group A spent a third of the kernel at the
// sub-tile barrier (ncu).  Fixed-order sums (lane-strided per thread, warp butterfly, warp order); k_linearize_cm /
// k_schur_cm remain for small maps, for a linearisation without Schur pieces and for re-damping after a rejected step.
#pragma once
#include "glba_kernels.cuh"
#include "glba_pipe.cuh"

namespace glba {

constexpr int CP_NA = 96, CP_NM = 160;      // threads of group A (Hessian half) and group M (Schur half)
constexpr int CP_SUB = 480;                 // observations per sub-tile: 5 per thread of group A, 3 per thread of group M
constexpr int CP_NT = CP_NA + CP_NM;
struct CamStage {
  double4 rec[CP_SUB];
  double2 uv[CP_SUB];
  int pt[CP_SUB + 8];                       // the slice starts at the 16-byte boundary at or below its first element
};
struct CamSmem {
  CamStage st[2];
  double red[2][27 * (CP_NM / 32)];
  double redo[2][27];
  uint64_t full[2];
};
static_assert(sizeof(CamSmem) <= 56 * 1024, "up to four CTAs of k_cam_pipe per SM");
static_assert(CP_SUB % CP_NA == 0 && CP_SUB % CP_NM == 0 && CP_NA % 32 == 0 && CP_NM % 32 == 0, "whole trips, whole warps");

__device__ __forceinline__ void cam_issue(CamStage& S, uint64_t* bar, const CmArgs& A, const double4* __restrict__ rec_cm, const int k0, const int k1) {
  const unsigned n = (unsigned)(k1 - k0);
  mbar_expect_tx(bar, n * 32u + n * 16u + slice_bytes(A.cm_pt, 4, k0, k1));
  bulk_g2s(S.rec, rec_cm + k0, n * 32u, bar);
  bulk_g2s(S.uv, A.cm_uv + k0, n * 16u, bar);
  bulk_slice(S.pt, A.cm_pt, 4, k0, k1, bar);
}

template <int MINB>       // resident CTAs per SM the register allocation aims at (2: 128 registers, 3: 85)
__global__ void __launch_bounds__(CP_NT, MINB)
k_cam_pipe(const CmArgs A, const double4* __restrict__ rec_cm, const double* __restrict__ camtab, const double* __restrict__ cinv,
           const double4* __restrict__ u0p, double* __restrict__ part_lin /* [n_chunks][27] */, double* __restrict__ part_schur /* [n_chunks][27] */) {
  pdl_grid_sync();
  extern __shared__ __align__(128) unsigned char cp_raw[];
  CamSmem& S = *reinterpret_cast<CamSmem*>(cp_raw);
  const int tid = threadIdx.x;
  const int grp = tid < CP_NA ? 0 : 1, t = tid - (grp ? CP_NA : 0);
  const int ch = blockIdx.x;
  const int cam = A.chunk_cam[ch];
  const int b = A.chunk_begin[ch], e = A.chunk_end[ch];
  const bool fr = A.cam_free[cam] != 0;                 // uniform per CTA
  const int n_sub = fr ? (e - b + CP_SUB - 1) / CP_SUB : 0;
  if (tid == 0) {
    mbar_init(&S.full[0], 1); mbar_init(&S.full[1], 1); mbar_fence_init();
    if (n_sub > 0) cam_issue(S.st[0], &S.full[0], A, rec_cm, b, min(e, b + CP_SUB));
    if (n_sub > 1) cam_issue(S.st[1], &S.full[1], A, rec_cm, b + CP_SUB, min(e, b + 2 * CP_SUB));
  }
  double acc[27];
#pragma unroll
  for (int q = 0; q < 27; ++q) acc[q] = 0.0;
  const double* ct = camtab + (size_t)CAMTAB * cam;
  double R[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) R[q] = ct[q];
  const double sv0 = ct[21], sv1 = ct[22], sv2 = ct[23];
  __syncthreads();                                       // barrier objects initialised before anyone waits on them
  for (int s = 0; s < n_sub; ++s) {
    CamStage& T = S.st[s & 1];
    const int k0 = b + s * CP_SUB, cnt = min(e - k0, CP_SUB);
    mbar_wait(&S.full[s & 1], (unsigned)((s >> 1) & 1));
    if (grp == 0) {
      // Hessian half: the arithmetic of k_linearize_cm
#pragma unroll
      for (int j = 0; j < CP_SUB / CP_NA; ++j) {
        const int l = j * CP_NA + t;
        if (l < cnt) {
          const double4 rec = T.rec[l];
          const double2 uv = T.uv[l];
          double a[6], bb[6];
          jhat_rows(rec, sv0, sv1, sv2, A.K, a, bb);
          const double r0 = rec.w * (A.K.fx * rec.x + A.K.cx - uv.x);
          const double r1 = rec.w * (A.K.fy * rec.y + A.K.cy - uv.y);
          acc_sym_sparse(acc, a, bb, a, bb);
          acc_vec_sparse(acc + 21, a, bb, r0, r1);
        }
      }
    } else {
      // Schur half: the arithmetic of k_schur_cm.  This group is the critical path (ncu: it waits on the gathered inverse point
      // blocks while group A waits at the barrier), so the block of the NEXT trip is requested before this trip's arithmetic.
      const int o_pt = slice_off(A.cm_pt, 4, k0);
      double Cn[6] = {0, 0, 0, 0, 0, 0}, un[3] = {0, 0, 0};
      if (t < cnt) load_pblk(cinv, u0p, T.pt[o_pt + t], Cn, un);
#pragma unroll
      for (int j = 0; j < CP_SUB / CP_NM; ++j) {
        const int l = j * CP_NM + t;
        double Ci[6], u0[3];
#pragma unroll
        for (int q = 0; q < 6; ++q) Ci[q] = Cn[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) u0[q] = un[q];
        if (j + 1 < CP_SUB / CP_NM && l + CP_NM < cnt) load_pblk(cinv, u0p, T.pt[o_pt + l + CP_NM], Cn, un);
        if (l < cnt) {
          const double4 rec = T.rec[l];
          double ap[3], bp[3];
          jp_rows(rec, R, A.K, ap, bp);
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
    }
    __syncthreads();                                     // both groups are done with this stage
    if (tid == 0 && s + 2 < n_sub) cam_issue(T, &S.full[s & 1], A, rec_cm, k0 + 2 * CP_SUB, min(e, k0 + 3 * CP_SUB));
  }
  // per group: warp butterfly, then the group's warp sums in warp order
  {
    const int lane = t & 31, wid = t >> 5;
#pragma unroll
    for (int i = 0; i < 27; ++i) {
      double x = acc[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
      if (lane == 0) S.red[grp][wid * 27 + i] = x;
    }
    __syncthreads();
    if (t < 27) {
      double x = S.red[grp][t];
      for (int w = 1; w < (grp ? CP_NM : CP_NA) / 32; ++w) x += S.red[grp][w * 27 + t];
      (grp == 0 ? part_lin : part_schur)[(size_t)27 * ch + t] = x;
    }
  }
}

}  // namespace glba
