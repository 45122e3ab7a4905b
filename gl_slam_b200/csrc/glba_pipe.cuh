// glba_pipe.cuh — persistent, TMA-fed point-major tile kernels (sm_100a): the large-map versions of
// k_linearize_tile / k_point_tile (glba_tiles.cuh).
//
// Round-1 ncu of the tile kernels: DRAM 38-48 % of the copy bandwidth, 59 % of the stall samples on the long scoreboard —
// a CTA issued no loads while it was in a compute phase, and with 4-6 CTAs per SM too few loads were in flight.  A tile's
// inputs are CONTIGUOUS slices (whole tracks: observations [k0, k1), points [j0, j1)), so here
//   * the grid is persistent (SM count x resident CTAs) and every CTA walks tiles blockIdx.x, +gridDim.x, ...;
//   * thread 0 fetches the per-observation slices of tile i+1 with 1-D bulk copies (cp.async.bulk.shared.global,
//     mbarrier complete_tx; SASS UBLKCP + SYNCS) into the other half of a two-stage shared-memory ring while the CTA
//     computes tile i, and the per-point slices that only phase 2 of tile i needs while phase 1 of tile i runs;
//   * the consumer threads therefore touch global memory only to STORE (and for rare out-of-capacity fall-backs):
//     the memory pipe is kept busy by the copy engine, not by warps in flight.
// Camera rows are staged per tile from a precomputed list of the tile's (up to TSLOTS) distinct cameras, whatever their
// ids: tracks over non-consecutive cameras (revisits, loop closures, street grids) are served from shared memory like
// banded ones; observations of a 33rd camera fall back to a global gather.
// The arithmetic per observation / per point is the arithmetic of glba_tiles.cuh, in the same order; the scalar
// partial sums are per CTA (thread-private accumulation over the CTA's tiles, one fixed-order reduction at the end).
#pragma once
#include "glba_tiles.cuh"

namespace glba {

constexpr int TSLOTS = 32;                  // camera rows staged per tile
#ifndef GLBA_NPCAP
#define GLBA_NPCAP 128
#endif
constexpr int NPCAP = GLBA_NPCAP;           // points staged per tile (more: global fall-back); < 255 (u8 local index)
constexpr int P_OBS = NT_T * OPT_LARGE;     // observations per tile (tile capacity of the large-map path)
#ifndef GLBA_P_NT
#define GLBA_P_NT 128
#endif
constexpr int P_NT = GLBA_P_NT;             // threads per CTA of the pipelined kernels
constexpr int P_OPT = P_OBS / P_NT;         // observations per thread in phase 1
static_assert(P_OBS % P_NT == 0 && P_NT % 32 == 0, "P_NT");
static_assert(NPCAP < 255 && NPCAP % 8 == 0, "NPCAP");
constexpr int BM_WORDS = 512;               // camera bitmap of k_tile_meta: 16 384 cameras
static_assert(BM_WORDS % NT_T == 0, "bitmap words per thread");

struct TileMeta {
  const int4* desc;        // [n_tiles] {j0, j1, k0, k1}
  const int* cams;         // [n_tiles][TSLOTS] distinct cameras of the tile (ascending), -1 = unused
  const uint8_t* slot;     // [n_obs] index of the observation's camera in its tile's list, 255 = not staged
  const int* pm_cam;       // [n_obs] camera of every observation (fall-backs only)
  const int* pm_pt;        // [n_obs] point of every observation (fall-backs only)
  int n_tiles;
  // fused product (k_pt_pipe<2>): the tile's observations grouped by camera slot
  const uint16_t* sobs;    // [n_obs] tile-local observation indices, slot by slot (each tile owns positions [k0, k1))
  const uint16_t* sstart;  // [n_tiles][SSTART] first list position of every slot, [TSLOTS] = end
  // observations whose camera found no slot in their tile (a tile with more than TSLOTS distinct cameras; 0.1 % of the street
  // grid): listed by ascending observation index; the fused product writes their contributions to ovf_c one by one
  const int* ovf_k;        // [n_ovf] ascending
  double* ovf_c;           // [n_ovf][6]
  int n_ovf;
};
constexpr int SSTART = 40;                  // uint16 per tile: TSLOTS + 1 used, padded to 80 bytes (16-byte multiple)

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, const unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, const unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, 16-byte aligned on both sides, size a multiple of 16; completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, const unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Slice [i0, i1) of an array of esz-byte elements: the copy runs from the 16-byte boundary at or below element i0 to the
// one at or above element i1 (every array carries >= 256 bytes of slack behind its last element); element i0 lands
// slice_off() ELEMENTS into dst.  slice_bytes() is what the copy adds to the barrier's transaction count.
__device__ __forceinline__ unsigned slice_bytes(const void* base, const size_t esz, const long i0, const long i1) {
  if (i1 <= i0) return 0u;
  const size_t a0 = (size_t)base + (size_t)i0 * esz, a1 = (size_t)base + (size_t)i1 * esz;
  return (unsigned)(((a1 + 15) & ~(size_t)15) - (a0 & ~(size_t)15));
}
__device__ __forceinline__ int slice_off(const void* base, const size_t esz, const long i0) {
  return (int)((((size_t)base + (size_t)i0 * esz) & 15) / esz);
}
__device__ __forceinline__ void bulk_slice(void* dst, const void* base, const size_t esz, const long i0, const long i1, uint64_t* bar) {
  if (i1 <= i0) return;
  const size_t a0 = ((size_t)base + (size_t)i0 * esz) & ~(size_t)15;
  bulk_g2s(dst, reinterpret_cast<const void*>(a0), slice_bytes(base, esz, i0, i1), bar);
}

// ---- load time: per-tile descriptor, distinct-camera list, per-observation slot ------------------------------------
// Distinct cameras by a shared-memory bitmap over [base, base + 16 384): the rank of a camera's bit IS its slot.
__global__ void __launch_bounds__(NT_T)
k_tile_meta(const int* __restrict__ tile_pt, const int* __restrict__ pt_start, const int* __restrict__ pm_cam, const int n_cam,
            int4* __restrict__ desc, int* __restrict__ cams /* pre-filled with -1 */, uint8_t* __restrict__ slot,
            uint16_t* __restrict__ sobs, uint16_t* __restrict__ sstart, int* __restrict__ n_overflow, int* __restrict__ ovf_raw, const int ovf_cap) {
  pdl_grid_sync();
  __shared__ unsigned bm[BM_WORDS];
  __shared__ int pre[BM_WORDS];
  __shared__ int wsum[NT_T / 32];
  __shared__ int s_base;
  __shared__ uint8_t s_slot[P_OBS];
  __shared__ int s_cnt[TSLOTS], s_pos[TSLOTS + 1];
  const int t = blockIdx.x, tid = threadIdx.x;
  const int j0 = tile_pt[t], j1 = tile_pt[t + 1];
  const int k0 = pt_start[j0], k1 = pt_start[j1];
  if (tid == 0) desc[t] = make_int4(j0, j1, k0, k1);
  for (int w = tid; w < BM_WORDS; w += NT_T) bm[w] = 0u;
  int base = 0;
  if (n_cam > BM_WORDS * 32) {       // more cameras than bits: window from the tile's smallest camera
    int m = 0x7fffffff;
    for (int k = k0 + tid; k < k1; k += NT_T) m = min(m, pm_cam[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) wsum[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) { for (int w = 1; w < NT_T / 32; ++w) m = min(m, wsum[w]); s_base = (m == 0x7fffffff) ? 0 : m; }
    __syncthreads();
    base = s_base;
  }
  __syncthreads();
  for (int k = k0 + tid; k < k1; k += NT_T) {
    const unsigned c = (unsigned)(pm_cam[k] - base);
    if (c < (unsigned)(BM_WORDS * 32)) atomicOr(&bm[c >> 5], 1u << (c & 31));
  }
  __syncthreads();
  // exclusive prefix of the word popcounts: BM_WORDS / NT_T consecutive words per thread, warp scan, warp totals
  constexpr int WPT = BM_WORDS / NT_T;
  int cnt[WPT], tot = 0;
#pragma unroll
  for (int q = 0; q < WPT; ++q) { cnt[q] = __popc(bm[tid * WPT + q]); tot += cnt[q]; }
  int inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if ((tid & 31) >= o) inc += y; }
  if ((tid & 31) == 31) wsum[tid >> 5] = inc;
  __syncthreads();
  int off = inc - tot;
  for (int w = 0; w < (tid >> 5); ++w) off += wsum[w];
#pragma unroll
  for (int q = 0; q < WPT; ++q) { pre[tid * WPT + q] = off; off += cnt[q]; }
  __syncthreads();
  for (int k = k0 + tid; k < k1; k += NT_T) {
    const int cam = pm_cam[k];
    const unsigned c = (unsigned)(cam - base);
    int s = 255;
    if (c < (unsigned)(BM_WORDS * 32)) {
      const int r = pre[c >> 5] + __popc(bm[c >> 5] & ((1u << (c & 31)) - 1u));
      if (r < TSLOTS) { s = r; cams[(size_t)TSLOTS * t + r] = cam; }      // same value from every observation of the camera
    }
    slot[k] = (uint8_t)s;
    if (k - k0 < P_OBS) s_slot[k - k0] = (uint8_t)s;
    if (s == 255) { const int q = atomicAdd(n_overflow, 1); if (q < ovf_cap) ovf_raw[q] = k; }      // (sorted afterwards: order-independent)
  }
  // the tile's observations grouped by slot, in observation order inside a slot (fixed order: the per-camera sums of the
  // fused product are reproducible): one thread per slot walks the tile
  __syncthreads();
  const int nobs = min(k1 - k0, P_OBS);
  if (tid < TSLOTS) {
    int c = 0;
    for (int l = 0; l < nobs; ++l) c += (s_slot[l] == tid) ? 1 : 0;
    s_cnt[tid] = c;
  }
  __syncthreads();
  if (tid == 0) {
    int acc = 0;
    for (int q = 0; q < TSLOTS; ++q) { s_pos[q] = acc; acc += s_cnt[q]; }
    s_pos[TSLOTS] = acc;
  }
  __syncthreads();
  if (tid <= TSLOTS) sstart[(size_t)SSTART * t + tid] = (uint16_t)s_pos[tid];
  if (tid < TSLOTS) {
    int pos = s_pos[tid];
    for (int l = 0; l < nobs; ++l)
      if (s_slot[l] == tid) sobs[k0 + pos++] = (uint16_t)l;
  }
}

// ---- camera rows of a tile's camera list ---------------------------------------------------------------------------
// PARTS 16-byte pieces per row, requested into registers (the loads fly during phase 1) and stored into the ring later
template <int PARTS>
struct RowFetch {
  static constexpr int NQ = (TSLOTS * PARTS + P_NT - 1) / P_NT;
  double2 v[NQ];
  bool ok[NQ];
  // cams: the tile's camera list, in shared memory (it travels with the previous tile's ring fill) or, for a CTA's
  // first tile, in global memory
  __device__ __forceinline__ void fetch(const int* cams, const double* __restrict__ tab, const int stride) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int t = threadIdx.x + q * P_NT;
      ok[q] = false;
      if (t < TSLOTS * PARTS) {
        const int row = t / PARTS, part = t - row * PARTS;
        const int cam = cams[row];
        if (cam >= 0) { v[q] = __ldg(reinterpret_cast<const double2*>(tab + (size_t)stride * cam) + part); ok[q] = true; }
      }
    }
  }
  template <int ROWSTRIDE>
  __device__ __forceinline__ void store(double* dst) const {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int t = threadIdx.x + q * P_NT;
      if (t < TSLOTS * PARTS && ok[q]) {
        const int row = t / PARTS, part = t - row * PARTS;
        *reinterpret_cast<double2*>(&dst[row * ROWSTRIDE + 2 * part]) = v[q];
      }
    }
  }
};

// ---- shared-memory stages --------------------------------------------------------------------------------------------
// "ring": per-observation slices and what phase 1 needs, two copies (tile i computes while tile i+1 lands)
// "late": per-point slices only phase 2 needs, one copy, fetched while phase 1 of the same tile runs
struct __align__(128) LinRing {
  double2 uv[P_OBS];
  double4 pt[NPCAP];
  double cs[TSLOTS * CROW];
  int pm2cm[P_OBS + 4];
  int ptst[NPCAP + 8];
  uint8_t slot[P_OBS + 16];
  int cams_next[TSLOTS];     // camera list and descriptor of the CTA's NEXT tile: they ride with this tile's fill, so
  int4 desc_next;            // neither the row fetch nor the next fill waits for an index load
  int4 desc;
};
struct __align__(128) LinSmem {
  LinRing ring[2];
  double val[8][P_OBS];
  uint8_t pidx[P_OBS];
  double red[5 * P_NT / 32 + 8];
  uint64_t full[2];
};

// next_tile: the tile this CTA processes after the one being fetched (>= n_tiles: none)
__device__ __forceinline__ void lin_issue_ring(LinRing& R, uint64_t* bar, const int4 d, const int next_tile, const PmArgs& A, const TileMeta& M,
                                               const double4* pt) {
  R.desc = d;
  const int np = min(d.y - d.x, NPCAP);
  const unsigned b_uv = (unsigned)(d.w - d.z) * 16u, b_pt = (unsigned)np * 32u;
  const unsigned b_next = next_tile < M.n_tiles ? (unsigned)(TSLOTS * 4 + 16) : 0u;
  const unsigned bytes = b_uv + b_pt + b_next + slice_bytes(A.pm2cm, 4, d.z, d.w) + slice_bytes(A.pt_start, 4, d.x, d.x + np + 1) +
                         slice_bytes(M.slot, 1, d.z, d.w);
  mbar_expect_tx(bar, bytes);
  if (b_next) {
    bulk_g2s(R.cams_next, M.cams + (size_t)TSLOTS * next_tile, TSLOTS * 4, bar);
    bulk_g2s(&R.desc_next, M.desc + next_tile, 16, bar);
  }
  if (b_uv) bulk_g2s(R.uv, A.pm_uv + d.z, b_uv, bar);
  if (b_pt) bulk_g2s(R.pt, pt + d.x, b_pt, bar);
  bulk_slice(R.pm2cm, A.pm2cm, 4, d.z, d.w, bar);
  bulk_slice(R.ptst, A.pt_start, 4, d.x, d.x + np + 1, bar);
  bulk_slice(R.slot, M.slot, 1, d.z, d.w, bar);
}

// K_A + point half of K_B, persistent + TMA-fed.  Same outputs as k_linearize_tile; the scalar partials are per CTA.
__global__ void __launch_bounds__(P_NT)
k_lin_pipe(const PmArgs A, const TileMeta M, const double4* __restrict__ pt, const double* __restrict__ camtab,
           double4* __restrict__ rec_pm, double4* __restrict__ rec_cm, double* __restrict__ Craw, double4* __restrict__ sp4,
           double4* __restrict__ lam4, double* __restrict__ cinv, double4* __restrict__ u0p, const int first, const int jacobi,
           const double min_diag, const double max_diag, const double inv_radius_arg, double* __restrict__ part /* [grid][5] */,
           const RedArgs RA) {
  pdl_grid_sync();
  extern __shared__ __align__(128) unsigned char dsm_raw[];
  LinSmem& S = *reinterpret_cast<LinSmem*>(dsm_raw);
  if (ctl_skip(RA.ctl, RA.gate)) return;
  const double inv_radius = ctl_inv_radius(RA.ctl, inv_radius_arg);
  const int tid = threadIdx.x;
  const int n_my = (M.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) { mbar_init(&S.full[0], 1); mbar_init(&S.full[1], 1); mbar_fence_init(); }
  __syncthreads();
  {   // prologue: first tile of this CTA
    RowFetch<6> F;
    F.fetch(M.cams + (size_t)TSLOTS * blockIdx.x, camtab, CAMTAB);
    if (tid == 0) lin_issue_ring(S.ring[0], &S.full[0], __ldg(M.desc + blockIdx.x), blockIdx.x + gridDim.x, A, M, pt);
    F.store<CROW>(S.ring[0].cs);
  }
  __syncthreads();
  double cost = 0.0, bad = 0.0, xn2 = 0.0, gmax = 0.0, notpd = 0.0;
  for (int i = 0; i < n_my; ++i) {
    const int s = i & 1;
    LinRing& R = S.ring[s];
    const int tile = blockIdx.x + i * gridDim.x;
    mbar_wait(&S.full[s], (unsigned)((i >> 1) & 1));
    const int4 d = R.desc;
    const int j0 = d.x, k0 = d.z;
    const int np = d.y - d.x, n = d.w - d.z;
    const int o_c = slice_off(A.pm2cm, 4, k0), o_s = slice_off(M.slot, 1, k0), o_j = slice_off(A.pt_start, 4, j0);
    auto pts = [&](const int p) -> int { return (p <= NPCAP) ? R.ptst[o_j + p] : __ldg(A.pt_start + j0 + p); };
    // local point index of every observation (replaces the pm_pt stream); 255 = beyond the staged points
    for (int p = tid; p < np; p += P_NT) {
      const int b = pts(p) - k0, e = pts(p + 1) - k0;
      const uint8_t pv = (uint8_t)(p < NPCAP ? p : 255);
      for (int m = b; m < e; ++m) S.pidx[m] = pv;
    }
    __syncthreads();            // (A) every thread is past phase 2 of the previous tile: ring[s^1], late and val are free
    RowFetch<6> F;
    const bool more = (i + 1 < n_my);
    if (more) F.fetch(R.cams_next, camtab, CAMTAB);
    if (more && tid == 0) lin_issue_ring(S.ring[s ^ 1], &S.full[s ^ 1], R.desc_next, tile + 2 * gridDim.x, A, M, pt);
    // what phase 2 needs of this thread's point besides the ring: requested now, consumed after phase 1
    double4 s4_pre = make_double4(1.0, 1.0, 1.0, 0.0);
    uint8_t free_pre = 0;
    if (tid < np) {
      free_pre = __ldg(A.pt_free + j0 + tid);
      if (!first) s4_pre = ldg4(sp4 + j0 + tid);
    }
    // ---- phase 1: one thread per observation ---------------------------------------------------------------------
#pragma unroll
    for (int m = 0; m < P_OPT; ++m) {
      const int l = m * P_NT + tid;
      if (l < n) {
        const int k = k0 + l;
        const double2 uv = R.uv[l];
        const int pl = S.pidx[l];
        const double4 X = (pl != 255) ? R.pt[pl] : ldg4(pt + __ldg(M.pm_pt + k));
        const int sl = R.slot[o_s + l];
        double Rm[9], cc[3];
        if (sl != 255) {
          const double2* cr = reinterpret_cast<const double2*>(&R.cs[sl * CROW]);
          const double2 a0 = cr[0], a1 = cr[1], a2 = cr[2], a3 = cr[3], a4 = cr[4], a5 = cr[5];
          Rm[0] = a0.x; Rm[1] = a0.y; Rm[2] = a1.x; Rm[3] = a1.y; Rm[4] = a2.x; Rm[5] = a2.y; Rm[6] = a3.x; Rm[7] = a3.y; Rm[8] = a4.x;
          cc[0] = a4.y; cc[1] = a5.x; cc[2] = a5.y;
        } else {
          load_Rc(camtab + (size_t)CAMTAB * __ldg(M.pm_cam + k), Rm, cc);
        }
        const double qx = X.x - cc[0], qy = X.y - cc[1], qz = X.z - cc[2];
        const double px = Rm[0] * qx + Rm[1] * qy + Rm[2] * qz;
        const double py = Rm[3] * qx + Rm[4] * qy + Rm[5] * qz;
        const double pz = Rm[6] * qx + Rm[7] * qy + Rm[8] * qz;
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
        st4(rec_cm + R.pm2cm[o_c + l], rec);
        double ap[3], bp[3];
        jp_rows(rec, Rm, A.K, ap, bp);
        S.val[0][l] = ap[0]; S.val[1][l] = ap[1]; S.val[2][l] = ap[2]; S.val[3][l] = bp[0]; S.val[4][l] = bp[1]; S.val[5][l] = bp[2];
        S.val[6][l] = w * rx; S.val[7][l] = w * ry;
      }
    }
    if (more) F.store<CROW>(S.ring[s ^ 1].cs);
    __syncthreads();            // (B)
    // ---- phase 2: one thread per point -----------------------------------------------------------------------------
    for (int p = tid; p < np; p += P_NT) {
      const int j = j0 + p;
      const int b = pts(p) - k0, e = pts(p + 1) - k0;
      const bool staged = p < NPCAP;
      const bool free_pt = (p == tid ? free_pre : A.pt_free[j]) != 0;
      const double4 Xp = staged ? R.pt[p] : ldg4(pt + j);
      double4 s4 = make_double4(1.0, 1.0, 1.0, 0.0);
      if (!first && free_pt) s4 = (p == tid) ? s4_pre : ldg4(sp4 + j);
      double C[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
      for (int m = b; m < e; ++m) {
        const double a0 = S.val[0][m], a1 = S.val[1][m], a2 = S.val[2][m], b0 = S.val[3][m], b1 = S.val[4][m], b2 = S.val[5][m];
        const double r0 = S.val[6][m], r1 = S.val[7][m];
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
        double sc[3], lam[3];
        if (first) {
#pragma unroll
          for (int q = 0; q < 3; ++q) sc[q] = jacobi ? 1.0 / (1.0 + sqrt(h[q])) : 1.0;
          st4(sp4 + j, make_double4(sc[0], sc[1], sc[2], 0.0));
        } else {
          sc[0] = s4.x; sc[1] = s4.y; sc[2] = s4.z;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          // lam = clamp(s^2 h, min, max) / s^2, which is h itself unless the clamp acts (then, and for NaN, divide)
          const double s2 = sc[q] * sc[q], t = s2 * h[q];
          lam[q] = (t >= min_diag && t <= max_diag) ? h[q] : fmin(fmax(t, min_diag), max_diag) / s2;
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
  }
  __syncthreads();
  double v[4] = {cost, xn2, bad, notpd};
  block_reduce<4, P_NT>(v, S.red, S.red + 4 * P_NT / 32);
  if (tid < 4) part[(size_t)5 * blockIdx.x + tid] = S.red[4 * P_NT / 32 + tid];
  __syncthreads();
  double mx[1] = {gmax};
  block_reduce<1, P_NT, true>(mx, S.red, S.red + P_NT / 32);
  if (tid == 0) part[(size_t)5 * blockIdx.x + 4] = S.red[P_NT / 32];
  __syncthreads();
  last_block_reduce5<4, P_NT>(part, gridDim.x, RA.slots, RA.counter, RA.scal, S.red, S.red + 5 * P_NT / 32);
}
static_assert(sizeof(LinSmem) + 1024 <= 233472 / 3, "k_lin_pipe: three CTAs per SM");

// ---- implicit product (MODE 0) / back-substitution + candidate cost (MODE 1) -------------------------------------------
template <int MODE>
struct __align__(128) PtRing {
  double4 rec[P_OBS];
  double xs[TSLOTS * XROW];
  double cs[MODE == 1 ? TSLOTS * CROW : 2];
  int ptst[NPCAP + 8];
  uint8_t slot[P_OBS + 16];
  uint16_t sobs[MODE == 2 ? P_OBS + 8 : 8];   // fused product: observations by slot
  uint16_t sstart[MODE == 2 ? SSTART : 8];
  int cams_next[TSLOTS];
  int4 desc_next;
  int4 desc;
};
template <int MODE>
struct __align__(128) PtSmem {
  PtRing<MODE> ring[2];
  double val[3][P_OBS];
  double4 ptc[MODE == 1 ? NPCAP : 1];       // candidate points of the tile (phase 3 re-reads them)
  double us[MODE == 2 ? 3 * NPCAP : 2];     // fused product: u_j of the tile's points
  // local point index per observation (phase 3).  MODE 1 has no barrier behind phase 3, so the next tile's indices (written
  // before its first barrier) go to the other half: one copy per ring stage
  uint8_t pidx[MODE == 1 ? 2 * P_OBS : (MODE == 2 ? P_OBS : 16)];
  double red[5 * P_NT / 32 + 8];
  uint64_t full[2];
};

template <int MODE>
__device__ __forceinline__ void pt_issue_ring(PtRing<MODE>& R, uint64_t* bar, const int4 d, const int this_tile, const int next_tile,
                                              const PmArgs& A, const TileMeta& M, const double4* rec_pm) {
  R.desc = d;
  const int np = min(d.y - d.x, NPCAP);
  const unsigned b_rec = (unsigned)(d.w - d.z) * 32u;
  const unsigned b_next = next_tile < M.n_tiles ? (unsigned)(TSLOTS * 4 + 16) : 0u;
  const unsigned b_lists = (MODE == 2) ? slice_bytes(M.sobs, 2, d.z, d.w) + (unsigned)(SSTART * 2) : 0u;
  // MODE 2 overwrites consumed records of this stage with generic stores: order them before the copy engine's writes
  if (MODE == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(bar, b_rec + b_next + b_lists + slice_bytes(A.pt_start, 4, d.x, d.x + np + 1) + slice_bytes(M.slot, 1, d.z, d.w));
  if (b_next) {
    bulk_g2s(R.cams_next, M.cams + (size_t)TSLOTS * next_tile, TSLOTS * 4, bar);
    bulk_g2s(&R.desc_next, M.desc + next_tile, 16, bar);
  }
  if (MODE == 2) {
    bulk_slice(R.sobs, M.sobs, 2, d.z, d.w, bar);
    bulk_g2s(R.sstart, M.sstart + (size_t)SSTART * this_tile, SSTART * 2, bar);
  }
  if (b_rec) bulk_g2s(R.rec, rec_pm + d.z, b_rec, bar);
  bulk_slice(R.ptst, A.pt_start, 4, d.x, d.x + np + 1, bar);
  bulk_slice(R.slot, M.slot, 1, d.z, d.w, bar);
}

static_assert(sizeof(PtSmem<0>) + 1024 <= 233472 / 4, "k_pt_pipe<0>: four CTAs per SM");
static_assert(sizeof(PtSmem<1>) + 1024 <= 233472 / 3, "k_pt_pipe<1>: three CTAs per SM");
static_assert(sizeof(PtSmem<2>) + 1024 <= 233472 / 3, "k_pt_pipe<2>: three CTAs per SM");
// xtab row of camera i: [xg(6) = T_i x_i | R_i (9) | small-angle flag] — 128 bytes, eight 16-byte pieces
template <int MODE>
__global__ void __launch_bounds__(P_NT)
k_pt_pipe(const PmArgs A, const TileMeta M, const double4* __restrict__ rec_pm, const double* __restrict__ camtab,
          const double* __restrict__ xtab, const double* __restrict__ cinv, const double4* __restrict__ u0p, double4* __restrict__ u4,
          const CgState* __restrict__ cg, const int li,
          // MODE 1 only:
          const double4* __restrict__ pt, double4* pt_c, const double* __restrict__ camtab_c, const double* __restrict__ Craw,
          const double4* __restrict__ lam4, const double inv_radius_arg, double* __restrict__ part /* [grid][5]; MODE 2: [n_tiles][TSLOTS][6] */, const RedArgs RA) {
  pdl_grid_sync();
  extern __shared__ __align__(128) unsigned char dsm_raw[];
  PtSmem<MODE>& S = *reinterpret_cast<PtSmem<MODE>*>(dsm_raw);
  if (MODE != 1 && cg && cg->done_at <= li) return;
  if (MODE == 1 && ctl_skip(RA.ctl, RA.gate)) return;
  const double inv_radius = (MODE == 1) ? ctl_inv_radius(RA.ctl, inv_radius_arg) : inv_radius_arg;
  const int tid = threadIdx.x;
  const int n_my = (M.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) { mbar_init(&S.full[0], 1); mbar_init(&S.full[1], 1); mbar_fence_init(); }
  __syncthreads();
  {
    RowFetch<8> FX;
    FX.fetch(M.cams + (size_t)TSLOTS * blockIdx.x, xtab, XTAB);
    if (tid == 0) pt_issue_ring<MODE>(S.ring[0], &S.full[0], __ldg(M.desc + blockIdx.x), blockIdx.x, blockIdx.x + gridDim.x, A, M, rec_pm);
    FX.template store<XROW>(S.ring[0].xs);
    if (MODE == 1) {
      RowFetch<6> FC;
      FC.fetch(M.cams + (size_t)TSLOTS * blockIdx.x, camtab_c, CAMTAB);
      FC.template store<CROW>(S.ring[0].cs);
    }
  }
  __syncthreads();
  double cost_c = 0.0, bad = 0.0, yn2 = 0.0, yg = 0.0, yly = 0.0;
  const double* gpl[3] = {Craw, Craw, Craw};
  if (MODE == 1) { gpl[0] = Craw + (size_t)6 * A.n_pt; gpl[1] = Craw + (size_t)7 * A.n_pt; gpl[2] = Craw + (size_t)8 * A.n_pt; }
  for (int i = 0; i < n_my; ++i) {
    const int s = i & 1;
    PtRing<MODE>& R = S.ring[s];
    const int tile = blockIdx.x + i * gridDim.x;
    mbar_wait(&S.full[s], (unsigned)((i >> 1) & 1));
    const int4 d = R.desc;
    const int j0 = d.x, k0 = d.z;
    const int np = d.y - d.x, n = d.w - d.z;
    const int o_s = slice_off(M.slot, 1, k0), o_j = slice_off(A.pt_start, 4, j0);
    auto pts = [&](const int p) -> int { return (p <= NPCAP) ? R.ptst[o_j + p] : __ldg(A.pt_start + j0 + p); };
    uint8_t* const pidx = S.pidx + (MODE == 1 ? s * P_OBS : 0);
    if (MODE != 0)
      for (int p = tid; p < np; p += P_NT) {
        const int b = pts(p) - k0, e = pts(p + 1) - k0;
        const uint8_t pv = (uint8_t)(p < NPCAP ? p : 255);
        for (int m = b; m < e; ++m) pidx[m] = pv;
      }
    __syncthreads();            // (A)
    RowFetch<8> FX;
    RowFetch<6> FC;
    const bool more = (i + 1 < n_my);
    if (more) {
      FX.fetch(R.cams_next, xtab, XTAB);
      if (MODE == 1) FC.fetch(R.cams_next, camtab_c, CAMTAB);
    }
    if (more && tid == 0) pt_issue_ring<MODE>(S.ring[s ^ 1], &S.full[s ^ 1], R.desc_next, tile + gridDim.x, tile + 2 * gridDim.x, A, M, rec_pm);
    // what phase 2 needs of this thread's point: requested now, consumed after phase 1
    double Ci_pre[6] = {0, 0, 0, 0, 0, 0};
    uint8_t free_pre = 0;
    double4 u0_pre = make_double4(0, 0, 0, 0), X_pre = u0_pre, l4_pre = u0_pre;
    double g_pre[3] = {0, 0, 0};
    double2 uv_pre[MODE == 1 ? P_OPT : 1];      // measurements of this thread's observations, for the candidate cost (phase 3)
    if (MODE == 1) {
#pragma unroll
      for (int m = 0; m < P_OPT; ++m) { const int l = m * P_NT + tid; uv_pre[m] = (l < n) ? __ldg(A.pm_uv + k0 + l) : make_double2(0.0, 0.0); }
    }
    if (tid < np) {
      const int j = j0 + tid;
      free_pre = __ldg(A.pt_free + j);
      if (MODE == 1) {
        double u3[3];
        load_pblk(cinv, u0p, j, Ci_pre, u3);
        u0_pre = make_double4(u3[0], u3[1], u3[2], 0.0);
        X_pre = ldg4(pt + j); l4_pre = ldg4(lam4 + j);
        g_pre[0] = __ldg(gpl[0] + j); g_pre[1] = __ldg(gpl[1] + j); g_pre[2] = __ldg(gpl[2] + j);
      } else {
        load_cinv(cinv, j, Ci_pre);
      }
    }
    // ---- phase 1: one thread per observation: t contribution J~p' (J~c x) -----------------------------------------
#pragma unroll
    for (int m = 0; m < P_OPT; ++m) {
      const int l = m * P_NT + tid;
      if (l < n) {
        const double4 rec = R.rec[l];
        const int sl = R.slot[o_s + l];
        double xg[6], Rm[9], flag;
        int cam_i = -1;
        if (sl != 255) {
          const double2* xr = reinterpret_cast<const double2*>(&R.xs[sl * XROW]);
          const double2 a0 = xr[0], a1 = xr[1], a2 = xr[2], a3 = xr[3], a4 = xr[4], a5 = xr[5], a6 = xr[6], a7 = xr[7];
          xg[0] = a0.x; xg[1] = a0.y; xg[2] = a1.x; xg[3] = a1.y; xg[4] = a2.x; xg[5] = a2.y;
          Rm[0] = a3.x; Rm[1] = a3.y; Rm[2] = a4.x; Rm[3] = a4.y; Rm[4] = a5.x; Rm[5] = a5.y; Rm[6] = a6.x; Rm[7] = a6.y; Rm[8] = a7.x;
          flag = a7.y;
        } else {
          cam_i = __ldg(M.pm_cam + k0 + l);
          const double4* xr = reinterpret_cast<const double4*>(xtab + (size_t)XTAB * cam_i);
          const double4 x0 = ldg4(xr), x1 = ldg4(xr + 1), x2 = ldg4(xr + 2), x3 = ldg4(xr + 3);
          xg[0] = x0.x; xg[1] = x0.y; xg[2] = x0.z; xg[3] = x0.w; xg[4] = x1.x; xg[5] = x1.y;
          Rm[0] = x1.z; Rm[1] = x1.w; Rm[2] = x2.x; Rm[3] = x2.y; Rm[4] = x2.z; Rm[5] = x2.w; Rm[6] = x3.x; Rm[7] = x3.y; Rm[8] = x3.z;
          flag = x3.w;
        }
        double sv0 = 0.0, sv1 = 0.0, sv2 = 0.0;
        if (flag != 0.0) {               // Ceres' small-angle branch: only an (almost) exactly-identity keyframe
          if (cam_i < 0) cam_i = __ldg(M.pm_cam + k0 + l);
          const double* ct = camtab + (size_t)CAMTAB * cam_i;
          sv0 = ct[CT_SV]; sv1 = ct[CT_SV + 1]; sv2 = ct[CT_SV + 2];
        }
        double a[6], bb[6];
        jhat_rows(rec, sv0, sv1, sv2, A.K, a, bb);
        double al0 = 0.0, al1 = 0.0;
#pragma unroll
        for (int r = 0; r < 6; ++r) { al0 += a[r] * xg[r]; al1 += bb[r] * xg[r]; }
        double ap[3], bp[3];
        jp_rows(rec, Rm, A.K, ap, bp);
        S.val[0][l] = ap[0] * al0 + bp[0] * al1;
        S.val[1][l] = ap[1] * al0 + bp[1] * al1;
        S.val[2][l] = ap[2] * al0 + bp[2] * al1;
      }
    }
    if (more) {
      FX.template store<XROW>(S.ring[s ^ 1].xs);
      if (MODE == 1) FC.template store<CROW>(S.ring[s ^ 1].cs);
    }
    __syncthreads();            // (B)
    // ---- phase 2: one thread per point ---------------------------------------------------------------------------
    for (int p = tid; p < np; p += P_NT) {
      const int j = j0 + p;
      const int b = pts(p) - k0, e = pts(p + 1) - k0;
      const bool staged = p < NPCAP;
      const bool mine = (p == tid);
      const bool free_pt = (mine ? free_pre : A.pt_free[j]) != 0;
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      if (free_pt)
        for (int m = b; m < e; ++m) { t0 += S.val[0][m]; t1 += S.val[1][m]; t2 += S.val[2][m]; }
      double Ci[6], u3[3] = {u0_pre.x, u0_pre.y, u0_pre.z};
      if (mine) {
#pragma unroll
        for (int q = 0; q < 6; ++q) Ci[q] = Ci_pre[q];
      } else if (MODE == 1) {
        load_pblk(cinv, u0p, j, Ci, u3);
      } else {
        load_cinv(cinv, j, Ci);
      }
      const double v0 = Ci[0] * t0 + Ci[1] * t1 + Ci[2] * t2;
      const double v1 = Ci[1] * t0 + Ci[3] * t1 + Ci[4] * t2;
      const double v2 = Ci[2] * t0 + Ci[4] * t1 + Ci[5] * t2;
      if (MODE == 0) {
        st4(u4 + j, make_double4(v0, v1, v2, 0.0));
      } else if (MODE == 2) {
        if (staged) { S.us[3 * p] = v0; S.us[3 * p + 1] = v1; S.us[3 * p + 2] = v2; }
        else st4(u4 + j, make_double4(v0, v1, v2, 0.0));           // beyond the staged points: through global memory
      } else {
        const double y0 = u3[0] - v0, y1 = u3[1] - v1, y2 = u3[2] - v2;
        const double4 X = mine ? X_pre : ldg4(pt + j);
        const double4 Xc = make_double4(X.x - y0, X.y - y1, X.z - y2, X.w);
        st4(pt_c + j, Xc);
        if (staged) S.ptc[p] = Xc;
        if (free_pt) {
          const double4 l4 = mine ? l4_pre : ldg4(lam4 + j);
          const double g0 = mine ? g_pre[0] : gpl[0][j];
          const double g1 = mine ? g_pre[1] : gpl[1][j];
          const double g2 = mine ? g_pre[2] : gpl[2][j];
          yn2 += y0 * y0 + y1 * y1 + y2 * y2;
          yg += y0 * g0 + y1 * g1 + y2 * g2;
          yly += (l4.x * y0 * y0 + l4.y * y1 * y1 + l4.z * y2 * y2) * inv_radius;
        }
      }
    }
    if (MODE == 2) {
      // ---- camera half of the product, fused: yhat_i += J^' (J~p u_j) summed per camera SLOT of the tile -----------------
      // EXPERIMENT (GLBA_FUSED=1, off by default).  It removes k_spmv_cm and its 68 B/observation, but measured on C4 the PCG
      // iteration got SLOWER (221 vs 156 us): the recomputation of the Jacobian rows, two more barriers per tile, the serial
      // per-slot sums of phase 4 and 3 instead of 4 CTAs per SM cost more than the saved 58 us kernel.  Kept because it is
      // parity-tested (same trajectories as the two-kernel product) and is the starting point for a cheaper slot reduction.
      __syncthreads();                  // (C) u_j of the tile's points are in S.us
#pragma unroll
      for (int m = 0; m < P_OPT; ++m) {
        const int l = m * P_NT + tid;
        if (l < n) {
          const double4 rec = R.rec[l];
          const int sl = R.slot[o_s + l];
          double Rm[9], sv0 = 0.0, sv1 = 0.0, sv2 = 0.0;
          if (sl != 255) {
            const double2* xr = reinterpret_cast<const double2*>(&R.xs[sl * XROW]);
            const double2 a3 = xr[3], a4 = xr[4], a5 = xr[5], a6 = xr[6], a7 = xr[7];
            Rm[0] = a3.x; Rm[1] = a3.y; Rm[2] = a4.x; Rm[3] = a4.y; Rm[4] = a5.x; Rm[5] = a5.y; Rm[6] = a6.x; Rm[7] = a6.y; Rm[8] = a7.x;
            if (a7.y != 0.0) {
              const double* ct = camtab + (size_t)CAMTAB * __ldg(M.pm_cam + k0 + l);
              sv0 = ct[CT_SV]; sv1 = ct[CT_SV + 1]; sv2 = ct[CT_SV + 2];
            }
          } else {
            const double* ct = camtab + (size_t)CAMTAB * __ldg(M.pm_cam + k0 + l);
            double cc[3];
            load_Rc(ct, Rm, cc);
            if (ct[CT_SV] != 0.0 || ct[CT_SV + 1] != 0.0 || ct[CT_SV + 2] != 0.0) { sv0 = ct[CT_SV]; sv1 = ct[CT_SV + 1]; sv2 = ct[CT_SV + 2]; }
          }
          const int pl = pidx[l];
          double u0, u1, u2;
          if (pl != 255) { u0 = S.us[3 * pl]; u1 = S.us[3 * pl + 1]; u2 = S.us[3 * pl + 2]; }
          else { const double4 uu = ld4(u4 + __ldg(M.pm_pt + k0 + l)); u0 = uu.x; u1 = uu.y; u2 = uu.z; }
          double ap[3], bp[3], a[6], bb[6];
          jp_rows(rec, Rm, A.K, ap, bp);
          const double f0 = ap[0] * u0 + ap[1] * u1 + ap[2] * u2;
          const double f1 = bp[0] * u0 + bp[1] * u1 + bp[2] * u2;
          jhat_rows(rec, sv0, sv1, sv2, A.K, a, bb);
          // the record has been consumed: its 32 bytes and two of the (free) val rows take the six contributions
          double4 c03;
          c03.x = a[0] * f0 + bb[0] * f1; c03.y = a[1] * f0 + bb[1] * f1; c03.z = a[2] * f0 + bb[2] * f1; c03.w = a[3] * f0;
          const double c4 = bb[4] * f1, c5 = a[5] * f0 + bb[5] * f1;
          if (sl != 255) {
            R.rec[l] = c03;
            S.val[0][l] = c4;
            S.val[1][l] = c5;
          } else {
            // no slot: the contribution goes to this observation's row of the overflow list (binary search: rare path)
            const int k = k0 + l;
            int lo = 0, hi = M.n_ovf;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(M.ovf_k + mid) < k) lo = mid + 1; else hi = mid; }
            double* oc = M.ovf_c + (size_t)6 * lo;
            oc[0] = c03.x; oc[1] = c03.y; oc[2] = c03.z; oc[3] = c03.w; oc[4] = c4; oc[5] = c5;
          }
        }
      }
      // the stores above went into a ring stage the copy engine refills two tiles from now: order them before the async proxy
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();                  // (D)
      const int o_l = slice_off(M.sobs, 2, k0);
      for (int e = tid; e < TSLOTS * 6; e += P_NT) {
        const int sl = e / 6, comp = e - sl * 6;
        const int q0 = R.sstart[sl], q1 = R.sstart[sl + 1];
        double acc = 0.0;
        for (int q = q0; q < q1; ++q) {
          const int l = R.sobs[o_l + q];
          acc += (comp < 4) ? reinterpret_cast<const double*>(&R.rec[l])[comp] : S.val[comp - 4][l];
        }
        part[((size_t)tile * TSLOTS + sl) * 6 + comp] = acc;
      }
    }
    if (MODE == 1) {
      __syncthreads();                  // (C) the tile's candidate points are in S.ptc (and in global for the fall-back)
      // ---- phase 3: candidate cost, one thread per observation ---------------------------------------------------
#pragma unroll
      for (int m = 0; m < P_OPT; ++m) {
        const int l = m * P_NT + tid;
        if (l < n) {
          const int k = k0 + l;
          const double2 uv = uv_pre[m];
          const int pl = pidx[l];
          const double4 xc = (pl != 255) ? S.ptc[pl] : ld4(pt_c + __ldg(M.pm_pt + k));      // written by this CTA: coherent load
          const int sl = R.slot[o_s + l];
          double Rc[9], cc[3];
          if (sl != 255) {
            const double2* cr = reinterpret_cast<const double2*>(&R.cs[sl * CROW]);
            const double2 a0 = cr[0], a1 = cr[1], a2 = cr[2], a3 = cr[3], a4 = cr[4], a5 = cr[5];
            Rc[0] = a0.x; Rc[1] = a0.y; Rc[2] = a1.x; Rc[3] = a1.y; Rc[4] = a2.x; Rc[5] = a2.y; Rc[6] = a3.x; Rc[7] = a3.y; Rc[8] = a4.x;
            cc[0] = a4.y; cc[1] = a5.x; cc[2] = a5.y;
          } else {
            load_Rc(camtab_c + (size_t)CAMTAB * __ldg(M.pm_cam + k), Rc, cc);
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
    }
  }
  if (MODE == 1) {
    __syncthreads();
    double v[5] = {cost_c, yn2, yg, yly, bad};
    block_reduce<5, P_NT>(v, S.red, S.red + 5 * P_NT / 32);
    if (tid < 5) part[(size_t)5 * blockIdx.x + tid] = S.red[5 * P_NT / 32 + tid];
    __syncthreads();
    last_block_reduce5<-1, P_NT>(part, gridDim.x, RA.slots, RA.counter, RA.scal, S.red, S.red + 5 * P_NT / 32, &RA.hook);
  }
}

}  // namespace glba
