// glba_sparse.cuh — the reduced camera matrix as an explicit block-sparse matrix.
//
// The implicit product of glba_pipe.cuh / k_spmv_cm streams every observation twice per PCG iteration (0.13 ms on a 5 M
// observation map).  A SLAM map has banded covisibility: camera a shares points with a few dozen cameras only, so the
// off-diagonal blocks of S are few (C4: ~30 k blocks of 6x6, 8 MB) and one assembly per LM iteration (every observation
// pair of every track, once) replaces the two streaming passes of EVERY PCG iteration by a product that reads 8 MB from L2.
// This is what the reference's solvers do on the CPU (Ceres SPARSE_SCHUR / g2o's BlockSolver build the reduced camera
// matrix; /root/reference/src/core/slam_core.cpp:1079-1082 selects SPARSE_SCHUR).
//
// The sums run in the "hat" space of the camera kernels (glba_cam.cuh), S = B + lam/radius - T' S^ T with
//     S^_ab = sum over points j seen by a and b of  J^_a' (J~p_a Cinv_j J~p_b') J^_b      (6x6, a != b),
// and the assembly kernel stores S_ab = -T_a' S^_ab T_b, so the PCG works on plain camera vectors: w_a = Md_a u_a + sum_b S_ab u_b
// with Md_a = S_aa from k_cam_schur_fin.
//
// Structure (built once per loaded problem, on the device, lazily at the first PCG solve):
//   instance = (observation of a, observation of b, point) for every unordered pair of observations of one point whose
//              cameras are both free; sorted (stable radix sort) by the pair key a * n_cam + b with a < b, so the instances
//              of a block are contiguous and in point order;
//   block p  = one distinct key; pair_start[p] .. pair_start[p+1] its instances;
//   row list = for camera a the entries (b, p, transposed?) of row a of S^, both triangles, sorted by b.
// Every sum runs in a fixed order (lane-strided partial sums + xor butterfly): results do not depend on scheduling.
#pragma once
#include "glba_kernels.cuh"
#include "glba_cam.cuh"

namespace glba {

// number of instances of point j: f (f - 1) / 2 with f = its observations in free cameras; cnt[n_pt] = 0
__global__ void k_pair_count(const int n_pt, const int* __restrict__ pt_start, const int* __restrict__ pm_cam, const uint8_t* __restrict__ cam_free,
                             const uint8_t* __restrict__ pt_free, long long* __restrict__ cnt) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n_pt) return;
  long long f = 0;
  if (j < n_pt && pt_free[j])
    for (int o = pt_start[j]; o < pt_start[j + 1]; ++o) f += cam_free[pm_cam[o]] ? 1 : 0;
  cnt[j] = f * (f - 1) / 2;
}

__global__ void k_pair_emit(const int n_pt, const int* __restrict__ pt_start, const int* __restrict__ pm_cam, const uint8_t* __restrict__ cam_free,
                            const uint8_t* __restrict__ pt_free, const long long* __restrict__ off, const int n_cam /* of the whole map */,
                            const int* __restrict__ l2g /* sharded: local -> global camera (monotonic); else null */,
                            unsigned long long* __restrict__ key, int4* __restrict__ val) {
  pdl_grid_sync();
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pt || !pt_free[j]) return;
  long long pos = off[j];
  const int e = pt_start[j + 1];
  for (int oa = pt_start[j]; oa < e; ++oa) {
    const int a = pm_cam[oa];
    if (!cam_free[a]) continue;
    for (int ob = oa + 1; ob < e; ++ob) {
      const int b = pm_cam[ob];
      if (!cam_free[b]) continue;
      const bool sw = a > b;
      const int lo = sw ? b : a, hi = sw ? a : b;
      key[pos] = (unsigned long long)(l2g ? l2g[lo] : lo) * (unsigned long long)n_cam + (unsigned long long)(l2g ? l2g[hi] : hi);
      val[pos] = make_int4(sw ? ob : oa, sw ? oa : ob, j, 0);
      ++pos;
    }
  }
}

// distinct keys of this rank -> (local) cameras of the block
__global__ void k_pair_cams(const int n_pairs, const unsigned long long* __restrict__ ukey, const int n_cam, const int* __restrict__ g2l /* or null */,
                            int* __restrict__ pair_a, int* __restrict__ pair_b) {
  pdl_grid_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const unsigned long long k = ukey[p];
  const int a = (int)(k / (unsigned long long)n_cam), b = (int)(k % (unsigned long long)n_cam);
  pair_a[p] = g2l ? g2l[a] : a; pair_b[p] = g2l ? g2l[b] : b;
}
// distinct keys of the whole map -> the two row entries of every block (upper: as stored, lower: transposed)
__global__ void k_pair_rows(const int n_pairs, const unsigned long long* __restrict__ ukey, const int n_cam,
                            unsigned long long* __restrict__ ent_key, int* __restrict__ ent_val) {
  pdl_grid_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const unsigned long long k = ukey[p];
  const int a = (int)(k / (unsigned long long)n_cam), b = (int)(k % (unsigned long long)n_cam);
  ent_key[2 * (size_t)p] = k;                                                                   ent_val[2 * (size_t)p] = p;
  ent_key[2 * (size_t)p + 1] = (unsigned long long)b * (unsigned long long)n_cam + (unsigned long long)a; ent_val[2 * (size_t)p + 1] = p | (int)0x80000000;
}
// sharded maps: the blocks of the whole map are the union of the ranks' blocks.  presence[key] (summed over ranks) -> 0/1,
// its exclusive scan numbers the blocks in key order, identically on every rank.
__global__ void k_pair_mark(const int n_pairs, const unsigned long long* __restrict__ ukey, int* __restrict__ pres) {
  pdl_grid_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pairs) pres[ukey[p]] = 1;
}
__global__ void k_pair_flag01(const long n, int* __restrict__ pres) {     // pres[n] = 0 closes the scan
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k <= n) pres[k] = (k < n && pres[k] > 0) ? 1 : 0;
}
__global__ void k_pair_compact(const long n, const int* __restrict__ pres, const int* __restrict__ scan, unsigned long long* __restrict__ gkey) {
  pdl_grid_sync();
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n && pres[k]) gkey[scan[k]] = (unsigned long long)k;
}
__global__ void k_pair_gid(const int n_pairs, const unsigned long long* __restrict__ ukey, const int* __restrict__ scan, int* __restrict__ gid) {
  pdl_grid_sync();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pairs) gid[p] = scan[ukey[p]];
}
// rows of `width` doubles: out[a] = in[l2g[a]]
__global__ void k_gather_rows(const int n_act, const int* __restrict__ l2g, const double* __restrict__ in, const int width, double* __restrict__ out) {
  pdl_grid_sync();
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)n_act * width) return;
  const int a = (int)(t / width), q = (int)(t - (long)a * width);
  out[t] = in[(size_t)l2g[a] * width + q];
}
__global__ void k_pair_entries(const int n_ent, const unsigned long long* __restrict__ skey, const int* __restrict__ sval, const int n_cam,
                               int* __restrict__ ent_row, int2* __restrict__ ent) {
  pdl_grid_sync();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_ent) return;
  const unsigned long long k = skey[e];
  ent_row[e] = (int)(k / (unsigned long long)n_cam);
  ent[e] = make_int2((int)(k % (unsigned long long)n_cam), sval[e]);
}

// One warp per block: S^_ab = sum over its instances of J^_a' E J^_b, E = J~p_a Cinv J~p_b' (2x2).
// The instance list is a two-level gather (index entry -> two records + the point's inverse block) and the arithmetic per
// instance is ~190 FP64 operations at 3 CTAs/SM: the first version waited on the scoreboard 65 % of the time (ncu), so the
// loop is software-pipelined: index entries two trips ahead, records and inverse block one trip ahead.
constexpr int NT_SP = 128;
__device__ __forceinline__ void pair_accumulate(const double4 ra, const double4 rb, const double* Ci, const double* Ra, const double* Rb,
                                                const double* sva, const double* svb, const Intr K, double* acc) {
  double apa[3], bpa[3], apb[3], bpb[3];
  jp_rows(ra, Ra, K, apa, bpa);
  jp_rows(rb, Rb, K, apb, bpb);
  const double t0 = Ci[0] * apb[0] + Ci[1] * apb[1] + Ci[2] * apb[2];
  const double t1 = Ci[1] * apb[0] + Ci[3] * apb[1] + Ci[4] * apb[2];
  const double t2 = Ci[2] * apb[0] + Ci[4] * apb[1] + Ci[5] * apb[2];
  const double s0 = Ci[0] * bpb[0] + Ci[1] * bpb[1] + Ci[2] * bpb[2];
  const double s1 = Ci[1] * bpb[0] + Ci[3] * bpb[1] + Ci[4] * bpb[2];
  const double s2 = Ci[2] * bpb[0] + Ci[4] * bpb[1] + Ci[5] * bpb[2];
  const double E00 = apa[0] * t0 + apa[1] * t1 + apa[2] * t2;
  const double E01 = apa[0] * s0 + apa[1] * s1 + apa[2] * s2;
  const double E10 = bpa[0] * t0 + bpa[1] * t1 + bpa[2] * t2;
  const double E11 = bpa[0] * s0 + bpa[1] * s1 + bpa[2] * s2;
  double aa[6], ba[6], ab[6], bb[6];
  jhat_rows(ra, sva[0], sva[1], sva[2], K, aa, ba);
  jhat_rows(rb, svb[0], svb[1], svb[2], K, ab, bb);
  // E J^_b: rows ea = E00 a_b + E01 b_b, eb = E10 a_b + E11 b_b   (a_b[4] = b_b[3] = 0 structurally)
  double ea[6], eb[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) { ea[c] = E00 * ab[c] + E01 * bb[c]; eb[c] = E10 * ab[c] + E11 * bb[c]; }
  ea[3] = E00 * ab[3]; eb[3] = E10 * ab[3];
  ea[4] = E01 * bb[4]; eb[4] = E11 * bb[4];
  ea[5] = E00 * ab[5] + E01 * bb[5]; eb[5] = E10 * ab[5] + E11 * bb[5];
  // J^_a' (E J^_b): row r = a_a[r] ea + b_a[r] eb   (a_a[4] = b_a[3] = 0)
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      if (r == 3) acc[r * 6 + c] += aa[3] * ea[c];
      else if (r == 4) acc[r * 6 + c] += ba[4] * eb[c];
      else acc[r * 6 + c] += aa[r] * ea[c] + ba[r] * eb[c];       // (chained FMAs, as in acc_sym_sparse, are 16 % SLOWER here: 0.494 vs 0.424 ms)
    }
}

__global__ void __launch_bounds__(NT_SP, 4)
k_schur_pairs(const int n_pairs, const int* __restrict__ pair_a, const int* __restrict__ pair_b, const int* __restrict__ pair_start,
              const int4* __restrict__ inst, const double4* __restrict__ rec_pm, const double* __restrict__ camtab, const double* __restrict__ cinv,
              const Intr K, const int* __restrict__ gid /* sharded: block number in the whole map; else null */,
              double* __restrict__ blocks /* [blocks of the whole map][36] */) {
  pdl_grid_sync();
  __shared__ double scam[NT_SP / 32][2][12];      // R[9], sv[3] of the warp's two cameras
  const int wid = threadIdx.x >> 5;
  const int p = (blockIdx.x * NT_SP + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (p >= n_pairs) return;
  const double* cta = camtab + (size_t)CAMTAB * pair_a[p];
  const double* ctb = camtab + (size_t)CAMTAB * pair_b[p];
  if (lane < 24) {
    const int q = lane % 12;
    scam[wid][lane / 12][q] = (lane < 12 ? cta : ctb)[q < 9 ? q : CT_SV + q - 9];
  }
  __syncwarp();
  const double* Ra = scam[wid][0]; const double* sva = Ra + 9;
  const double* Rb = scam[wid][1]; const double* svb = Rb + 9;
  double acc[36];
#pragma unroll
  for (int q = 0; q < 36; ++q) acc[q] = 0.0;
  const int i1 = pair_start[p + 1];
  int i = pair_start[p] + lane;
  if (i < i1) {
    int4 in1 = __ldg(inst + min(i + 32, i1 - 1));
    const int4 in0 = __ldg(inst + i);
    double4 ra = ldg4(rec_pm + in0.x), rb = ldg4(rec_pm + in0.y);
    double Ci[6];
    load_cinv(cinv, in0.z, Ci);
    for (; i < i1; i += 32) {
      const int4 in2 = __ldg(inst + min(i + 64, i1 - 1));              // clamped: a harmless re-read past the lane's last trip
      const double4 ra_n = ldg4(rec_pm + in1.x), rb_n = ldg4(rec_pm + in1.y);
      double Ci_n[6];
      load_cinv(cinv, in1.z, Ci_n);
      pair_accumulate(ra, rb, Ci, Ra, Rb, sva, svb, K, acc);
      ra = ra_n; rb = rb_n; in1 = in2;
#pragma unroll
      for (int q = 0; q < 6; ++q) Ci[q] = Ci_n[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 36; ++q)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
  // every lane holds S^_ab; S_ab = -T_a' S^_ab T_b with T = blkdiag(G, R)  (all lanes alike: static register indexing)
  double Ga[9], Gb[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) { Ga[q] = cta[CT_G + q]; Gb[q] = ctb[CT_G + q]; }
  double out[36];
#pragma unroll
  for (int I = 0; I < 2; ++I)
#pragma unroll
    for (int J = 0; J < 2; ++J) {
      const double* L = I == 0 ? Ga : Ra;
      const double* Rm = J == 0 ? Gb : Rb;
      double t[9];       // t = S^_IJ Rm
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          t[r * 3 + c] = acc[(3 * I + r) * 6 + 3 * J] * Rm[c] + acc[(3 * I + r) * 6 + 3 * J + 1] * Rm[3 + c] + acc[(3 * I + r) * 6 + 3 * J + 2] * Rm[6 + c];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          out[(3 * I + r) * 6 + 3 * J + c] = -(L[r] * t[c] + L[3 + r] * t[3 + c] + L[6 + r] * t[6 + c]);
    }
  if (lane < 9) {       // lanes 0..8 write four values each (256-bit stores)
    double4 v = make_double4(0.0, 0.0, 0.0, 0.0);
#pragma unroll
    for (int l = 0; l < 9; ++l)
      if (lane == l) v = make_double4(out[4 * l], out[4 * l + 1], out[4 * l + 2], out[4 * l + 3]);
    st4(reinterpret_cast<double4*>(blocks + (size_t)36 * (gid ? gid[p] : p)) + lane, v);
  }
}

// ---------------------------------------------------------------------------------------------
// The whole block-Jacobi PCG on the assembled matrix in ONE cooperative launch (one CTA per SM at most, all resident):
// same single-reduction recurrence and stop rules as k_cg_w / k_cg_update (glba_cam.cuh), two grid barriers per iteration.
//   phase 1 (one warp per camera row):  w_a = Md_a u_a + sum_b S_ab u_b,  partial gamma' = r.u and delta = u.w per CTA
//   barrier; warp 0 of every CTA adds the CTA partials in CTA order (lane-strided, all loads in flight + butterfly):
//            the same totals and therefore the same decisions in every CTA
//   phase 2 (lanes 0..5 of the row's warp own the six components):  p = u + beta p, s = w + beta s, x += alpha p,
//            r -= alpha s, u = Minv r
//   barrier (u is read by other rows' products)
// A row always belongs to the same warp.
// REG (at most one row per warp: up to 16 rows per CTA on up to 148 CTAs): the row's vectors and its rows of Md / Minv live in
//   registers for the whole solve, only u (every iteration) and x (at the end) go to memory; the blocks of the CTA's rows sit
//   in shared memory in row-entry order, already transposed where the entry is a lower-triangle one; the host deals the rows
//   out in contiguous ranges of (nearly) equal entry count (glba.cu, ensure_explicit).
// otherwise: several rows per warp, vectors re-read from global memory by the lanes that wrote them, blocks through L2.
// What other CTAs write (u, partial sums) is read with ld.cg.  A gpu-scope fence invalidates the SM's L1 (CCTL.IVALL), so
// nothing read through L1 survives an iteration: that is why the REG form keeps the blocks in shared memory.
// The barrier is a monotonic counter (zeroed by the host before the launch), polled with relaxed loads by one thread per CTA,
// one fence before the arrival and one after the last poll (1.4 us on 113-148 CTAs: tools/micro/grid_barrier.cu); a waiter that
// spins implausibly long flags reason 4 and leaves, so a scheduling accident cannot hang the device.
// ---------------------------------------------------------------------------------------------
constexpr int NT_CGP = 512;
constexpr unsigned CG_SPIN_LIMIT = 1u << 22;
constexpr size_t CG_SMEM_CAP = 780;                                        // row entries per CTA kept in shared memory (292 B each)
constexpr size_t CG_SMEM_BYTES = CG_SMEM_CAP * (36 * sizeof(double) + sizeof(int));
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// polls are relaxed loads (an acquire load invalidates the SM's L1 every time: ncu showed a CCTL.IVALL per poll); ONE fence
// after the last poll orders everything behind it
__device__ __forceinline__ bool grid_barrier(unsigned* bar, unsigned& target, const unsigned n_cta) {
  __shared__ int ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    target += n_cta;
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    unsigned spins = 0;
    bool fine = true;
    while (ld_relaxed_u32(bar) < target) {
      if (++spins > CG_SPIN_LIMIT) { fine = false; break; }
    }
    __threadfence();
    ok = fine ? 1 : 0;
  }
  __syncthreads();
  return ok != 0;
}
// lanes 0..5 of a warp own the six components of the row's vectors (the other lanes mirror lane 0 and store nothing)
__device__ __forceinline__ double pick6(const double* a, const int q) {
  double v = a[0];
#pragma unroll
  for (int c = 1; c < 6; ++c) v = (q == c) ? a[c] : v;
  return v;
}
__device__ __forceinline__ double sum6_lanes(const double v) {        // v of lanes 0..5 added in lane order, result in every lane
  double t = __shfl_sync(0xffffffffu, v, 0);
#pragma unroll
  for (int c = 1; c < 6; ++c) t += __shfl_sync(0xffffffffu, v, c);
  return t;
}
__device__ __forceinline__ double row6_dot_reg(const double* m /* 6 values of row q */, const double v) {   // sum_c M[q][c] v_c, v_c from lane c
  double t = 0.0;
#pragma unroll
  for (int c = 0; c < 6; ++c) t += m[c] * __shfl_sync(0xffffffffu, v, c);
  return t;
}
__device__ __forceinline__ double row6_dot(const double* __restrict__ Mrow, const double v) {
  double m[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) m[c] = __ldg(Mrow + c);
  return row6_dot_reg(m, v);
}
// sum_b S_ab u_b over the entries of one row, all lanes; totals in every lane
__device__ __forceinline__ void bsr_row(const int e0, const int e1, const int lane, const int2* __restrict__ ent, const double* __restrict__ blocks,
                                        const double* u, double* acc) {
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = 0.0;
  for (int e = e0 + lane; e < e1; e += 32) {
    const int2 en = __ldg(ent + e);
    const double2* ub = reinterpret_cast<const double2*>(u + (size_t)6 * en.x);
    const double2 x01 = __ldcg(ub), x23 = __ldcg(ub + 1), x45 = __ldcg(ub + 2);
    const double xv[6] = {x01.x, x01.y, x23.x, x23.y, x45.x, x45.y};
    const double4* bp = reinterpret_cast<const double4*>(blocks + (size_t)36 * (en.y & 0x7fffffff));
    double B[36];
#pragma unroll
    for (int k = 0; k < 9; ++k) { const double4 v = ldg4(bp + k); B[4 * k] = v.x; B[4 * k + 1] = v.y; B[4 * k + 2] = v.z; B[4 * k + 3] = v.w; }
    if (en.y >= 0) {
#pragma unroll
      for (int rr = 0; rr < 6; ++rr)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[rr] += B[rr * 6 + c] * xv[c];
    } else {
#pragma unroll
      for (int rr = 0; rr < 6; ++rr)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] += B[rr * 6 + c] * xv[rr];
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
}

// the same sum with the row's blocks in shared memory (entry le = e - E0 of the CTA, component-major); the gather of u for
// the next trip is in flight while the current trip multiplies
__device__ __forceinline__ void bsr_row_smem(const int e0, const int e1, const int E0, const int cap, const int lane, const double* sblk,
                                             const int* scol, const double* u, double* acc) {
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = 0.0;
  int e = e0 + lane;
  bool v = e < e1;
  double2 n01 = make_double2(0.0, 0.0), n23 = n01, n45 = n01;
  if (v) {
    const double2* ub = reinterpret_cast<const double2*>(u + (size_t)6 * scol[e - E0]);
    n01 = __ldcg(ub); n23 = __ldcg(ub + 1); n45 = __ldcg(ub + 2);
  }
  while (v) {
    const double xv[6] = {n01.x, n01.y, n23.x, n23.y, n45.x, n45.y};
    const int le = e - E0;
    e += 32; v = e < e1;
    if (v) {
      const double2* ub = reinterpret_cast<const double2*>(u + (size_t)6 * scol[e - E0]);
      n01 = __ldcg(ub); n23 = __ldcg(ub + 1); n45 = __ldcg(ub + 2);
    }
#pragma unroll
    for (int rr = 0; rr < 6; ++rr)
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[rr] += sblk[(size_t)(rr * 6 + c) * cap + le] * xv[c];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
}

template <bool REG>
__global__ void __launch_bounds__(NT_CGP, 1)
k_cg_bsr(const int n_cam, const int* __restrict__ row_start, const int2* __restrict__ ent,
         const double* __restrict__ blocks, const double* __restrict__ Md, const double* __restrict__ Minv, const double* __restrict__ rhs,
         double* x, double* r, double* u, double* p, double* sv, double* w, double* part /* [gridDim.x][2] */,
         unsigned* sync /* barrier counter, zeroed before the launch */,
         CgState* cg, const double tol, const int max_iters, const int cap /* REG: row entries the CTA keeps in shared memory */,
         const int* __restrict__ cta_rows /* REG: rows [cta_rows[c], cta_rows[c+1]) of CTA c, at most one per warp, balanced by entry count */,
         long long* prof /* diagnostic (GLBA_CG_PROF=1): SM cycles of CTA 0 per phase, summed over the iterations; else null */) {
  pdl_grid_sync();
  // REG: the blocks of the CTA's rows, in row-entry order and already transposed where the entry is a lower-triangle one, live in
  // shared memory for the whole solve (component-major: lane-consecutive entries hit consecutive banks), with their column
  // indices.  Every grid synchronisation invalidates L1 (fence), so without this each iteration re-fetched ~90 KB per SM from
  // L2 behind a two-level dependent load; entries beyond `cap` (none on C4: ~600 per CTA of ~790) are read from global memory.
  extern __shared__ double sblk[];            // [36][cap] doubles, then cap ints
  // rows of fixed cameras have no entries, Md = Minv = 0 and rhs = 0: they stay zero without a branch
  __shared__ double sm_dot[2][NT_CGP / 32];
  __shared__ double sm_tot[2];
  __shared__ int sm_ok;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int q = lane < 6 ? lane : 0;
  const bool owner = lane < 6;
  const int gw = blockIdx.x * (NT_CGP / 32) + wid, nw = gridDim.x * (NT_CGP / 32);
  unsigned* bar = sync;
  unsigned target = 0;
  const int first = REG ? cta_rows[blockIdx.x] : 0, last = REG ? cta_rows[blockIdx.x + 1] : 0;
  const bool has_row = REG ? (first + wid < last) : false;          // REG: this warp's only row
  const int my = has_row ? first + wid : 0;
  int re0 = 0, re1 = 0;
  double xq = 0.0, rq = 0.0, uq = 0.0, pq = 0.0, sq = 0.0, wq = 0.0;
  // x = 0, r = rhs, u = Minv r, p = s = 0
  double mdr[6] = {0, 0, 0, 0, 0, 0}, mir[6] = {0, 0, 0, 0, 0, 0};      // REG: rows q of the row's Md and Minv
  int E0 = 0;
  int* scol = reinterpret_cast<int*>(sblk + (size_t)36 * cap);
  if (REG) {
    E0 = row_start[first];
    const int nloc = row_start[last] - E0;             // <= cap by construction of cta_rows
    for (int i = threadIdx.x; i < nloc * 36; i += NT_CGP) {
      const int le = i / 36, k = i - le * 36;
      const int2 en = __ldg(ent + E0 + le);
      const int kk = (en.y >= 0) ? k : (k % 6) * 6 + k / 6;          // lower-triangle entry: store the transpose
      sblk[(size_t)kk * cap + le] = __ldg(blocks + (size_t)36 * (en.y & 0x7fffffff) + k);
      if (k == 0) scol[le] = en.x;
    }
    __syncthreads();
    if (has_row) {
      re0 = row_start[my]; re1 = row_start[my + 1];
      rq = rhs[6 * my + q];
#pragma unroll
      for (int c = 0; c < 6; ++c) { mdr[c] = __ldg(Md + (size_t)36 * my + 6 * q + c); mir[c] = __ldg(Minv + (size_t)36 * my + 6 * q + c); }
      uq = row6_dot_reg(mir, rq);
      if (owner) u[6 * my + q] = uq;
    }
  } else {
    for (int row = gw; row < n_cam; row += nw) {
      const double rr = rhs[6 * row + q];
      const double z = row6_dot(Minv + (size_t)36 * row + 6 * q, rr);
      if (owner) { x[6 * row + q] = 0.0; r[6 * row + q] = rr; p[6 * row + q] = 0.0; sv[6 * row + q] = 0.0; u[6 * row + q] = z; }
    }
  }
  int iters = 0, reason = 0;
  bool alive = grid_barrier(bar, target, gridDim.x);
  double g_old = 0.0, a_old = 0.0, g0 = 0.0;
  long long tprof[6] = {0, 0, 0, 0, 0, 0}, tc = clock64();
#define CG_PROF(slot) do { if (prof != nullptr) { const long long t__ = clock64(); tprof[slot] += t__ - tc; tc = t__; } } while (0)
  for (int li = 0; alive; ++li) {
    double ru = 0.0, uw = 0.0;
    CG_PROF(5);
    if (REG) {
      if (has_row) {
        double acc[6];
        bsr_row_smem(re0, re1, E0, cap, lane, sblk, scol, u, acc);
        wq = pick6(acc, q) + row6_dot_reg(mdr, uq);
        ru = sum6_lanes(rq * uq);
        uw = sum6_lanes(wq * uq);
      }
    } else {
      for (int row = gw; row < n_cam; row += nw) {
        const double uo = u[6 * row + q], ro = r[6 * row + q];       // written by this lane (phase 2 / start)
        double acc[6];
        bsr_row(row_start[row], row_start[row + 1], lane, ent, blocks, u, acc);
        const double wo = pick6(acc, q) + row6_dot(Md + (size_t)36 * row + 6 * q, uo);
        if (owner) w[6 * row + q] = wo;
        ru += sum6_lanes(ro * uo);
        uw += sum6_lanes(wo * uo);
      }
    }
    CG_PROF(0);                 // product + partial dot products of this warp
    if (lane == 0) { sm_dot[0][wid] = ru; sm_dot[1][wid] = uw; }
    __syncthreads();
    CG_PROF(1);                 // waiting for the CTA's slowest warp
    if (threadIdx.x == 0) {
      double a = 0.0, b = 0.0;
#pragma unroll
      for (int k = 0; k < NT_CGP / 32; ++k) { a += sm_dot[0][k]; b += sm_dot[1][k]; }
      part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b;
    }
    const bool bar_ok = grid_barrier(bar, target, gridDim.x);
    CG_PROF(2);                 // grid barrier (incl. waiting for the slowest CTA)
    if (wid == 0) {       // add the CTAs' partials in CTA order: lane-strided sums + butterfly, identical in every CTA
      const bool fine = bar_ok;
      double a = 0.0, b = 0.0;
      double2 pv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {       // all of a lane's partials in flight together (cta = lane + 32 j < 256)
        const unsigned k = lane + 32u * j;
        pv[j] = (k < gridDim.x) ? __ldcg(reinterpret_cast<const double2*>(part) + k) : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { a += pv[j].x; b += pv[j].y; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
      if (lane == 0) { sm_tot[0] = a; sm_tot[1] = b; sm_ok = fine ? 1 : 0; }
    }
    __syncthreads();
    CG_PROF(3);                 // totals
    if (!sm_ok) { reason = 4; break; }
    const double gp = sm_tot[0], dl = sm_tot[1];
    const bool first = (li == 0);
    if (first) g0 = gp;
    const bool zero_rhs = first && !(gp > 0.0);
    const bool converged = !first && sqrt(gp) <= tol * sqrt(g0);
    const double beta = first ? 0.0 : ((g_old > 0.0) ? gp / g_old : 0.0);
    const double denom = first ? dl : dl - beta * gp / a_old;
    const bool breakdown = !(denom > 0.0);
    if (zero_rhs || converged || breakdown) { iters = li; reason = (breakdown && !converged && !zero_rhs) ? 2 : 1; break; }
    const double alpha = gp / denom;
    g_old = gp; a_old = alpha;
    if (REG) {
      if (has_row) {
        pq = uq + beta * pq;
        sq = wq + beta * sq;
        xq += alpha * pq;
        rq -= alpha * sq;
        uq = row6_dot_reg(mir, rq);
        if (owner) u[6 * my + q] = uq;
      }
    } else {
      for (int row = gw; row < n_cam; row += nw) {
        const double pn = u[6 * row + q] + beta * p[6 * row + q];
        const double sn = w[6 * row + q] + beta * sv[6 * row + q];
        const double xn = x[6 * row + q] + alpha * pn;
        const double rr = r[6 * row + q] - alpha * sn;
        const double z = row6_dot(Minv + (size_t)36 * row + 6 * q, rr);
        if (owner) { p[6 * row + q] = pn; sv[6 * row + q] = sn; x[6 * row + q] = xn; r[6 * row + q] = rr; u[6 * row + q] = z; }
      }
    }
    iters = li + 1;
    CG_PROF(4);                 // vector updates
    if (li + 1 >= max_iters) { reason = 3; break; }
    if (!grid_barrier(bar, target, gridDim.x)) { reason = 4; break; }
  }
  if (prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) prof[k] = tprof[k];
    prof[6] = iters;
  }
#undef CG_PROF
  if (!alive) reason = 4;
  if (REG && has_row && owner) x[6 * my + q] = xq;
  if (blockIdx.x == 0 && threadIdx.x == 0) { cg->iters = iters; cg->reason = reason; cg->done_at = 0; cg->gamma0 = g0; cg->tol = tol; cg->max_iters = max_iters; }
}

}  // namespace glba
