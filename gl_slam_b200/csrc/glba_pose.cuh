// glba_pose.cuh — pose-only BA (GL-SLAM slam_core::pose_only_ba, src/core/slam_core.cpp:1092-1140)
// as ONE kernel launch: one CTA per frame runs the complete Ceres-semantics Levenberg-Marquardt
// loop (evaluate -> 6x6 normal equations -> Cholesky -> candidate cost -> accept/reject) on device.
// The reference spins up a Ceres problem, a thread pool and (with ceres::CUDA, :1120) a cuSOLVER
// round trip for a 6x6 system; here nothing leaves the SM until the pose is final.
#pragma once
#include "glba_kernels.cuh"

namespace glba {

constexpr int NT_POSE = 256;

struct PoseOpts {
  LossP loss;
  int max_iters;
  double function_tol, gradient_tol, parameter_tol;
  double initial_radius, max_radius, min_radius, min_relative_decrease, min_diag, max_diag;
  int jacobi, max_invalid;
};

struct PoseTrace {        // optional per-iteration trace of problem 0..batch-1, stride = GLBA_MAX_ITERS+1
  double* cost; double* cost_candidate; double* radius; double* step_norm; double* relative_decrease;
  double* gradient_max_norm; uint8_t* accepted; int stride;
};

// evaluate at camera table `ct`: cost, and (if WITH_J) H = sum J~c'J~c (21 upper) and g = J~c' r~ (6),
// accumulated in the pre-transform basis (J^) and rotated by T once at the end by thread 0.
template <bool WITH_J>
__device__ __forceinline__ void pose_eval(const double* ct /* smem CAMTAB */, const int n, const double* __restrict__ X,
                                          const double* __restrict__ uv, const Intr K, const LossP loss, double* sm,
                                          double* out /* smem: [0]=cost [1..21]=A [22..27]=ghat [28]=bad */) {
  constexpr int NV = WITH_J ? 29 : 2;
  double acc[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) acc[q] = 0.0;
  for (int k = threadIdx.x; k < n; k += NT_POSE) {
    double xh, yh, iz, rx, ry;
    project_obs(ct, X[3 * k], X[3 * k + 1], X[3 * k + 2], K, uv[2 * k], uv[2 * k + 1], xh, yh, iz, rx, ry);
    double rho, w;
    loss_eval(loss, rx * rx + ry * ry, rho, w);
    acc[0] += 0.5 * rho;
    const double isbad = (!isfinite(rx) || !isfinite(ry)) ? 1.0 : 0.0;
    if (WITH_J) {
      acc[28] += isbad;
      double a[6], b[6];
      jhat_rows(make_double4(xh, yh, iz, w), ct[21], ct[22], ct[23], K, a, b);
      const double r0 = w * rx, r1 = w * ry;
      int q = 1;
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = r; c < 6; ++c) acc[q++] += a[r] * a[c] + b[r] * b[c];
#pragma unroll
      for (int r = 0; r < 6; ++r) acc[22 + r] += a[r] * r0 + b[r] * r1;
    } else {
      acc[1] += isbad;
    }
  }
  block_reduce<NV, NT_POSE>(acc, sm, out);
}

__global__ void __launch_bounds__(NT_POSE)
k_pose_only(const int batch, double* __restrict__ cams, const int* __restrict__ offset, const double* __restrict__ Xall,
            const double* __restrict__ uvall, const Intr K, const PoseOpts O, uint8_t* __restrict__ usable,
            int* __restrict__ n_iters_out, double* __restrict__ final_cost, int* __restrict__ term_out, const PoseTrace tr) {
  pdl_grid_sync();
  __shared__ double sm[29 * NT_POSE / 32];
  __shared__ double red[32];
  __shared__ double ct[CAMTAB], ctc[CAMTAB];
  __shared__ double s_cam[6], s_camc[6], s_H[36], s_g[6], s_scale[6], s_lam[6], s_y[6];
  __shared__ double s_cost, s_radius, s_dec, s_gmax, s_xnorm, s_model;
  __shared__ int s_state;   // 0 = compute step, 1 = stop
  __shared__ int s_it, s_invalid, s_term, s_stop;
  const int pb = blockIdx.x;
  if (pb >= batch) return;
  const int o0 = offset[pb], n = offset[pb + 1] - o0;
  const double* X = Xall + 3 * (size_t)o0;
  const double* uv = uvall + 2 * (size_t)o0;
  const size_t tb = (size_t)pb * tr.stride;
  if (threadIdx.x < 6) s_cam[threadIdx.x] = cams[6 * (size_t)pb + threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) {
    cam_table_row(s_cam, ct);
    s_radius = O.initial_radius; s_dec = 2.0; s_it = 0; s_invalid = 0; s_state = 0; s_term = 1; s_stop = 0;
  }
  __syncthreads();
  bool need_lin = true, first = true;
  for (;;) {
    if (need_lin) {
      pose_eval<true>(ct, n, X, uv, K, O.loss, sm, red);
      if (threadIdx.x == 0) {
        // H = T' A T, g = T' ghat
        double Af[36], T[36], AT[36];
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) Af[r * 6 + c] = (r <= c) ? red[1 + tri(r, c)] : red[1 + tri(c, r)];
        for (int q = 0; q < 36; ++q) T[q] = 0.0;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { T[r * 6 + c] = ct[CT_G + r * 3 + c]; T[(3 + r) * 6 + 3 + c] = ct[r * 3 + c]; }
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double s = 0; for (int k = 0; k < 6; ++k) s += Af[r * 6 + k] * T[k * 6 + c]; AT[r * 6 + c] = s; }
        for (int r = 0; r < 6; ++r) for (int c = 0; c < 6; ++c) { double s = 0; for (int k = 0; k < 6; ++k) s += T[k * 6 + r] * AT[k * 6 + c]; s_H[r * 6 + c] = s; }
        double gmax = 0, xn = 0;
        for (int r = 0; r < 6; ++r) {
          double s = 0; for (int k = 0; k < 6; ++k) s += T[k * 6 + r] * red[22 + k];
          s_g[r] = s; gmax = fmax(gmax, fabs(s)); xn += s_cam[r] * s_cam[r];
          if (first) s_scale[r] = O.jacobi ? 1.0 / (1.0 + sqrt(s_H[r * 7])) : 1.0;
          const double s2 = s_scale[r] * s_scale[r];
          s_lam[r] = fmin(fmax(s2 * s_H[r * 7], O.min_diag), O.max_diag) / s2;
        }
        s_cost = red[0]; s_gmax = gmax; s_xnorm = sqrt(xn);
        if (first) {
          if (red[28] > 0.0 || !isfinite(red[0])) { s_state = 1; s_term = 2; s_stop = 7; }
          if (tr.cost) { tr.cost[tb] = red[0]; tr.cost_candidate[tb] = red[0]; tr.radius[tb] = s_radius;
                         tr.gradient_max_norm[tb] = gmax; tr.accepted[tb] = 0; tr.step_norm[tb] = 0; tr.relative_decrease[tb] = 0; }
        }
      }
      __syncthreads();
      need_lin = false; first = false;
    }
    if (threadIdx.x == 0 && s_state == 0) {
      // loop guards (FinalizeIterationAndCheckIfMinimizerCanContinue)
      if (s_it >= O.max_iters) { s_state = 1; s_term = 1; s_stop = 1; }
      else if (s_gmax <= O.gradient_tol) { s_state = 1; s_term = 0; s_stop = 2; }
      else if (s_radius <= O.min_radius) { s_state = 1; s_term = 0; s_stop = 5; }
      else {
        s_it += 1;
        double M[36], Mi[36];
        for (int q = 0; q < 36; ++q) M[q] = s_H[q];
        for (int r = 0; r < 6; ++r) M[r * 7] += s_lam[r] / s_radius;
        const bool ok = inv6_spd(M, Mi);
        double yg = 0, yly = 0;
        for (int r = 0; r < 6; ++r) {
          double s = 0; for (int c = 0; c < 6; ++c) s += Mi[r * 6 + c] * s_g[c];
          s_y[r] = s; yg += s * s_g[r]; yly += s_lam[r] / s_radius * s * s;
          s_camc[r] = s_cam[r] - s;
        }
        s_model = ok ? 0.5 * (yg + yly) : -1.0;
        cam_table_row(s_camc, ctc);
      }
    }
    __syncthreads();
    if (s_state) break;
    const int it = s_it;
    if (!(s_model > 0.0)) {      // invalid step (uniform branch: s_model is shared)
      if (threadIdx.x == 0) {
        s_invalid += 1;
        if (tr.cost) { tr.cost[tb + it] = s_cost; tr.cost_candidate[tb + it] = s_cost; tr.step_norm[tb + it] = 0; tr.relative_decrease[tb + it] = 0;
                       tr.gradient_max_norm[tb + it] = s_gmax; tr.accepted[tb + it] = 0; }
        if (s_invalid >= O.max_invalid) { s_state = 1; s_term = 2; s_stop = 6; }
        else { s_radius = s_radius / s_dec; s_dec *= 2.0; }
        if (tr.cost) tr.radius[tb + it] = s_radius;
      }
      __syncthreads();
      if (s_state) break;
      continue;
    }
    pose_eval<false>(ctc, n, X, uv, K, O.loss, sm, red);
    if (threadIdx.x == 0) {
      s_invalid = 0;
      double cand = red[0];
      if (red[1] > 0.0 || !isfinite(cand)) cand = DBL_MAX;
      double sn = 0; for (int r = 0; r < 6; ++r) sn += s_y[r] * s_y[r];
      sn = sqrt(sn);
      double rel = 0.0; int acc = 0;
      if (sn <= O.parameter_tol * (s_xnorm + O.parameter_tol)) { s_state = 1; s_term = 0; s_stop = 3; }
      else if (fabs(s_cost - cand) <= O.function_tol * s_cost) { s_state = 1; s_term = 0; s_stop = 4; }
      else {
        rel = (cand >= DBL_MAX) ? -DBL_MAX : (s_cost - cand) / s_model;
        if (rel > O.min_relative_decrease) {
          acc = 1;
          for (int r = 0; r < 6; ++r) s_cam[r] = s_camc[r];
          for (int q = 0; q < CAMTAB; ++q) ct[q] = ctc[q];
          const double t = 2.0 * rel - 1.0;
          s_radius = fmin(O.max_radius, s_radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
          s_dec = 2.0;
        } else { s_radius = s_radius / s_dec; s_dec *= 2.0; }
      }
      if (tr.cost) { tr.cost_candidate[tb + it] = cand; tr.step_norm[tb + it] = sn; tr.relative_decrease[tb + it] = rel;
                     tr.accepted[tb + it] = (uint8_t)acc; tr.radius[tb + it] = s_radius; tr.cost[tb + it] = acc ? cand : s_cost;
                     tr.gradient_max_norm[tb + it] = s_gmax; }
      s_model = acc ? 1.0 : 0.0;   // reuse as the "accepted" broadcast
    }
    __syncthreads();
    if (s_state) break;
    need_lin = (s_model > 0.5);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const bool ok = (s_term != 2);   // IsSolutionUsable(): CONVERGENCE or NO_CONVERGENCE
    if (ok) for (int r = 0; r < 6; ++r) cams[6 * (size_t)pb + r] = s_cam[r];
    if (usable) usable[pb] = ok ? 1 : 0;
    if (n_iters_out) n_iters_out[pb] = s_it;
    if (final_cost) final_cost[pb] = s_cost;
    if (term_out) { term_out[2 * pb] = s_term; term_out[2 * pb + 1] = s_stop; }
  }
}

}  // namespace glba
