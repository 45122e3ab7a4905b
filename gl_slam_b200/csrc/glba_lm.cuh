// glba_lm.cuh — the trust-region decisions of a Levenberg-Marquardt iteration ON THE DEVICE (north star (4): "LM damping and
// accept/reject step, all on device").
//
// Ceres' TrustRegionMinimizer loop as run_lm (glba.cu) restates it, cut where the data is: k_lm_decide consumes the scalars
// of a computed step (model decrease, candidate cost, step norm) and writes accept / reject / invalid, the new radius and the
// termination tests into an LmCtl in device memory; k_lm_absorb does the same for the scalars of the re-linearisation that
// follows an accepted step.  Every kernel of the iteration reads that LmCtl (ctl_skip / ctl_inv_radius, glba_kernels.cuh), so
// the host enqueues iterations blindly and synchronises once per few iterations only to learn whether the loop is done.
// The per-iteration trace goes straight into a glba_summary in device memory.
#pragma once
#include "../../include/glba.h"
#include "glba_kernels.cuh"

namespace glba {

struct LmParams {
  int max_iters, max_invalid, n_free_cam;
  double function_tol, gradient_tol, parameter_tol, max_radius, min_radius, min_rel;
};

__device__ __forceinline__ void lm_finish(LmCtl* c, glba_summary* sum, const int termination, const int stop_reason) {
  c->done = 1; c->accepted = 0; c->need_redamp = 0;
  sum->termination = termination; sum->stop_reason = stop_reason;
  sum->n_iters = c->it; sum->final_cost = c->cost;
}
// the tests Ceres makes before it starts another iteration (run_lm's loop top), in its order
__device__ __forceinline__ void lm_top_tests(LmCtl* c, const LmParams& P, glba_summary* sum) {
  if (c->it >= P.max_iters) lm_finish(c, sum, GLBA_TERM_NO_CONVERGENCE, GLBA_STOP_MAX_ITERS);
  else if (c->gmax <= P.gradient_tol) lm_finish(c, sum, GLBA_TERM_CONVERGENCE, GLBA_STOP_GRADIENT_TOL);
  else if (c->radius <= P.min_radius) lm_finish(c, sum, GLBA_TERM_CONVERGENCE, GLBA_STOP_MIN_RADIUS);
}

// After back-substitution + candidate cost: step valid?  converged?  accept or reject, new radius.
__device__ inline void lm_decide(LmCtl* c, const LmParams& P, const double* S, glba_summary* sum) {
  if (c->done) return;
  const int it = ++c->it;
  c->accepted = 0; c->need_redamp = 0;
  sum->cg_iters[it] = 0;
  const bool solver_ok = (S[S_NOTPD_P] + (P.n_free_cam > 0 ? S[S_NOTPD_C] : 0.0)) == 0.0;
  const double model_cost_change = 0.5 * ((S[S_YG_P] + S[S_YG_C]) + (S[S_YLY_P] + S[S_YLY_C]));
  const bool valid = solver_ok && (model_cost_change > 0.0);
  double radius = c->radius;
  if (!valid) {
    ++c->n_invalid;
    sum->cost[it] = c->cost; sum->cost_candidate[it] = c->cost; sum->step_norm[it] = 0.0; sum->relative_decrease[it] = 0.0;
    sum->gradient_max_norm[it] = c->gmax; sum->accepted[it] = 0;
    if (c->n_invalid >= P.max_invalid) { sum->radius[it] = radius; lm_finish(c, sum, GLBA_TERM_FAILURE, GLBA_STOP_INVALID_STEPS); return; }
    radius = radius / c->decrease_factor; c->decrease_factor *= 2.0; c->need_redamp = 1;
    sum->radius[it] = radius;
    c->radius = radius; c->inv_radius = 1.0 / radius;
    lm_top_tests(c, P, sum);
    return;
  }
  c->n_invalid = 0;
  double cand = S[S_COST_C];
  if (S[S_BAD_C] > 0.0 || !isfinite(cand) || !solver_ok) cand = DBL_MAX;
  const double step_norm = sqrt(S[S_YN2_P] + S[S_YN2_C]);
  sum->cost_candidate[it] = cand; sum->step_norm[it] = step_norm; sum->cost[it] = c->cost; sum->radius[it] = radius;
  sum->gradient_max_norm[it] = c->gmax;
  if (step_norm <= P.parameter_tol * (c->x_norm + P.parameter_tol)) { lm_finish(c, sum, GLBA_TERM_CONVERGENCE, GLBA_STOP_PARAMETER_TOL); return; }
  const double cost_change = c->cost - cand;
  if (fabs(cost_change) <= P.function_tol * c->cost) { lm_finish(c, sum, GLBA_TERM_CONVERGENCE, GLBA_STOP_FUNCTION_TOL); return; }
  const double rel = (cand >= DBL_MAX) ? -DBL_MAX : cost_change / model_cost_change;
  sum->relative_decrease[it] = rel;
  if (rel > P.min_rel) {
    c->accepted = 1;
    const double t = __dsub_rn(__dmul_rn(2.0, rel), 1.0);      // no FMA contraction: the host loop (run_lm) rounds the same way
    radius = radius / fmax(1.0 / 3.0, __dsub_rn(1.0, __dmul_rn(__dmul_rn(t, t), t)));
    radius = fmin(P.max_radius, radius);
    c->decrease_factor = 2.0; c->n_rejected = 0;
    c->cost = cand;            // provisional: k_lm_absorb replaces it by the re-evaluated value
    sum->n_linearizations++; sum->n_successful++; sum->accepted[it] = 1;
  } else {
    radius = radius / c->decrease_factor; c->decrease_factor *= 2.0; c->need_redamp = 1;
    sum->accepted[it] = 0;
    ++c->n_rejected;
  }
  c->radius = radius; c->inv_radius = 1.0 / radius;
  sum->cost[it] = c->cost; sum->radius[it] = radius; sum->gradient_max_norm[it] = c->gmax;
  if (!c->accepted) lm_top_tests(c, P, sum);      // accepted: k_lm_absorb makes them after the re-linearisation
}

// After the re-linearisation that follows an accepted step: its cost, |g|_inf, |x| become the loop's.
__device__ inline void lm_absorb(LmCtl* c, const LmParams& P, const double* S, glba_summary* sum) {
  if (c->done || !c->accepted) return;
  if (S[S_BAD] > 0.0 || !isfinite(S[S_COST])) { lm_finish(c, sum, GLBA_TERM_FAILURE, GLBA_STOP_NUMERIC); return; }
  c->cost = S[S_COST]; c->gmax = fmax(S[S_GMAX_P], S[S_GMAX_C]); c->x_norm = sqrt(S[S_XN2_P] + S[S_XN2_C]);
  sum->cost[c->it] = c->cost; sum->gradient_max_norm[c->it] = c->gmax;
  sum->final_cost = c->cost;
  lm_top_tests(c, P, sum);
}

// The two decisions ride in the tail of the kernel that produces their last input: the CTA that finishes the scalar
// reduction of the back-substitution pass calls lm_decide, the one that finishes the camera blocks of a re-linearisation
// calls lm_absorb (no extra launches).  LmHook is what those kernels carry.
struct LmHook {
  LmCtl* ctl;              // null: host-driven loop, nothing to do
  glba_summary* sum;
  LmParams P;
};

// accepted step: the candidate state (cameras, their table, points) becomes the current one
__global__ void k_accept_copy(const LmCtl* __restrict__ ctl, const int n_cam, const int n_pt, const double* __restrict__ cam_c,
                              const double* __restrict__ camtab_c, const double4* __restrict__ pt_c, double* __restrict__ cam,
                              double* __restrict__ camtab, double4* __restrict__ pt) {
  pdl_grid_sync();
  if (ctl_skip(ctl, GATE_ACCEPTED)) return;
  const long n1 = 6L * n_cam, n2 = n1 + (long)CAMTAB * n_cam, n3 = n2 + n_pt;
  for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n3; t += (long)gridDim.x * blockDim.x) {
    if (t < n1) cam[t] = cam_c[t];
    else if (t < n2) camtab[t - n1] = camtab_c[t - n1];
    else st4(pt + (t - n2), ldg4(pt_c + (t - n2)));
  }
}

}  // namespace glba
