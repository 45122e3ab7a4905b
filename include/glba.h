/*
 * glba.h — C ABI of the B200-native bundle-adjustment backend for GL-SLAM.
 *
 * This is the drop-in boundary for the two BA entry points of the reference
 * (all file:line citations are relative to the GL-SLAM tree):
 *
 *   slam_core::full_ba(std::mutex&, Map&, cv::Mat& K, int window)  -> bool
 *       include/core/slam_core.h:59, src/core/slam_core.cpp:744-883
 *       (called from thread_pool::map_optimizing_thread, src/threading/thread_pool.cpp:351)
 *   slam_core::pose_only_ba(cv::Mat& R, cv::Mat& t, p3d, p2d, K)     -> bool
 *       include/core/slam_core.h:65-68, src/core/slam_core.cpp:1092-1140
 *       (called from thread_pool::tracking_thread, src/threading/thread_pool.cpp:195)
 *
 * The reference has no FFI: those two C++ functions hand flat double arrays to
 * Ceres (camera_params / point_params, slam_core.cpp:750-751, 1099) and read them
 * back (slam_core.cpp:859-871, 1135-1137).  This header exposes exactly that flat
 * hand-off: a camera block is [angle-axis(3), camera centre(3)] (camera-to-world,
 * slam_core.cpp:771-776), a point block is [X,Y,Z] (slam_core.cpp:794-796), an
 * observation is (camera index, point index, u, v) (slam_core.cpp:806-817).
 * include/glba_slam.hpp restates the packing/unpacking around it.
 *
 * Conventions: plain C, no exceptions cross the boundary, every entry point
 * returns GLBA_OK (0) or a negative glba_status; on failure the caller's
 * cam/pt arrays are left untouched.  A glba_ctx is NOT thread-safe; create one
 * per calling thread (the mapping thread and the tracking thread each own one,
 * both may be in flight simultaneously on different CUDA streams).
 */
#ifndef GLBA_H_
#define GLBA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLBA_VERSION 100          /* 0.1.0 */
#define GLBA_MAX_ITERS 256        /* capacity of the per-iteration summary arrays */
#define GLBA_NCCL_ID_BYTES 128

typedef enum {
  GLBA_OK = 0,
  GLBA_E_INVALID_ARG = -1,       /* null pointer, negative size, index out of range */
  GLBA_E_CUDA = -2,              /* CUDA runtime error (glba_last_error has the text) */
  GLBA_E_NO_DEVICE = -3,         /* no CUDA device: there is NO CPU fallback */
  GLBA_E_OOM = -4,
  GLBA_E_NCCL = -5,
  GLBA_E_NUMERIC = -6,           /* initial evaluation produced non-finite residuals */
  GLBA_E_UNSUPPORTED = -7
} glba_status;

typedef enum { GLBA_LOSS_NONE = 0, GLBA_LOSS_HUBER = 1, GLBA_LOSS_CAUCHY = 2 } glba_loss;

/* Linear solver for the reduced camera system. */
typedef enum {
  GLBA_LINSOLVE_AUTO = 0,        /* dense when 6*n_free_cam <= dense_max_dim, else PCG */
  GLBA_LINSOLVE_PCG = 1,         /* block-Jacobi PCG on the implicit Schur complement */
  GLBA_LINSOLVE_DENSE = 2        /* explicit S + on-device Cholesky (small windows)   */
} glba_linsolve;

/* ceres::TerminationType equivalents (slam_core.cpp:1132 IsSolutionUsable()). */
typedef enum {
  GLBA_TERM_CONVERGENCE = 0,
  GLBA_TERM_NO_CONVERGENCE = 1,
  GLBA_TERM_FAILURE = 2
} glba_termination;

/* Why the loop stopped (finer than glba_termination). */
typedef enum {
  GLBA_STOP_NONE = 0,
  GLBA_STOP_MAX_ITERS = 1,
  GLBA_STOP_GRADIENT_TOL = 2,
  GLBA_STOP_PARAMETER_TOL = 3,
  GLBA_STOP_FUNCTION_TOL = 4,
  GLBA_STOP_MIN_RADIUS = 5,
  GLBA_STOP_INVALID_STEPS = 6,
  GLBA_STOP_NUMERIC = 7,
  GLBA_STOP_TRIALS = 8           /* GLBA_MODE_G2O: g2o_max_trials rejected trials in a row, or a trial with rho == 0 ("Terminate") */
} glba_stop_reason;

/* Which of the reference's two BA formulations:
 *  CERES — the live path, slam_core.cpp:699-849: cam = [angle-axis of R_wc | centre], additive update, residual =
 *          projection - observation, Ceres trust-region LM (Jacobi scaling, D = sqrt(clamp(diag)/radius), tolerances).
 *  G2O   — the archived path, Old/mult_img_recoverpose_single_ba:258-314, docs/old_unorganized/4image_pnp_ba.txt:321-430:
 *          cam = [angle-axis of R_cw | t] WORLD-TO-CAMERA (VertexSE3Expmap), update T <- exp([dw, dv]) * T, error =
 *          observation - projection with unit information, robust kernel as rho' reweighting, g2o's Levenberg: (H + lambda I),
 *          lambda0 = g2o_tau * max diag H, rho = (chi2 - chi2') / (d'(lambda d + b) + 1e-3), accept on rho > 0 with
 *          lambda *= clamp(1 - (2 rho - 1)^3, 1/3, 2/3), else lambda *= nu, nu *= 2, at most g2o_max_trials trials per
 *          iteration; max_iters counts g2o iterations (= accepted steps), no tolerance stops.  In the summary every TRIAL
 *          is one entry (radius[] = 1 / lambda), n_successful = completed g2o iterations.  The Ceres-only option fields
 *          (tolerances, radii, lm diagonal bounds, jacobi_scaling, min_relative_decrease) are ignored.
 *          Applies to glba_solve / glba_load + glba_solve_resident / glba_map_solve_window (a map holds poses in the
 *          convention of the mode it is solved with); glba_linearize, pose-only BA and culling are CERES-convention only. */
typedef enum { GLBA_MODE_CERES = 0, GLBA_MODE_G2O = 1 } glba_mode;

typedef enum { GLBA_MEM_HOST = 0, GLBA_MEM_DEVICE = 1 } glba_memspace;

typedef struct glba_ctx glba_ctx;

typedef struct {
  int32_t device;                /* CUDA ordinal */
  int32_t rank;                  /* this context's shard, 0 <= rank < world */
  int32_t world;                 /* 1 = single GPU; >1 = point tracks partitioned over `world` contexts */
  const void* nccl_unique_id;    /* GLBA_NCCL_ID_BYTES bytes from glba_nccl_unique_id(), same on every rank; NULL if world==1 */
  void* stream;                  /* optional cudaStream_t to run on; NULL = context-owned non-blocking stream */
} glba_device_cfg;

/*
 * A BA problem in the flat form the reference hands to Ceres.
 *   cam[6*i..]  = [w0,w1,w2, c0,c1,c2]: angle-axis of R_wc and camera centre (in/out)
 *   pt[3*j..]   = world point (in/out)
 *   observation k ties camera obs_cam[k] to point obs_pt[k] with pixel (obs_u[k], obs_v[k]).
 * Observations may come in any order; track-contiguous (sorted by point, then camera)
 * input skips the on-device sort.  With world>1 every rank passes ALL cameras and its
 * OWN shard of points/observations (whole tracks; point indices are local to the shard).
 * cam_fixed / pt_fixed: non-zero = held constant (slam_core.cpp:831-833); NULL = none.
 */
typedef struct {
  int32_t n_cam;
  int32_t n_pt;
  int64_t n_obs;
  double* cam;
  double* pt;
  const int32_t* obs_cam;
  const int32_t* obs_pt;
  const double* obs_u;
  const double* obs_v;
  const uint8_t* cam_fixed;
  const uint8_t* pt_fixed;
  double fx, fy, cx, cy;         /* cameraMatrix(0,0),(1,1),(0,2),(1,2) — slam_core.cpp:720-723 */
  int32_t memspace;              /* glba_memspace of every pointer above and below */
  const double* pt_info;         /* optional [n_pt]: information weight of every observation of point j (residual' Omega residual with
                                  * Omega = pt_info[j] * I, robust kernel applied to the weighted square); NULL = 1.  The archived g2o BA
                                  * uses 1 / z^2 of the point (docs/old_unorganized/4image_pnp_ba.txt:400-403). */
} glba_problem;

/* Defaults (glba_default_options) = the values hard-coded at slam_core.cpp:814, 842-847
 * plus the Ceres 2.x defaults that call relies on. */
typedef struct {
  int32_t loss;                  /* glba_loss; reference: CauchyLoss(1.0) */
  double loss_scale;             /* a in rho(s); 1.0 */
  int32_t max_iters;             /* 30 */
  double function_tol;           /* 1e-6 */
  double gradient_tol;           /* 1e-10 */
  double parameter_tol;          /* 1e-8 */
  double initial_radius;         /* 1e4 */
  double max_radius;             /* 1e16 */
  double min_radius;             /* 1e-32 */
  double min_relative_decrease;  /* 1e-3 */
  double min_lm_diagonal;        /* 1e-6 */
  double max_lm_diagonal;        /* 1e32 */
  int32_t jacobi_scaling;        /* 1 */
  int32_t max_consecutive_invalid_steps; /* 5 */
  int32_t linsolve;              /* glba_linsolve */
  int32_t dense_max_dim;         /* AUTO switches to PCG above this reduced dimension (default 96 = 16 cameras, the dense kernels' limit) */
  double cg_rel_tol;             /* PCG stop: sqrt(r'M^-1 r) <= tol * sqrt(r0'M^-1 r0); 1e-13 = parity mode */
  int32_t cg_max_iters;          /* 0 = 4*reduced dimension, capped at 4000 */
  int32_t verbose;
  int32_t mode;                  /* glba_mode; GLBA_MODE_CERES */
  double g2o_tau;                /* 1e-5  (OptimizationAlgorithmLevenberg::_tau) */
  int32_t g2o_max_trials;        /* 10    (maxTrialsAfterFailure) */
} glba_options;

typedef struct {
  int32_t status;                /* glba_status */
  int32_t termination;           /* glba_termination */
  int32_t stop_reason;           /* glba_stop_reason */
  int32_t n_iters;               /* LM iterations started (trust-region steps computed) */
  int32_t n_successful;          /* accepted steps */
  int32_t n_linearizations;      /* Jacobian evaluations (1 + n_successful) */
  double initial_cost;
  double final_cost;
  /* index 0 = state before the first step; index it = state after LM iteration it */
  double cost[GLBA_MAX_ITERS + 1];            /* cost of the current (accepted) point */
  double cost_candidate[GLBA_MAX_ITERS + 1];  /* cost evaluated at the trial point of iteration it */
  double radius[GLBA_MAX_ITERS + 1];          /* trust-region radius after the iteration */
  double step_norm[GLBA_MAX_ITERS + 1];
  double relative_decrease[GLBA_MAX_ITERS + 1];
  double gradient_max_norm[GLBA_MAX_ITERS + 1];
  int32_t cg_iters[GLBA_MAX_ITERS + 1];
  uint8_t accepted[GLBA_MAX_ITERS + 1];
  /* device time, milliseconds, summed over the solve.  The per-phase figures are recorded only for problems with
   * >= 200 000 observations: below that the CUDA-event calls themselves rival the kernels and are skipped (fields = 0). */
  double t_setup_ms;             /* H2D + sort + index build */
  double t_linearize_ms;         /* residual/weight/Jacobian records + Hessian blocks */
  double t_schur_ms;             /* point-block inverses, preconditioner, reduced rhs */
  double t_solve_ms;             /* PCG / dense solve */
  double t_update_ms;            /* back-substitution, candidate cost, accept/reject */
  double t_total_ms;
  double t_comm_ms;              /* world > 1: the fused all-reduce of every linearisation with its chunk sum, measured in place
                                  * (so it includes waiting for the slowest rank); part of t_total_ms, not of the fields above */
} glba_summary;

/* Output of glba_linearize: the reduced camera system of ONE linearisation at a given radius.
 * Any pointer may be NULL (not copied out).  Host pointers. */
typedef struct {
  double cost;                   /* 1/2 sum rho(|r|^2) */
  double* residuals;             /* [2*n_obs]  loss-corrected r~, caller's observation order */
  double* jac_cam;               /* [12*n_obs] loss-corrected 2x6 block, row-major */
  double* jac_pt;                /* [6*n_obs]  loss-corrected 2x3 block, row-major */
  double* grad_cam;              /* [6*n_cam]  J~c' r~ (zero rows for fixed cameras) */
  double* grad_pt;               /* [3*n_pt] */
  double* hess_cam;              /* [36*n_cam] B_i = sum J~c'J~c, row-major, undamped */
  double* hess_pt;               /* [9*n_pt]   C_j = sum J~p'J~p, undamped */
  double* schur_diag;            /* [36*n_cam] diagonal blocks of S = B + D_c^2 - W (C+D_p^2)^-1 W' */
  double* schur_rhs;             /* [6*n_cam]  g_c - W (C+D_p^2)^-1 g_p */
  double t_linearize_ms;
  double t_schur_ms;
} glba_linearization;

/* Average device time per launch (CUDA events on the context stream) of each hot kernel on the
 * resident problem, for roofline accounting (bench.py).  Milliseconds. */
typedef struct {
  double linearize_pm_ms;   /* K_A + point half of K_B: residual, weight, records, C_j, g_j, damped inverse */
  double linearize_cm_ms;   /* camera half of K_B: B_i, g_i */
  double schur_cm_ms;       /* Schur diagonal blocks (preconditioner) + reduced rhs */
  double spmv_pm_ms;        /* implicit S x, point-major half */
  double spmv_cm_ms;        /* implicit S x, camera-major half */
  double backsub_cost_ms;   /* back-substitution + candidate cost */
  double point_damp_ms;     /* re-damping of the point blocks after a rejected step */
  double small_kernels_ms;  /* all camera-sized / reduction kernels of one linearise+Schur pass */
  double allreduce_ms;      /* world > 1: the one fused NCCL all-reduce of a linearise+Schur pass ([accB|accA|scalars]) */
  double chunk_sum_ms;      /* world > 1: the two per-camera chunk sums that feed it */
  double exchange_bytes;    /* world > 1: payload of that all-reduce on this rank */
  int32_t n_local_cams;     /* world > 1: cameras this rank holds (owner-computes layout: the ones its own tracks observe) */
  int32_t n_shared_cams;    /* world > 1: cameras whose partial sums are exchanged (owner-computes: observed by >= 2 ranks) */
  /* explicit block-sparse reduced camera matrix (single GPU, glba_sparse.cuh); zeros when the map keeps the matrix-free product */
  double schur_pairs_ms;    /* assembly of the off-diagonal blocks, once per LM iteration */
  double bsr_spmv_ms;       /* one PCG iteration of the cooperative kernel on the blocks (product + dot products + updates + two grid barriers) */
  double pair_setup_ms;     /* structure of the loaded problem (sorts), once per load; 0 if it was already built */
  int64_t n_pair_instances; /* (observation, observation) pairs summed by the assembly */
  int32_t n_pair_blocks;    /* distinct off-diagonal 6x6 blocks (upper triangle) */
  int32_t reserved_;
  double cam_pipe_ms;       /* large maps: both camera-major passes of a linearisation in one kernel (k_cam_pipe); 0 if the map runs them separately */
} glba_kernel_times;

void glba_default_options(glba_options* opt);
const char* glba_strerror(int status);
const char* glba_last_error(const glba_ctx* ctx);
int glba_version(void);
/* Number of kernels this library has launched since the process started (bench evidence). */
int64_t glba_kernel_launch_count(void);

int glba_nccl_unique_id(void* out_id /* GLBA_NCCL_ID_BYTES */);
int glba_create(const glba_device_cfg* cfg, glba_ctx** out);
void glba_destroy(glba_ctx* ctx);

/* Replaces the ceres::Problem build + ceres::Solve of slam_core.cpp:799-849. */
int glba_solve(glba_ctx* ctx, const glba_problem* prob, const glba_options* opt, glba_summary* summary);

/* Replaces slam_core.cpp:1099-1137.  cam[6] in/out; X[3*n], uv[2*n] host arrays (points fixed). */
int glba_pose_only(glba_ctx* ctx, double* cam, int32_t n, const double* X, const double* uv,
                   double fx, double fy, double cx, double cy, const glba_options* opt,
                   glba_summary* summary);

/* `batch` independent pose-only problems in one launch; problem b owns
 * observations [offset[b], offset[b+1]).  usable[b] = IsSolutionUsable(). */
int glba_pose_only_batch(glba_ctx* ctx, int32_t batch, double* cams /* [6*batch] */,
                         const int32_t* offset /* [batch+1] */, const double* X, const double* uv,
                         double fx, double fy, double cx, double cy, const glba_options* opt,
                         uint8_t* usable, int32_t* n_iters, double* final_cost);

/* One pass of residual + robust weight + Jacobians + Hessian blocks + Schur complement
 * (the kernels the headline metric times), state is NOT updated. */
int glba_linearize(glba_ctx* ctx, const glba_problem* prob, const glba_options* opt, double radius,
                   glba_linearization* out);

/* Benchmark/serving form: upload + index once, then run passes on the resident problem. */
int glba_load(glba_ctx* ctx, const glba_problem* prob, const glba_options* opt);
/* cost == NULL: enqueue only (no host synchronisation) */
int glba_linearize_resident(glba_ctx* ctx, const glba_options* opt, double radius, double* cost /* may be NULL */);
int glba_solve_resident(glba_ctx* ctx, const glba_options* opt, glba_summary* summary);
int glba_time_kernels(glba_ctx* ctx, const glba_options* opt, double radius, int32_t reps, glba_kernel_times* out);
int glba_reset_resident(glba_ctx* ctx);     /* restore the parameters uploaded by glba_load */
int glba_read_resident(glba_ctx* ctx, double* cam, double* pt); /* D2H of the current state */
int glba_synchronize(glba_ctx* ctx);
void* glba_stream(glba_ctx* ctx);            /* cudaStream_t the context launches on */

/* Post-BA map maintenance on device (slam_core.cpp:977-1038, post_ba_map_point_culling):
 * mean reprojection error and cheirality per point over ALL its observations.
 * bad[j] = 1 if any depth <= 0, or n_obs_j < min_obs, or mean error > max_mean_err. */
int glba_cull_points(glba_ctx* ctx, const glba_problem* prob, int32_t min_obs, double max_mean_err,
                     uint8_t* bad /* [n_pt] */, double* mean_err /* [n_pt], may be NULL */);

/* Two-view triangulation + cheirality / reprojection filter on device, one thread per match
 * (slam_core.cpp:173-256, triangulate_and_filter_3d_points; cv::triangulatePoints = DLT).  [R|t] are world-to-camera
 * as that function receives them (row-major R1[9], t1[3], ...); p0/p1 = matched pixels [2n]; X[3n] receives X/w for every
 * match, keep[n] = 1 where the reference would have kept the match (|w| >= 1e-9, 0 < depth <= distance_threshold in both
 * views, reprojection error <= reprojection_threshold in both views).  Host pointers. */
int glba_triangulate_filter(glba_ctx* ctx, const double* R1, const double* t1, const double* R2, const double* t2,
                            double fx, double fy, double cx, double cy, int32_t n, const double* p0, const double* p1,
                            double distance_threshold, double reprojection_threshold, double* X, uint8_t* keep);

/* ---- Persistent device-resident map -------------------------------------------------------------------
 * An append-only SoA mirror of slam_types.h:13-61 kept in HBM, so that a local BA no longer re-walks the host hash maps
 * and re-uploads the window (slam_core.cpp:750-819).  Feed it where GL-SLAM grows its map
 * (update_map_and_keyframe_data, slam_core.cpp:287-426): keyframe ids and point ids are the dense counters GL-SLAM
 * already uses (run_window / next_point_id), handed back in *first_id.  One map per context, one GPU; host pointers.
 * The map must be destroyed before its context. */
typedef struct glba_map glba_map;
int glba_map_create(glba_ctx* ctx, double fx, double fy, double cx, double cy, glba_map** out);
void glba_map_destroy(glba_map* map);
int glba_map_size(const glba_map* map, int32_t* n_kf, int32_t* n_pt, int64_t* n_obs);
int glba_map_add_keyframes(glba_map* map, int32_t n, const double* cam /* [6n] camera-to-world (w, centre) */, int32_t* first_id);
int glba_map_add_points(glba_map* map, int32_t n, const double* xyz /* [3n] */, int32_t* first_id);
int glba_map_add_observations(glba_map* map, int32_t n, const int32_t* kf, const int32_t* pt, const double* uv /* [2n] u,v interleaved */);
int glba_map_set_bad(glba_map* map, int32_t n, const int32_t* pt_ids, uint8_t value);   /* MapPoint::is_bad */
int glba_map_write_keyframes(glba_map* map, int32_t first, int32_t n, const double* cam);   /* host edits, e.g. slam_core.cpp:916-973 */
int glba_map_read_keyframes(glba_map* map, int32_t first, int32_t n, double* cam);
int glba_map_write_points(glba_map* map, int32_t first, int32_t n, const double* xyz);
int glba_map_read_points(glba_map* map, int32_t first, int32_t n, double* xyz, uint8_t* bad /* may be NULL */);
/* full_ba (slam_core.cpp:744-883) on the resident map: keyframes [first_kf, first_kf+window), the first n_fixed constant
 * (2 in the reference), all non-bad points with >= min_obs observations inside the window (reference: 1).  Refined poses
 * and points are written into the map unless the solve FAILED.  n_pt_used / n_obs_used (may be NULL) = window size. */
int glba_map_solve_window(glba_map* map, int32_t first_kf, int32_t window, int32_t n_fixed, int32_t min_obs,
                          const glba_options* opt, glba_summary* summary, int32_t* n_pt_used, int64_t* n_obs_used);
/* post_ba_map_point_culling (slam_core.cpp:977-1038) on the resident map: candidates = non-bad points first observed by
 * keyframes [first_kf, last_kf] (reference: [run_window - local_ba_window, run_window - 4]); a candidate becomes bad if it
 * lies behind one of its cameras, has < min_obs observations, or a mean reprojection error > max_mean_err over ALL its
 * observations.  The flags are set in the map; up to `cap` culled point ids are written to culled_ids (may be NULL). */
int glba_map_cull_points(glba_map* map, int32_t first_kf, int32_t last_kf, int32_t min_obs, double max_mean_err,
                         int32_t* n_candidates, int32_t* n_culled, int32_t* culled_ids, int32_t cap);

/* post_ba_map_update_for_new_keyframes (slam_core.cpp:916-973) on the resident map.  (R_before[9] row-major, t_before[3]) =
 * pose of keyframe kf_last saved before the BA write-back (slam_core.cpp:853-854); its pose NOW in the map is "after".
 * delta = ComputeDeltaPose_SO3 (slam_core.cpp:885-912: SVD projection to SO(3) with the reflection flip) is applied on the
 * device to keyframes kf_ids (R <- dR R, t <- dR t + dt: kpid_to_correct) and points pt_ids (X <- dR X + dt:
 * mpid_to_correct).  dR_out[9] / dt_out[3] may be NULL.  Host pointers. */
int glba_map_propagate(glba_map* map, const double* R_before, const double* t_before, int32_t kf_last, int32_t n_kf,
                       const int32_t* kf_ids, int32_t n_pt, const int32_t* pt_ids, double* dR_out, double* dt_out);

#ifdef __cplusplus
}
#endif
#endif /* GLBA_H_ */
