"""ctypes mirror of include/glba.h (struct layouts, enums, prototypes).

Shared by the product loader (gl_slam_b200/__init__.py -> libglba.so) and by the test-only
checker under oracle/ (which re-uses these POD structs); nothing here depends on that checker.
"""
import ctypes as C

import numpy as np

GLBA_MAX_ITERS = 256
GLBA_NCCL_ID_BYTES = 128

# glba_status
OK, E_INVALID_ARG, E_CUDA, E_NO_DEVICE, E_OOM, E_NCCL, E_NUMERIC, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6, -7
# glba_loss
LOSS_NONE, LOSS_HUBER, LOSS_CAUCHY = 0, 1, 2
# glba_linsolve
LINSOLVE_AUTO, LINSOLVE_PCG, LINSOLVE_DENSE = 0, 1, 2
# glba_termination
TERM_CONVERGENCE, TERM_NO_CONVERGENCE, TERM_FAILURE = 0, 1, 2
# glba_stop_reason
(STOP_NONE, STOP_MAX_ITERS, STOP_GRADIENT_TOL, STOP_PARAMETER_TOL, STOP_FUNCTION_TOL, STOP_MIN_RADIUS,
 STOP_INVALID_STEPS, STOP_NUMERIC) = range(8)
MEM_HOST, MEM_DEVICE = 0, 1
MODE_CERES, MODE_G2O = 0, 1

_N = GLBA_MAX_ITERS + 1


class DeviceCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("nccl_unique_id", C.c_void_p), ("stream", C.c_void_p)]


class Problem(C.Structure):
    _fields_ = [("n_cam", C.c_int32), ("n_pt", C.c_int32), ("n_obs", C.c_int64),
                ("cam", C.c_void_p), ("pt", C.c_void_p),
                ("obs_cam", C.c_void_p), ("obs_pt", C.c_void_p), ("obs_u", C.c_void_p), ("obs_v", C.c_void_p),
                ("cam_fixed", C.c_void_p), ("pt_fixed", C.c_void_p),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("memspace", C.c_int32), ("pt_info", C.c_void_p)]


class Options(C.Structure):
    _fields_ = [("loss", C.c_int32), ("loss_scale", C.c_double), ("max_iters", C.c_int32),
                ("function_tol", C.c_double), ("gradient_tol", C.c_double), ("parameter_tol", C.c_double),
                ("initial_radius", C.c_double), ("max_radius", C.c_double), ("min_radius", C.c_double),
                ("min_relative_decrease", C.c_double), ("min_lm_diagonal", C.c_double),
                ("max_lm_diagonal", C.c_double), ("jacobi_scaling", C.c_int32),
                ("max_consecutive_invalid_steps", C.c_int32), ("linsolve", C.c_int32),
                ("dense_max_dim", C.c_int32), ("cg_rel_tol", C.c_double), ("cg_max_iters", C.c_int32),
                ("verbose", C.c_int32), ("mode", C.c_int32), ("g2o_tau", C.c_double), ("g2o_max_trials", C.c_int32)]


class Summary(C.Structure):
    _fields_ = [("status", C.c_int32), ("termination", C.c_int32), ("stop_reason", C.c_int32),
                ("n_iters", C.c_int32), ("n_successful", C.c_int32), ("n_linearizations", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("cost", C.c_double * _N), ("cost_candidate", C.c_double * _N), ("radius", C.c_double * _N),
                ("step_norm", C.c_double * _N), ("relative_decrease", C.c_double * _N),
                ("gradient_max_norm", C.c_double * _N), ("cg_iters", C.c_int32 * _N),
                ("accepted", C.c_uint8 * _N),
                ("t_setup_ms", C.c_double), ("t_linearize_ms", C.c_double), ("t_schur_ms", C.c_double),
                ("t_solve_ms", C.c_double), ("t_update_ms", C.c_double), ("t_total_ms", C.c_double), ("t_comm_ms", C.c_double)]

    def as_dict(self):
        n = self.n_iters + 1
        d = {k: getattr(self, k) for k in ("status", "termination", "stop_reason", "n_iters", "n_successful",
                                            "n_linearizations", "initial_cost", "final_cost", "t_setup_ms",
                                            "t_linearize_ms", "t_schur_ms", "t_solve_ms", "t_update_ms",
                                            "t_total_ms", "t_comm_ms")}
        for k in ("cost", "cost_candidate", "radius", "step_norm", "relative_decrease", "gradient_max_norm",
                  "cg_iters", "accepted"):
            d[k] = list(getattr(self, k)[:n])
        return d


class Linearization(C.Structure):
    _fields_ = [("cost", C.c_double), ("residuals", C.c_void_p), ("jac_cam", C.c_void_p), ("jac_pt", C.c_void_p),
                ("grad_cam", C.c_void_p), ("grad_pt", C.c_void_p), ("hess_cam", C.c_void_p),
                ("hess_pt", C.c_void_p), ("schur_diag", C.c_void_p), ("schur_rhs", C.c_void_p),
                ("t_linearize_ms", C.c_double), ("t_schur_ms", C.c_double)]


class KernelTimes(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("linearize_pm_ms", "linearize_cm_ms", "schur_cm_ms", "spmv_pm_ms", "spmv_cm_ms",
                                          "backsub_cost_ms", "point_damp_ms", "small_kernels_ms", "allreduce_ms", "chunk_sum_ms",
                                          "exchange_bytes")] + [("n_local_cams", C.c_int32), ("n_shared_cams", C.c_int32)] + \
               [(k, C.c_double) for k in ("schur_pairs_ms", "bsr_spmv_ms", "pair_setup_ms")] + \
               [("n_pair_instances", C.c_int64), ("n_pair_blocks", C.c_int32), ("reserved_", C.c_int32), ("cam_pipe_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def _ptr(a):
    return None if a is None else a.ctypes.data


class HostProblem:
    """Numpy arrays in the layout of glba_problem; keeps them alive while the ctypes struct is in use."""

    def __init__(self, cam, pt, obs_cam, obs_pt, obs_u, obs_v, K, cam_fixed=None, pt_fixed=None, pt_info=None):
        self.cam = np.ascontiguousarray(cam, dtype=np.float64).reshape(-1, 6).copy()
        self.pt = np.ascontiguousarray(pt, dtype=np.float64).reshape(-1, 3).copy()
        self.obs_cam = np.ascontiguousarray(obs_cam, dtype=np.int32)
        self.obs_pt = np.ascontiguousarray(obs_pt, dtype=np.int32)
        self.obs_u = np.ascontiguousarray(obs_u, dtype=np.float64)
        self.obs_v = np.ascontiguousarray(obs_v, dtype=np.float64)
        self.cam_fixed = None if cam_fixed is None else np.ascontiguousarray(cam_fixed, dtype=np.uint8)
        self.pt_fixed = None if pt_fixed is None else np.ascontiguousarray(pt_fixed, dtype=np.uint8)
        self.pt_info = None if pt_info is None else np.ascontiguousarray(pt_info, dtype=np.float64)
        self.K = tuple(float(x) for x in K)  # fx, fy, cx, cy
        n = self.obs_cam.shape[0]
        if not (self.obs_pt.shape[0] == n == self.obs_u.shape[0] == self.obs_v.shape[0]):
            raise ValueError("observation arrays differ in length")

    @property
    def n_cam(self):
        return self.cam.shape[0]

    @property
    def n_pt(self):
        return self.pt.shape[0]

    @property
    def n_obs(self):
        return self.obs_cam.shape[0]

    def copy(self):
        return HostProblem(self.cam, self.pt, self.obs_cam, self.obs_pt, self.obs_u, self.obs_v, self.K,
                           self.cam_fixed, self.pt_fixed, self.pt_info)

    def struct(self):
        p = Problem()
        p.n_cam, p.n_pt, p.n_obs = self.n_cam, self.n_pt, self.n_obs
        p.cam, p.pt = _ptr(self.cam), _ptr(self.pt)
        p.obs_cam, p.obs_pt, p.obs_u, p.obs_v = _ptr(self.obs_cam), _ptr(self.obs_pt), _ptr(self.obs_u), _ptr(self.obs_v)
        p.cam_fixed, p.pt_fixed = _ptr(self.cam_fixed), _ptr(self.pt_fixed)
        p.fx, p.fy, p.cx, p.cy = self.K
        p.memspace = MEM_HOST
        p.pt_info = _ptr(self.pt_info)
        return p


class LinearizationOut:
    """Host buffers for a glba_linearization (outputs of one linearisation)."""

    def __init__(self, n_cam, n_pt, n_obs, per_obs=True):
        z = np.zeros
        self.residuals = z((n_obs, 2)) if per_obs else None
        self.jac_cam = z((n_obs, 2, 6)) if per_obs else None
        self.jac_pt = z((n_obs, 2, 3)) if per_obs else None
        self.grad_cam, self.grad_pt = z((n_cam, 6)), z((n_pt, 3))
        self.hess_cam, self.hess_pt = z((n_cam, 6, 6)), z((n_pt, 3, 3))
        self.schur_diag, self.schur_rhs = z((n_cam, 6, 6)), z((n_cam, 6))
        self.cost = 0.0
        self.t_linearize_ms = self.t_schur_ms = 0.0

    def struct(self):
        s = Linearization()
        for k in ("residuals", "jac_cam", "jac_pt", "grad_cam", "grad_pt", "hess_cam", "hess_pt", "schur_diag",
                  "schur_rhs"):
            setattr(s, k, _ptr(getattr(self, k)))
        return s

    def take(self, s):
        self.cost, self.t_linearize_ms, self.t_schur_ms = s.cost, s.t_linearize_ms, s.t_schur_ms


def make_options(default_fn, **kw):
    o = Options()
    default_fn(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(f"glba_options has no field {k!r}")
        setattr(o, k, v)
    return o
