"""gl_slam_b200 — B200-native bundle-adjustment backend for GL-SLAM (host-side Python binding).

Thin ctypes layer over the C ABI of include/glba.h (gl_slam_b200/libglba.so, built from csrc/ by
`__graft_entry__.build()` or `make -C gl_slam_b200/csrc`).  The library is the product; this module
only marshals numpy buffers.  There is NO CPU fallback: if the shared library is missing, or no CUDA
device is present, calls raise.
"""
import ctypes as C
import weakref
import os

import numpy as np

from . import _abi
from ._abi import (HostProblem, LOSS_CAUCHY, LOSS_HUBER, LOSS_NONE, LINSOLVE_AUTO, LINSOLVE_DENSE,  # noqa: F401
                   LINSOLVE_PCG, TERM_CONVERGENCE, TERM_FAILURE, TERM_NO_CONVERGENCE)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libglba.so")
_LIB = None

# every symbol include/glba.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = (
    "glba_default_options", "glba_strerror", "glba_last_error", "glba_version", "glba_kernel_launch_count",
    "glba_nccl_unique_id", "glba_create", "glba_destroy", "glba_solve", "glba_pose_only", "glba_pose_only_batch",
    "glba_linearize", "glba_load", "glba_linearize_resident", "glba_solve_resident", "glba_reset_resident",
    "glba_read_resident", "glba_synchronize", "glba_stream", "glba_cull_points", "glba_time_kernels",
    "glba_triangulate_filter",
    "glba_map_create", "glba_map_destroy", "glba_map_size", "glba_map_add_keyframes", "glba_map_add_points",
    "glba_map_add_observations", "glba_map_set_bad", "glba_map_write_keyframes", "glba_map_read_keyframes",
    "glba_map_write_points", "glba_map_read_points", "glba_map_solve_window", "glba_map_cull_points", "glba_map_propagate",
)


class GlbaError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        msg = f"{where}: status {status}"
        try:
            msg += f" ({lib().glba_strerror(status).decode()})"
        except Exception:
            pass
        if detail:
            msg += f": {detail}"
        super().__init__(msg)


def lib():
    """Load libglba.so; raises loudly if the CUDA library has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the BA kernels)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
    L.glba_default_options.argtypes = [C.POINTER(_abi.Options)]
    L.glba_default_options.restype = None
    L.glba_strerror.argtypes = [C.c_int]
    L.glba_strerror.restype = C.c_char_p
    L.glba_last_error.argtypes = [vp]
    L.glba_last_error.restype = C.c_char_p
    L.glba_version.restype = C.c_int
    L.glba_kernel_launch_count.restype = C.c_int64
    L.glba_nccl_unique_id.argtypes = [vp]
    L.glba_create.argtypes = [C.POINTER(_abi.DeviceCfg), C.POINTER(vp)]
    L.glba_destroy.argtypes = [vp]
    L.glba_destroy.restype = None
    L.glba_solve.argtypes = [vp, C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.POINTER(_abi.Summary)]
    L.glba_pose_only.argtypes = [vp, vp, i32, vp, vp, f64, f64, f64, f64, C.POINTER(_abi.Options), C.POINTER(_abi.Summary)]
    L.glba_pose_only_batch.argtypes = [vp, i32, vp, vp, vp, vp, f64, f64, f64, f64, C.POINTER(_abi.Options), vp, vp, vp]
    L.glba_linearize.argtypes = [vp, C.POINTER(_abi.Problem), C.POINTER(_abi.Options), f64, C.POINTER(_abi.Linearization)]
    L.glba_load.argtypes = [vp, C.POINTER(_abi.Problem), C.POINTER(_abi.Options)]
    L.glba_linearize_resident.argtypes = [vp, C.POINTER(_abi.Options), f64, C.POINTER(f64)]
    L.glba_solve_resident.argtypes = [vp, C.POINTER(_abi.Options), C.POINTER(_abi.Summary)]
    L.glba_time_kernels.argtypes = [vp, C.POINTER(_abi.Options), f64, i32, C.POINTER(_abi.KernelTimes)]
    L.glba_reset_resident.argtypes = [vp]
    L.glba_read_resident.argtypes = [vp, vp, vp]
    L.glba_synchronize.argtypes = [vp]
    L.glba_stream.argtypes = [vp]
    L.glba_stream.restype = vp
    L.glba_cull_points.argtypes = [vp, C.POINTER(_abi.Problem), i32, f64, vp, vp]
    L.glba_triangulate_filter.argtypes = [vp, vp, vp, vp, vp, f64, f64, f64, f64, i32, vp, vp, f64, f64, vp, vp]
    L.glba_map_create.argtypes = [vp, f64, f64, f64, f64, C.POINTER(vp)]
    L.glba_map_destroy.argtypes = [vp]
    L.glba_map_destroy.restype = None
    L.glba_map_size.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_int64)]
    L.glba_map_add_keyframes.argtypes = [vp, i32, vp, C.POINTER(i32)]
    L.glba_map_add_points.argtypes = [vp, i32, vp, C.POINTER(i32)]
    L.glba_map_add_observations.argtypes = [vp, i32, vp, vp, vp]
    L.glba_map_set_bad.argtypes = [vp, i32, vp, C.c_uint8]
    L.glba_map_write_keyframes.argtypes = [vp, i32, i32, vp]
    L.glba_map_read_keyframes.argtypes = [vp, i32, i32, vp]
    L.glba_map_write_points.argtypes = [vp, i32, i32, vp]
    L.glba_map_read_points.argtypes = [vp, i32, i32, vp, vp]
    L.glba_map_solve_window.argtypes = [vp, i32, i32, i32, i32, C.POINTER(_abi.Options), C.POINTER(_abi.Summary), C.POINTER(i32),
                                        C.POINTER(C.c_int64)]
    L.glba_map_cull_points.argtypes = [vp, i32, i32, i32, f64, C.POINTER(i32), C.POINTER(i32), vp, i32]
    L.glba_map_propagate.argtypes = [vp, vp, vp, i32, i32, vp, i32, vp, vp, vp]
    _LIB = L
    return L


def options(**kw):
    """glba_options with the reference's hard-coded values (slam_core.cpp:814, 842-847) as defaults."""
    return _abi.make_options(lib().glba_default_options, **kw)


def kernel_launch_count():
    return int(lib().glba_kernel_launch_count())


def nccl_unique_id():
    buf = C.create_string_buffer(_abi.GLBA_NCCL_ID_BYTES)
    st = lib().glba_nccl_unique_id(buf)
    if st:
        raise GlbaError(st, "glba_nccl_unique_id")
    return buf.raw


class Context:
    """One glba_ctx: a CUDA stream, workspaces and (world>1) an NCCL communicator.  Not thread-safe."""

    def __init__(self, device=0, rank=0, world=1, nccl_id=None, stream=None):
        self._h = C.c_void_p()
        self._id = C.create_string_buffer(nccl_id, _abi.GLBA_NCCL_ID_BYTES) if nccl_id is not None else None
        cfg = _abi.DeviceCfg(device, rank, world, C.cast(self._id, C.c_void_p) if self._id is not None else None, stream)
        st = lib().glba_create(C.byref(cfg), C.byref(self._h))
        if st:
            self._h = C.c_void_p()
            raise GlbaError(st, "glba_create")
        self.rank, self.world, self.device = rank, world, device
        self._maps = weakref.WeakSet()      # resident maps created on this context: they must be destroyed first

    def close(self):
        if self._h:
            for m in list(getattr(self, "_maps", ())):
                m.close()
            lib().glba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st, where):
        if st:
            raise GlbaError(st, where, lib().glba_last_error(self._h).decode(errors="replace"))

    # -- full BA -------------------------------------------------------------------------------
    def solve(self, prob, opt=None):
        """Bundle-adjust a HostProblem.  Returns (refined copy, summary dict); input is not modified."""
        opt = opt or options()
        work = prob.copy()
        summ = _abi.Summary()
        st = lib().glba_solve(self._h, C.byref(work.struct()), C.byref(opt), C.byref(summ))
        self._check(st, "glba_solve")
        return work, summ.as_dict()

    def linearize(self, prob, radius, opt=None, per_obs=True):
        opt = opt or options()
        out = _abi.LinearizationOut(prob.n_cam, prob.n_pt, prob.n_obs, per_obs)
        s = out.struct()
        st = lib().glba_linearize(self._h, C.byref(prob.struct()), C.byref(opt), float(radius), C.byref(s))
        self._check(st, "glba_linearize")
        out.take(s)
        return out

    # -- resident (device) problem -------------------------------------------------------------
    def load(self, prob_struct, opt=None):
        """prob_struct: _abi.Problem (host or device pointers).  Uploads and builds the device index."""
        opt = opt or options()
        self._check(lib().glba_load(self._h, C.byref(prob_struct), C.byref(opt)), "glba_load")

    def linearize_resident(self, radius, opt=None, want_cost=True):
        opt = opt or options()
        c = C.c_double()
        st = lib().glba_linearize_resident(self._h, C.byref(opt), float(radius), C.byref(c) if want_cost else None)
        self._check(st, "glba_linearize_resident")
        return c.value if want_cost else None

    def solve_resident(self, opt=None):
        opt = opt or options()
        summ = _abi.Summary()
        self._check(lib().glba_solve_resident(self._h, C.byref(opt), C.byref(summ)), "glba_solve_resident")
        return summ.as_dict()

    def time_kernels(self, radius=1e4, reps=5, opt=None):
        opt = opt or options()
        kt = _abi.KernelTimes()
        self._check(lib().glba_time_kernels(self._h, C.byref(opt), float(radius), int(reps), C.byref(kt)), "glba_time_kernels")
        return kt.as_dict()

    def reset_resident(self):
        self._check(lib().glba_reset_resident(self._h), "glba_reset_resident")

    def read_resident(self, n_cam, n_pt):
        cam, pt = np.zeros((n_cam, 6)), np.zeros((n_pt, 3))
        self._check(lib().glba_read_resident(self._h, cam.ctypes.data, pt.ctypes.data), "glba_read_resident")
        return cam, pt

    def synchronize(self):
        self._check(lib().glba_synchronize(self._h), "glba_synchronize")

    @property
    def stream(self):
        return lib().glba_stream(self._h)

    # -- pose-only BA --------------------------------------------------------------------------
    def pose_only(self, cam, X, uv, K, opt=None):
        """Refine one camera [angle-axis, centre] against fixed points.  Returns (cam, summary dict)."""
        opt = opt or options()
        cam = np.ascontiguousarray(cam, dtype=np.float64).copy()
        X = np.ascontiguousarray(X, dtype=np.float64)
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        if X.shape[0] != uv.shape[0]:
            raise ValueError("p3d / p2d size mismatch")
        summ = _abi.Summary()
        st = lib().glba_pose_only(self._h, cam.ctypes.data, X.shape[0], X.ctypes.data, uv.ctypes.data,
                                  *map(float, K), C.byref(opt), C.byref(summ))
        self._check(st, "glba_pose_only")
        return cam, summ.as_dict()

    def pose_only_batch(self, cams, offsets, X, uv, K, opt=None):
        opt = opt or options()
        cams = np.ascontiguousarray(cams, dtype=np.float64).copy()
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        X = np.ascontiguousarray(X, dtype=np.float64)
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        b = cams.shape[0]
        usable, iters, cost = np.zeros(b, np.uint8), np.zeros(b, np.int32), np.zeros(b)
        st = lib().glba_pose_only_batch(self._h, b, cams.ctypes.data, offsets.ctypes.data, X.ctypes.data, uv.ctypes.data,
                                        *map(float, K), C.byref(opt), usable.ctypes.data, iters.ctypes.data,
                                        cost.ctypes.data)
        self._check(st, "glba_pose_only_batch")
        return cams, usable.astype(bool), iters, cost

    # -- two-view triangulation + filter (slam_core.cpp:173-256) ---------------------------------
    def triangulate_filter(self, R1, t1, R2, t2, K, p0, p1, distance_threshold, reprojection_threshold):
        """[R|t] world-to-camera; p0/p1 (n,2) matched pixels.  Returns (X (n,3) = X/w for every match, keep (n,) bool)."""
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (R1, t1, R2, t2, p0, p1)]
        n = a[4].shape[0]
        X, keep = np.zeros((n, 3)), np.zeros(n, np.uint8)
        st = lib().glba_triangulate_filter(self._h, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data,
                                           *map(float, K), n, a[4].ctypes.data, a[5].ctypes.data, float(distance_threshold),
                                           float(reprojection_threshold), X.ctypes.data, keep.ctypes.data)
        self._check(st, "glba_triangulate_filter")
        return X, keep.astype(bool)

    # -- post-BA culling -----------------------------------------------------------------------
    def cull_points(self, prob, min_obs=3, max_mean_err=1.0):
        bad = np.zeros(prob.n_pt, dtype=np.uint8)
        err = np.zeros(prob.n_pt)
        st = lib().glba_cull_points(self._h, C.byref(prob.struct()), min_obs, float(max_mean_err), bad.ctypes.data,
                                    err.ctypes.data)
        self._check(st, "glba_cull_points")
        return bad, err


class DeviceMap:
    """glba_map: the persistent device-resident mirror of GL-SLAM's Map (keyframes, points, observation log).

    Grown where the reference grows its map (slam_core.cpp:287-426); `solve_window` is full_ba (slam_core.cpp:744-883)
    with the window selected and packed on the device."""

    def __init__(self, ctx, K):
        self._ctx = ctx                     # keeps the context alive: the map must go first
        self._h = C.c_void_p()
        st = lib().glba_map_create(ctx._h, *(float(k) for k in K), C.byref(self._h))
        if st:
            self._h = C.c_void_p()
            ctx._check(st, "glba_map_create")
        ctx._maps.add(self)

    def close(self):
        if self._h:
            lib().glba_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        a, b, c = C.c_int32(), C.c_int32(), C.c_int64()
        self._ctx._check(lib().glba_map_size(self._h, C.byref(a), C.byref(b), C.byref(c)), "glba_map_size")
        return a.value, b.value, c.value

    def add_keyframes(self, cam):
        cam = np.ascontiguousarray(cam, dtype=np.float64).reshape(-1, 6)
        first = C.c_int32()
        self._ctx._check(lib().glba_map_add_keyframes(self._h, len(cam), cam.ctypes.data, C.byref(first)), "glba_map_add_keyframes")
        return first.value

    def add_points(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        first = C.c_int32()
        self._ctx._check(lib().glba_map_add_points(self._h, len(xyz), xyz.ctypes.data, C.byref(first)), "glba_map_add_points")
        return first.value

    def add_observations(self, kf, pt, uv):
        kf = np.ascontiguousarray(kf, dtype=np.int32)
        pt = np.ascontiguousarray(pt, dtype=np.int32)
        uv = np.ascontiguousarray(uv, dtype=np.float64).reshape(-1, 2)
        if not (len(kf) == len(pt) == len(uv)):
            raise ValueError("observation arrays differ in length")
        self._ctx._check(lib().glba_map_add_observations(self._h, len(kf), kf.ctypes.data, pt.ctypes.data, uv.ctypes.data),
                         "glba_map_add_observations")

    def set_bad(self, pt_ids, value=True):
        ids = np.ascontiguousarray(pt_ids, dtype=np.int32)
        self._ctx._check(lib().glba_map_set_bad(self._h, len(ids), ids.ctypes.data, 1 if value else 0), "glba_map_set_bad")

    def write_keyframes(self, first, cam):
        cam = np.ascontiguousarray(cam, dtype=np.float64).reshape(-1, 6)
        self._ctx._check(lib().glba_map_write_keyframes(self._h, int(first), len(cam), cam.ctypes.data), "glba_map_write_keyframes")

    def read_keyframes(self, first=0, n=None):
        n = self.size()[0] - first if n is None else n
        cam = np.zeros((n, 6))
        self._ctx._check(lib().glba_map_read_keyframes(self._h, int(first), int(n), cam.ctypes.data), "glba_map_read_keyframes")
        return cam

    def write_points(self, first, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, 3)
        self._ctx._check(lib().glba_map_write_points(self._h, int(first), len(xyz), xyz.ctypes.data), "glba_map_write_points")

    def read_points(self, first=0, n=None):
        n = self.size()[1] - first if n is None else n
        xyz, bad = np.zeros((n, 3)), np.zeros(n, np.uint8)
        self._ctx._check(lib().glba_map_read_points(self._h, int(first), int(n), xyz.ctypes.data, bad.ctypes.data), "glba_map_read_points")
        return xyz, bad

    def propagate(self, R_before, t_before, kf_last, kf_ids, pt_ids):
        """post_ba_map_update_for_new_keyframes (slam_core.cpp:916-973) on the device; returns (dR, dt)."""
        Rb = np.ascontiguousarray(R_before, dtype=np.float64).reshape(3, 3)
        tb = np.ascontiguousarray(t_before, dtype=np.float64).reshape(3)
        kf = np.ascontiguousarray(kf_ids, dtype=np.int32)
        pt = np.ascontiguousarray(pt_ids, dtype=np.int32)
        dR, dt = np.zeros((3, 3)), np.zeros(3)
        st = lib().glba_map_propagate(self._h, Rb.ctypes.data, tb.ctypes.data, int(kf_last), len(kf), kf.ctypes.data if len(kf) else None,
                                      len(pt), pt.ctypes.data if len(pt) else None, dR.ctypes.data, dt.ctypes.data)
        self._ctx._check(st, "glba_map_propagate")
        return dR, dt

    def solve_window(self, first_kf, window, opt=None, n_fixed=2, min_obs=1):
        """Returns the summary dict, with the window's size under 'n_pt' / 'n_obs'."""
        opt = opt or options()
        summ = _abi.Summary()
        npt, nobs = C.c_int32(), C.c_int64()
        st = lib().glba_map_solve_window(self._h, int(first_kf), int(window), int(n_fixed), int(min_obs), C.byref(opt), C.byref(summ),
                                         C.byref(npt), C.byref(nobs))
        self._ctx._check(st, "glba_map_solve_window")
        d = summ.as_dict()
        d["n_pt"], d["n_obs"] = npt.value, nobs.value
        return d

    def cull_points(self, first_kf, last_kf, min_obs=3, max_mean_err=1.0):
        """post_ba_map_point_culling on the resident map.  Returns (n_candidates, culled point ids)."""
        ncand, ncull = C.c_int32(), C.c_int32()
        cap = self.size()[1]
        ids = np.zeros(max(cap, 1), np.int32)
        st = lib().glba_map_cull_points(self._h, int(first_kf), int(last_kf), int(min_obs), float(max_mean_err), C.byref(ncand),
                                        C.byref(ncull), ids.ctypes.data, cap)
        self._ctx._check(st, "glba_map_cull_points")
        return ncand.value, ids[:ncull.value].copy()
