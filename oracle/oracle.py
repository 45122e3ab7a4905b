"""ctypes loader for oracle/libglba_oracle.so — TEST INFRASTRUCTURE ONLY.

May be imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never from gl_slam_b200 (the product).  Parity is unpinned at the Ceres boundary: see the
header of glba_oracle.cpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from gl_slam_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libglba_oracle.so")
    src = os.path.join(_HERE, "glba_oracle.cpp")
    hdr = os.path.join(_HERE, "..", "include", "glba.h")
    stale = (not os.path.exists(so)) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libglba_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libglba_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.glbao_default_options.argtypes = [C.POINTER(_abi.Options)]
        L.glbao_default_options.restype = None
        L.glbao_num_threads.restype = C.c_int
        L.glbao_cost.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.POINTER(C.c_double)]
        L.glbao_linearize.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.c_double,
                                      C.POINTER(_abi.Linearization)]
        L.glbao_step.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.c_double, C.POINTER(C.c_double),
                                 C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.glbao_solve.argtypes = [C.POINTER(_abi.Problem), C.POINTER(_abi.Options), C.POINTER(_abi.Summary)]
        L.glbao_pose_only.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.POINTER(_abi.Options), C.POINTER(_abi.Summary)]
        L.glbao_cull_points.argtypes = [C.POINTER(_abi.Problem), C.c_int32, C.c_double, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def options(**kw):
    return _abi.make_options(lib().glbao_default_options, **kw)


def num_threads():
    return lib().glbao_num_threads()


def set_num_threads(n):
    """OpenMP threads of the oracle (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().glbao_set_num_threads(int(n))


def cost(prob, opt=None):
    opt = opt or options()
    c = C.c_double()
    st = lib().glbao_cost(C.byref(prob.struct()), C.byref(opt), C.byref(c))
    if st:
        raise RuntimeError(f"glbao_cost -> {st}")
    return c.value


def linearize(prob, radius, opt=None, per_obs=True):
    opt = opt or options()
    out = _abi.LinearizationOut(prob.n_cam, prob.n_pt, prob.n_obs, per_obs)
    s = out.struct()
    st = lib().glbao_linearize(C.byref(prob.struct()), C.byref(opt), float(radius), C.byref(s))
    if st:
        raise RuntimeError(f"glbao_linearize -> {st}")
    out.take(s)
    return out


def step(prob, radius, opt=None):
    """One threaded linearise + Schur pass (what bench.py times on the CPU).  Returns (cost, t_eval_ms, t_schur_ms)."""
    opt = opt or options()
    c, te, ts = C.c_double(), C.c_double(), C.c_double()
    st = lib().glbao_step(C.byref(prob.struct()), C.byref(opt), float(radius), C.byref(c), C.byref(te), C.byref(ts))
    if st:
        raise RuntimeError(f"glbao_step -> {st}")
    return c.value, te.value, ts.value


def solve(prob, opt=None):
    """Returns (refined HostProblem copy, summary dict).  The input problem is not modified."""
    opt = opt or options()
    work = prob.copy()
    summ = _abi.Summary()
    st = lib().glbao_solve(C.byref(work.struct()), C.byref(opt), C.byref(summ))
    d = summ.as_dict()
    d["status"] = st
    return work, d


def pose_only(cam, X, uv, K, opt=None):
    opt = opt or options()
    cam = np.ascontiguousarray(cam, dtype=np.float64).copy()
    X = np.ascontiguousarray(X, dtype=np.float64)
    uv = np.ascontiguousarray(uv, dtype=np.float64)
    summ = _abi.Summary()
    st = lib().glbao_pose_only(cam.ctypes.data, X.shape[0], X.ctypes.data, uv.ctypes.data, *map(float, K),
                               C.byref(opt), C.byref(summ))
    d = summ.as_dict()
    d["status"] = st
    return cam, d


def cull_points(prob, min_obs=3, max_mean_err=1.0):
    bad = np.zeros(prob.n_pt, dtype=np.uint8)
    err = np.zeros(prob.n_pt)
    st = lib().glbao_cull_points(C.byref(prob.struct()), min_obs, float(max_mean_err), bad.ctypes.data,
                                 err.ctypes.data)
    if st:
        raise RuntimeError(f"glbao_cull_points -> {st}")
    return bad, err
