"""Where the end-to-end (host buffers in, camera blocks out) time of one C4 linearisation goes."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gl_slam_b200 as g  # noqa: E402
from gl_slam_b200 import _abi, scene  # noqa: E402

prob = scene.config("C4", float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
dev = torch.device("cuda", 0)
keys = ("cam", "pt", "obs_cam", "obs_pt", "obs_u", "obs_v", "cam_fixed")
host = {k: torch.from_numpy(np.ascontiguousarray(getattr(prob, k))).pin_memory() for k in keys}
devb = {k: torch.empty_like(host[k], device=dev) for k in keys}


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def h2d():
    for k in keys:
        devb[k].copy_(host[k], non_blocking=True)


nbytes = sum(host[k].numel() * host[k].element_size() for k in keys)
t_h2d = timeit(h2d)
print(f"H2D of the problem ({nbytes / 1e6:.1f} MB, pinned): {t_h2d:.3f} ms = {nbytes / t_h2d / 1e6:.1f} GB/s")

ctx = g.Context()
opt = g.options()


def problem(src, memspace):
    ps = _abi.Problem()
    ps.n_cam, ps.n_pt, ps.n_obs = prob.n_cam, prob.n_pt, prob.n_obs
    ps.cam, ps.pt = src["cam"].data_ptr(), src["pt"].data_ptr()
    ps.obs_cam, ps.obs_pt, ps.obs_u, ps.obs_v = (src[k].data_ptr() for k in ("obs_cam", "obs_pt", "obs_u", "obs_v"))
    ps.cam_fixed, ps.pt_fixed = src["cam_fixed"].data_ptr(), None
    ps.fx, ps.fy, ps.cx, ps.cy = prob.K
    ps.memspace = memspace
    return ps


pd, ph = problem(devb, _abi.MEM_DEVICE), problem(host, _abi.MEM_HOST)
print(f"glba_load, inputs already on the device (index build only): {timeit(lambda: ctx.load(pd, opt)):.3f} ms")
print(f"glba_load from pinned host buffers: {timeit(lambda: ctx.load(ph, opt)):.3f} ms")
print(f"linearise + Schur, resident: {timeit(lambda: ctx.linearize_resident(1e4, opt, want_cost=False), 20):.3f} ms")
out = _abi.LinearizationOut(prob.n_cam, prob.n_pt, prob.n_obs, per_obs=False)
out.grad_pt = out.hess_pt = None
ls = out.struct()
print(f"glba_linearize(host problem) end to end: {timeit(lambda: g.lib().glba_linearize(ctx._h, C.byref(ph), C.byref(opt), 1e4, C.byref(ls))):.3f} ms")
