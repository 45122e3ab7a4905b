"""Small end-to-end cases for compute-sanitizer (memcheck): every kernel family runs at least once."""
import sys; sys.path.insert(0, '.')
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
from gl_slam_b200._abi import HostProblem
ctx = g.Context(0)
p = scene.make_scene(8, 700, lambda rng, n: 2 + rng.poisson(2.0, size=n), seed=1, outlier_frac=0.05, rot_sigma=0.004, pos_sigma=0.03)
for ls in (g.LINSOLVE_DENSE, g.LINSOLVE_PCG):
    r, s = ctx.solve(p, g.options(linsolve=ls, max_iters=6)); print('solve', ls, s['n_iters'], s['final_cost'])
p2 = scene.make_scene(40, 3000, lambda rng, n: 2 + rng.poisson(3.0, size=n), seed=2, rot_sigma=0.003, pos_sigma=0.03)
perm = np.random.default_rng(0).permutation(p2.n_obs)
shuf = HostProblem(p2.cam, p2.pt, p2.obs_cam[perm], p2.obs_pt[perm], p2.obs_u[perm], p2.obs_v[perm], p2.K, p2.cam_fixed)
r, s = ctx.solve(shuf, g.options(max_iters=4, cg_max_iters=30)); print('pcg unsorted', s['n_iters'], s['cg_iters'])
L = ctx.linearize(p, 1e4); print('linearize', L.cost)
bad, err = ctx.cull_points(p); print('cull', int(bad.sum()))
cam0, X, uv, _ = scene.pose_only_scene(333, seed=3)
c, ps = ctx.pose_only(cam0, X, uv, scene.KITTI_K); print('pose', ps['n_iters'])
# long track (> TILE_OBS/2 observations of one point) -> thread-per-point fallback kernels
n_cam = 600
cam = np.zeros((n_cam, 6)); cam[:, 3] = np.linspace(0, 6, n_cam)
pt = np.array([[3.0, 0.2, 12.0], [2.0, -0.3, 9.0]])
oc = np.r_[np.arange(n_cam), np.arange(0, n_cam, 2)]; op = np.r_[np.zeros(n_cam, int), np.ones(n_cam // 2, int)]
u, v, _ = scene.project(cam, pt, oc, op, scene.KITTI_K)
lp = HostProblem(cam + 1e-4, pt + 0.01, oc, op, u, v, scene.KITTI_K, (np.arange(n_cam) < 2).astype(np.uint8))
r, s = ctx.solve(lp, g.options(max_iters=3, cg_max_iters=20)); print('long track fallback', s['n_iters'], s['final_cost'])
print('done')
