import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
from gl_slam_b200._abi import HostProblem
from oracle import oracle
from test_gpu_parity import _window
ctx = g.Context(0)
prob = scene.make_scene(66, 6600, lambda rng, n: 4 + rng.poisson(3.0, size=n), seed=3, rot_sigma=0.002, pos_sigma=0.02, n_fixed=0, depth=(4.0, 20.0), step=1.0, min_parallax_deg=3.0)
for mo in (2, 1):
    cam, pt = prob.cam.copy(), prob.pt.copy()
    for first in range(0, 57, 7):
        cur = HostProblem(cam, pt, prob.obs_cam, prob.obs_pt, prob.obs_u, prob.obs_v, prob.K)
        sub, keep = _window(cur, first, 10, mo)
        ref, so = oracle.solve(sub); got, s = ctx.solve(sub)
        n = min(len(s['cost']), len(so['cost']))
        rel = [abs(a-b)/abs(b) for a, b in zip(s['cost'][:n], so['cost'][:n])]
        # conditioning proxy: min parallax among points (in-window)
        print('min_obs', mo, 'first', first, 'iters', s['n_iters'], so['n_iters'], 'max rel %.1e at it %d' % (max(rel), int(np.argmax(rel))), 'cam diff %.1e' % np.abs(got.cam-ref.cam).max(),
              'radius@max %.1e' % so['radius'][int(np.argmax(rel))])
        cam[first:first+10] = ref.cam; pt[keep] = ref.pt
