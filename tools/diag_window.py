import sys; sys.path.insert(0, '.')
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
from gl_slam_b200._abi import HostProblem
from oracle import oracle
ctx = g.Context(0)
prob = scene.make_scene(66, 6600, lambda rng, n: 2 + rng.poisson(3.0, size=n), seed=3, rot_sigma=0.003, pos_sigma=0.03, n_fixed=0)
def window(prob, first, window, min_obs=1):
    in_win = (prob.obs_cam >= first) & (prob.obs_cam < first + window)
    cnt = np.bincount(prob.obs_pt[in_win], minlength=prob.n_pt)
    keep = cnt >= min_obs
    sel = in_win & keep[prob.obs_pt]
    new_id = -np.ones(prob.n_pt, int); new_id[keep] = np.arange(keep.sum())
    fixed = np.zeros(window, np.uint8); fixed[:2] = 1
    return HostProblem(prob.cam[first:first+window], prob.pt[keep], prob.obs_cam[sel]-first, new_id[prob.obs_pt[sel]], prob.obs_u[sel], prob.obs_v[sel], prob.K, fixed)
for mo in (1, 2):
    sub = window(prob, 0, 10, mo)
    ro, so = oracle.solve(sub)
    for ls in (g.LINSOLVE_DENSE, g.LINSOLVE_PCG):
        rg, sg = ctx.solve(sub, g.options(linsolve=ls))
        n = min(len(sg['cost']), len(so['cost']))
        print('min_obs', mo, 'ls', ls, 'iters', sg['n_iters'], so['n_iters'], ['%.0e' % (abs(a-b)/abs(b)) for a, b in zip(sg['cost'][:n], so['cost'][:n])])
        print('   radius rel', ['%.0e' % (abs(a-b)/abs(b)) for a, b in zip(sg['radius'][:n], so['radius'][:n])][:12], 'acc', sg['accepted'][:n] == so['accepted'][:n])
