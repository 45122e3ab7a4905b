import sys; sys.path.insert(0, '.')
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
from oracle import oracle
ctx = g.Context(0)
prob = scene.make_scene(60, 6000, lambda rng, n: 2 + rng.poisson(3.0, size=n), seed=41, rot_sigma=0.003, pos_sigma=0.03)
ref, so = oracle.solve(prob)
for tol, mi in ((1e-13, 0), (1e-15, 0), (1e-16, 4000), (1e-10, 0)):
    got, s = ctx.solve(prob, g.options(cg_rel_tol=tol, cg_max_iters=mi))
    print('tol', tol, 'iters', s['n_iters'], so['n_iters'], ['%.0e' % (abs(a - b) / abs(b)) for a, b in zip(s['cost'], so['cost'])], s['cg_iters'][1:], 'solve ms %.1f' % s['t_solve_ms'])
