import sys, time; sys.path.insert(0, '.')
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
ctx = g.Context(0)
for name, p in (("C2", scene.config("C2")), ("C1", scene.config("C1")), ("w7", scene.make_scene(7, 3000, 4, seed=5, outlier_frac=0.03, rot_sigma=0.004, pos_sigma=0.03))):
    ctx.solve(p)
    t0 = time.perf_counter(); reps = 10
    for _ in range(reps): _, s = ctx.solve(p)
    dt = (time.perf_counter() - t0) / reps
    print({k: round(s[k], 3) for k in s if k.startswith('t_')})
    print(name, 'solve ms %.3f' % (dt * 1e3), 'iters', s['n_iters'], 'us/iter %.1f' % (dt * 1e6 / max(1, s['n_iters'])), 'LM it/s %.0f' % (s['n_iters'] / dt))
