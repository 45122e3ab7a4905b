import sys; sys.path.insert(0, '.')
import gl_slam_b200 as g
from gl_slam_b200 import scene
ctx = g.Context(0)
p = scene.config("C2")
ctx.solve(p); _, s = ctx.solve(p); print(s['n_iters'])
