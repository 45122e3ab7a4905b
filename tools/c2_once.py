import sys; sys.path.insert(0, '.')
import gl_slam_b200 as g
from gl_slam_b200 import scene
ctx = g.Context(0)
p = scene.config("C2")
for _ in range(3):
    _, s = ctx.solve(p)
print(s["n_iters"], s["final_cost"])
