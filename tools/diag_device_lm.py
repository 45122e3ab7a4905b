import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
prob = scene.config("C2", rot_sigma=0.05, pos_sigma=0.3, pt_sigma=0.5)
with g.Context(device=0) as c:
    dev, sd = c.solve(prob, g.options(loss=1))
os.environ["GLBA_HOST_LM"] = "1"
with g.Context(device=0) as c:
    host, sh = c.solve(prob, g.options(loss=1))
n = sd["n_iters"] + 2
for k in ("accepted", "cost", "radius", "gradient_max_norm", "relative_decrease"):
    print(k); print(" dev ", list(sd[k][:n])); print(" host", list(sh[k][:n]))
print(sd["n_iters"], sh["n_iters"], sd["termination"], sh["termination"], sd["stop_reason"], sh["stop_reason"])
