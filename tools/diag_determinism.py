"""Run the bench's inexact-Newton LM on C4 several times in one process: per-iteration costs must repeat bit for bit."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gl_slam_b200 as g
from gl_slam_b200 import scene
prob = scene.config(sys.argv[1] if len(sys.argv) > 1 else "C4", scale=float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
cgmax = int(sys.argv[3]) if len(sys.argv) > 3 else 40
opt = g.options(max_iters=6, function_tol=0.0, parameter_tol=0.0, gradient_tol=0.0, cg_rel_tol=1e-2, cg_max_iters=cgmax)
with g.Context(device=0) as c:
    runs = []
    for rep in range(4):
        got, s = c.solve(prob, opt)
        runs.append(s)
        print(rep, "cost", ["%.17g" % x for x in s["cost_candidate"][:7]])
        print(rep, "rel ", ["%.6e" % x for x in s["relative_decrease"][:7]], s["accepted"][:7], s["cg_iters"][:7])
    print("bitwise reproducible:", all(r["cost_candidate"] == runs[0]["cost_candidate"] for r in runs))
