"""Stage times of the pair-structure build (glba.cu::ensure_explicit) on C4: first load of a context (buffers allocated) and
steady state.  Run with GLBA_CG_PROF=1; the stages are printed on stderr."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gl_slam_b200 as g
from gl_slam_b200 import scene
prob = scene.config("C4")
lopt = g.options(max_iters=2, function_tol=0.0, parameter_tol=0.0, gradient_tol=0.0, cg_rel_tol=1e-2, cg_max_iters=40)
with g.Context(device=0) as c:
    for k in range(3):
        sys.stderr.write(f"--- load {k}\n"); sys.stderr.flush()
        c.load(prob.struct(), lopt)
        c.solve_resident(lopt)
