"""Kernel times of the assembled reduced camera matrix on a bench workload (one GPU): structure build, assembly, one PCG
iteration of the cooperative kernel, and a short LM solve both ways (GLBA_EXPLICIT=0 = matrix-free product)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gl_slam_b200 as g  # noqa: E402
from gl_slam_b200 import scene  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C4"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    prob = scene.config(name, scale=scale)
    out = {"workload": name, "scale": scale, "n_cam": prob.n_cam, "n_obs": prob.n_obs}
    lopt = g.options(max_iters=6, function_tol=0.0, parameter_tol=0.0, gradient_tol=0.0, cg_rel_tol=1e-2, cg_max_iters=40)
    for mode in ("explicit", "matrix_free"):
        if mode == "matrix_free":
            os.environ["GLBA_EXPLICIT"] = "0"
        with g.Context(device=0) as c:
            c.load(prob.struct(), lopt)
            kt = c.time_kernels(1e4, reps=10, opt=lopt)
            c.reset_resident()
            c.solve_resident(lopt)
            c.reset_resident()
            t0 = time.perf_counter()
            s = c.solve_resident(lopt)
            wall = time.perf_counter() - t0
        os.environ.pop("GLBA_EXPLICIT", None)
        out[mode] = {"lm_iters_per_s": s["n_iters"] / wall, "wall_ms": wall * 1e3, "cg_iters": s["cg_iters"][1:s["n_iters"] + 1],
                     "final_cost": s["final_cost"],
                     "device_ms": {k: s[k] for k in ("t_linearize_ms", "t_schur_ms", "t_solve_ms", "t_update_ms")},
                     "kernels": {k: v for k, v in kt.items() if v}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
