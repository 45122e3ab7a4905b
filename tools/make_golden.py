"""Generate tests/golden/*.npz: seeded scenes + the oracle's answers (cost trajectory, radii, final state).

The reference ships no golden vectors, so these pin the oracle against regressions and
against the independent numpy restatement (oracle/py_oracle.py), which is run here and must agree
before a fixture is written.  Run from the repo root:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gl_slam_b200 import scene  # noqa: E402
from oracle import oracle, py_oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (scene kwargs, loss)
    "tiny_cauchy": (dict(n_cam=6, n_pt=120, track_len=4, seed=11, outlier_frac=0.05, rot_sigma=0.005, pos_sigma=0.03), 2),
    "tiny_huber": (dict(n_cam=6, n_pt=120, track_len=4, seed=12, outlier_frac=0.05, rot_sigma=0.005, pos_sigma=0.03), 1),
    "tiny_none": (dict(n_cam=6, n_pt=120, track_len=4, seed=13, outlier_frac=0.0, rot_sigma=0.005, pos_sigma=0.03), 0),
    "window10": (dict(n_cam=10, n_pt=400, track_len=4, seed=2, outlier_frac=0.05, rot_sigma=0.005, pos_sigma=0.03), 2),
    "twoview": (dict(n_cam=2, n_pt=200, track_len=2, seed=1, step=1.0, rot_sigma=0.0, pos_sigma=0.0, pt_sigma=0.3, n_fixed=2), 2),
    "ragged": (dict(n_cam=12, n_pt=300, track_len=lambda rng, n: 2 + rng.poisson(2.0, size=n), seed=21, rot_sigma=0.003,
                    pos_sigma=0.02), 2),
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (kw, loss) in CASES.items():
        prob = scene.make_scene(**kw)
        ref, s = oracle.solve(prob, oracle.options(loss=loss))
        q = py_oracle.solve(prob, loss_kind=loss)
        assert q["n_iters"] == s["n_iters"], (name, q["n_iters"], s["n_iters"])
        rel = max(abs(a - b) / abs(b) for a, b in zip(q["cost"], s["cost"]))
        assert rel < 1e-9, (name, rel)
        assert np.allclose(q["cam"], ref.cam, rtol=1e-6, atol=1e-9), name
        lin = oracle.linearize(prob, 1e4, oracle.options(loss=loss))
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), cam=prob.cam, pt=prob.pt, obs_cam=prob.obs_cam, obs_pt=prob.obs_pt,
            obs_u=prob.obs_u, obs_v=prob.obs_v, cam_fixed=prob.cam_fixed, K=np.array(prob.K), loss=loss,
            n_iters=s["n_iters"], stop_reason=s["stop_reason"], termination=s["termination"], cost=np.array(s["cost"]),
            cost_candidate=np.array(s["cost_candidate"]), radius=np.array(s["radius"]), accepted=np.array(s["accepted"]),
            cam_final=ref.cam, pt_final=ref.pt, lin_cost=lin.cost, lin_grad_cam=lin.grad_cam, lin_schur_rhs=lin.schur_rhs,
            lin_schur_diag=lin.schur_diag)
        print(f"{name}: {prob.n_cam} cams {prob.n_pt} pts {prob.n_obs} obs, {s['n_iters']} iterations, "
              f"cost {s['initial_cost']:.6e} -> {s['final_cost']:.6e}, numpy cross-check rel {rel:.1e}")
    cam0, X, uv, gt = scene.pose_only_scene(300, seed=7)
    cam, s = oracle.pose_only(cam0, X, uv, scene.KITTI_K)
    np.savez_compressed(os.path.join(OUT, "pose_only.npz"), cam0=cam0, X=X, uv=uv, K=np.array(scene.KITTI_K), cam_final=cam,
                        n_iters=s["n_iters"], cost=np.array(s["cost"]), radius=np.array(s["radius"]),
                        termination=s["termination"])
    print(f"pose_only: {s['n_iters']} iterations, cost {s['initial_cost']:.6e} -> {s['final_cost']:.6e}")


if __name__ == "__main__":
    main()
