"""GLBA_MODE_G2O, CPU side: the C++ oracle's restatement of the archived g2o BA against an independent numpy one, and
against the Ceres formulation where the two must agree (same objective, same minimiser)."""
import numpy as np
import pytest

from gl_slam_b200 import scene
from gl_slam_b200._abi import MODE_G2O
from oracle import py_oracle_g2o


def small(seed=11, **kw):
    args = dict(n_cam=6, n_pt=120, track_len=4, seed=seed, outlier_frac=0.05, rot_sigma=0.005, pos_sigma=0.03)
    args.update(kw)
    prob = scene.make_scene(**args)
    prob.cam_fixed[:] = 0
    prob.cam_fixed[0] = 1                      # the g2o path fixes camera 0 only (4image_pnp_ba.txt:355-357)
    return prob


def test_pose_convention_round_trip():
    prob = small()
    back = scene.to_camera_to_world(scene.to_world_to_camera(prob.cam))
    assert np.allclose(back, prob.cam, rtol=0, atol=1e-14)
    g = scene.as_g2o(prob)
    R = scene.rodrigues(g.cam[:, :3])
    p = np.einsum("nij,nj->ni", R[g.obs_cam], g.pt[g.obs_pt]) + g.cam[g.obs_cam, 3:]
    u, v, depth = scene.project(prob.cam, prob.pt, prob.obs_cam, prob.obs_pt, prob.K)
    fx, fy, cx, cy = prob.K
    assert np.allclose(fx * p[:, 0] / p[:, 2] + cx, u, rtol=1e-12) and np.allclose(fy * p[:, 1] / p[:, 2] + cy, v, rtol=1e-12)
    assert np.allclose(p[:, 2], depth, rtol=1e-12)


@pytest.mark.parametrize("loss", [0, 1, 2])
def test_cpp_oracle_matches_numpy_g2o(oracle, loss):
    g = scene.as_g2o(small())
    ref, s = oracle.solve(g, oracle.options(loss=loss, mode=MODE_G2O, max_iters=12))
    q = py_oracle_g2o.solve(g, loss_kind=loss, max_iters=12)
    assert (s["n_iters"], s["n_successful"]) == (q["n_iters"], q["n_successful"])
    assert list(s["accepted"]) == q["accepted"]
    assert np.allclose(s["cost"], q["cost"], rtol=1e-10, atol=0)
    assert np.allclose(1.0 / np.asarray(s["radius"]), q["lam"], rtol=1e-8)
    assert np.allclose(ref.cam, q["cam"], rtol=1e-7, atol=1e-9)
    assert np.allclose(ref.pt, q["pt"], rtol=1e-7, atol=1e-7)
    assert np.array_equal(ref.cam[0], g.cam[0])                      # the fixed vertex keeps its bits


def test_rejected_trials_and_termination(oracle):
    """A violently perturbed start makes the first trials fail: lambda grows by nu = 2, 4, 8 ...; with one trial per
    iteration allowed, the first failure terminates the optimisation (g2o's "Terminate")."""
    g = scene.as_g2o(small(seed=6, rot_sigma=0.15, pos_sigma=1.5, pt_sigma=4.0))
    ref, s = oracle.solve(g, oracle.options(loss=0, mode=MODE_G2O, max_iters=15))
    q = py_oracle_g2o.solve(g, loss_kind=0, max_iters=15)
    assert 0 in list(s["accepted"])[1:], "scene too tame: no rejected trial"
    assert list(s["accepted"]) == q["accepted"]
    assert np.allclose(s["cost"], q["cost"], rtol=1e-9)
    acc = np.asarray(s["accepted"])
    lam = 1.0 / np.asarray(s["radius"])
    k = int(np.nonzero(acc[1:] == 0)[0][0]) + 1
    assert np.isclose(lam[k] / lam[k - 1], 2.0)                      # first rejection: lambda *= nu (= 2)
    _, s1 = oracle.solve(g, oracle.options(loss=0, mode=MODE_G2O, max_iters=15, g2o_max_trials=1))
    assert s1["stop_reason"] == 8 and s1["n_iters"] == k and s1["n_successful"] == k - 1


def test_same_minimum_as_ceres_formulation(oracle):
    """No robust kernel, clean data, gauge fixed by two cameras: both formulations minimise the same sum of squares, so
    they must meet at the same poses and points (a check that needs no g2o: the Ceres-formulation oracle is pinned
    against OpenCV and complex-step derivatives in test_oracle.py)."""
    prob = scene.make_scene(n_cam=6, n_pt=150, track_len=4, seed=3, outlier_frac=0.0, rot_sigma=0.004, pos_sigma=0.03)
    a, sa = oracle.solve(prob, oracle.options(loss=0, max_iters=60, function_tol=1e-16, parameter_tol=1e-14))
    g = scene.as_g2o(prob)
    b, sb = oracle.solve(g, oracle.options(loss=0, mode=MODE_G2O, max_iters=40))
    assert abs(sa["final_cost"] - sb["final_cost"]) <= 1e-9 * sa["final_cost"]
    assert np.allclose(scene.to_camera_to_world(b.cam), a.cam, rtol=1e-6, atol=1e-7)
    assert np.allclose(b.pt, a.pt, rtol=1e-6, atol=1e-6)


def inv_depth2_information(prob):
    """docs/old_unorganized/4image_pnp_ba.txt:400-401: weight = 1/z^2 of the point (world z), 1/0.1^2 below 0.1."""
    z = prob.pt[:, 2]
    return np.where(z > 0.1, 1.0 / np.maximum(z, 0.1) ** 2, 1.0 / 0.01)


@pytest.mark.parametrize("loss,delta", [(1, 3.0), (0, 1.0)])
def test_information_weights_match_numpy_g2o(oracle, loss, delta):
    """The archived variant with edge information I/z^2 and RobustKernelHuber(delta = 3) (4image_pnp_ba.txt:400-406)."""
    g = scene.as_g2o(small())
    g.pt_info = inv_depth2_information(g)
    ref, s = oracle.solve(g, oracle.options(loss=loss, loss_scale=delta, mode=MODE_G2O, max_iters=10))
    q = py_oracle_g2o.solve(g, loss_kind=loss, loss_a=delta, max_iters=10)
    assert list(s["accepted"]) == q["accepted"]
    assert np.allclose(s["cost"], q["cost"], rtol=1e-10, atol=0)
    assert np.allclose(ref.cam, q["cam"], rtol=1e-7, atol=1e-9)
    # uniform information c is the same as scaling the residuals by sqrt(c): cost scales by c when there is no kernel
    g1 = scene.as_g2o(small())
    g1.pt_info = np.full(g1.n_pt, 4.0)
    g0 = scene.as_g2o(small())
    assert np.isclose(oracle.solve(g1, oracle.options(loss=0, mode=MODE_G2O, max_iters=0))[1]["initial_cost"],
                      4.0 * oracle.solve(g0, oracle.options(loss=0, mode=MODE_G2O, max_iters=0))[1]["initial_cost"], rtol=1e-14)
