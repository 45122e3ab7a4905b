"""GLBA_MODE_G2O on the GPU against the oracle's restatement of the archived g2o BA (SURVEY §8a "secondary formulation",
§8 f4): same trial sequence, per-trial cost within 1e-9, lambda within 1e-7, final poses / points within 1e-6."""
import numpy as np
import pytest

import gl_slam_b200 as g
from gl_slam_b200 import scene
from gl_slam_b200._abi import MODE_G2O
from helpers import well_conditioned_points

pytestmark = pytest.mark.gpu


def g2o_scene(n_cam, n_pt, seed, n_fixed=1, **kw):
    args = dict(track_len=4, outlier_frac=0.05, rot_sigma=0.005, pos_sigma=0.03)
    args.update(kw)
    prob = scene.make_scene(n_cam=n_cam, n_pt=n_pt, seed=seed, **args)
    prob.cam_fixed[:] = 0
    prob.cam_fixed[:n_fixed] = 1
    return scene.as_g2o(prob)


def compare(ctx, oracle, prob, rtol_cost=1e-9, **opt):
    ref, so = oracle.solve(prob, oracle.options(mode=MODE_G2O, **opt))
    got, s = ctx.solve(prob, g.options(mode=MODE_G2O, **opt))
    assert (s["n_iters"], s["n_successful"], s["stop_reason"], s["termination"]) == \
           (so["n_iters"], so["n_successful"], so["stop_reason"], so["termination"])
    assert list(s["accepted"]) == list(so["accepted"])
    assert np.allclose(s["cost"], so["cost"], rtol=rtol_cost, atol=0), np.max(np.abs(np.array(s["cost"]) / np.array(so["cost"]) - 1))
    assert np.allclose(s["radius"], so["radius"], rtol=1e-7)
    assert np.allclose(got.cam, ref.cam, rtol=1e-6, atol=1e-8)
    ok = well_conditioned_points(scene_ctw(prob), ref.pt)
    assert np.allclose(got.pt[ok], ref.pt[ok], rtol=1e-6, atol=1e-7)
    fixed = prob.cam_fixed.astype(bool)
    assert np.array_equal(got.cam[fixed], prob.cam[fixed])
    return s, so


def scene_ctw(prob):
    q = prob.copy()
    q.cam = scene.to_camera_to_world(prob.cam)
    return q


@pytest.mark.parametrize("loss", [0, 1, 2])
def test_window_dense_path(ctx, oracle, loss):
    """10 keyframes, camera 0 fixed (docs/old_unorganized/4image_pnp_ba.txt:355-357): the explicit reduced system."""
    compare(ctx, oracle, g2o_scene(10, 400, seed=2), loss=loss, max_iters=10)


def test_pcg_path(ctx, oracle):
    """40 cameras: implicit Schur complement + block-Jacobi PCG (parity tolerance of the PCG, 1e-13)."""
    compare(ctx, oracle, g2o_scene(40, 3000, seed=8, track_len=lambda rng, n: 3 + rng.poisson(2.0, size=n)), loss=1, max_iters=8)


def test_rejected_trials(ctx, oracle):
    """A start far from the optimum: trials are rejected (lambda x 2, 4, ...) and retried inside one g2o iteration."""
    prob = g2o_scene(6, 120, seed=6, rot_sigma=0.15, pos_sigma=1.5, pt_sigma=4.0)
    s, so = compare(ctx, oracle, prob, rtol_cost=1e-8, loss=0, max_iters=8)
    assert 0 in list(s["accepted"])[1:]
    got, s1 = ctx.solve(prob, g.options(mode=MODE_G2O, loss=0, max_iters=15, g2o_max_trials=1))
    ref, o1 = oracle.solve(prob, oracle.options(mode=MODE_G2O, loss=0, max_iters=15, g2o_max_trials=1))
    assert s1["stop_reason"] == o1["stop_reason"] == 8 and s1["n_iters"] == o1["n_iters"]


def test_same_minimum_as_ceres_mode(ctx):
    """Both formulations on the GPU, no robust kernel: one objective, one minimiser."""
    prob = scene.make_scene(n_cam=8, n_pt=300, track_len=4, seed=3, outlier_frac=0.0, rot_sigma=0.004, pos_sigma=0.03)
    a, sa = ctx.solve(prob, g.options(loss=0, max_iters=60, function_tol=1e-16, parameter_tol=1e-14))
    b, sb = ctx.solve(scene.as_g2o(prob), g.options(loss=0, mode=MODE_G2O, max_iters=40))
    assert abs(sa["final_cost"] - sb["final_cost"]) <= 1e-9 * sa["final_cost"]
    assert np.allclose(scene.to_camera_to_world(b.cam), a.cam, rtol=1e-6, atol=1e-7)
    assert np.allclose(b.pt, a.pt, rtol=1e-6, atol=1e-6)


def test_resident_and_mode_mismatch(ctx):
    prob = g2o_scene(10, 400, seed=2)
    got, s = ctx.solve(prob, g.options(mode=MODE_G2O, max_iters=5))
    ctx.load(prob.struct(), g.options(mode=MODE_G2O))
    s2 = ctx.solve_resident(g.options(mode=MODE_G2O, max_iters=5))
    assert s2["cost"] == s["cost"]
    with pytest.raises(g.GlbaError):
        ctx.solve_resident(g.options(max_iters=5))            # loaded in the g2o convention, solved as Ceres
    with pytest.raises(g.GlbaError):
        ctx.linearize(prob, 1e4, g.options(mode=MODE_G2O))


@pytest.mark.parametrize("mode", [MODE_G2O, 0])
def test_information_weights(ctx, oracle, mode):
    """glba_problem::pt_info: edge information I/z^2 with Huber(3) as in docs/old_unorganized/4image_pnp_ba.txt:400-406
    (g2o mode), and the same weighting through the Ceres formulation."""
    prob = g2o_scene(10, 400, seed=2) if mode == MODE_G2O else scene.make_scene(n_cam=10, n_pt=400, track_len=4, seed=2, outlier_frac=0.05,
                                                                                rot_sigma=0.005, pos_sigma=0.03)
    z = prob.pt[:, 2]
    prob.pt_info = np.where(z > 0.1, 1.0 / np.maximum(z, 0.1) ** 2, 100.0)
    o = dict(loss=1, loss_scale=3.0, max_iters=10, mode=mode)
    ref, so = oracle.solve(prob, oracle.options(**o))
    got, s = ctx.solve(prob, g.options(**o))
    assert s["n_iters"] == so["n_iters"] and list(s["accepted"]) == list(so["accepted"])
    assert np.allclose(s["cost"], so["cost"], rtol=1e-9, atol=0)
    assert np.allclose(got.cam, ref.cam, rtol=1e-6, atol=1e-8)
    plain = prob.copy()
    plain.pt_info = None
    _, s0 = ctx.solve(plain, g.options(**o))
    assert abs(s0["initial_cost"] - s["initial_cost"]) > 0.1 * s["initial_cost"]       # the weights do something
    ones = prob.copy()
    ones.pt_info = np.ones(prob.n_pt)
    _, s1 = ctx.solve(ones, g.options(**o))
    assert s1["cost"] == s0["cost"]                                                     # unit information is bit-for-bit the default
