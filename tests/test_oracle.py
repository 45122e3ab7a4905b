"""CPU tests: the oracle against its golden vectors, an independent numpy restatement, closed forms."""
import numpy as np
import pytest

from gl_slam_b200 import scene
from gl_slam_b200._abi import HostProblem
from oracle import py_oracle

from helpers import check_state, check_trajectory, golden_names, load_golden, max_rel


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden(oracle, name):
    prob, z = load_golden(name)
    ref, s = oracle.solve(prob, oracle.options(loss=int(z["loss"])))
    check_trajectory(s, z)
    assert max_rel(s["radius"], z["radius"]) < 1e-9
    assert s["stop_reason"] == int(z["stop_reason"]) and s["termination"] == int(z["termination"])
    check_state(prob, ref.cam, ref.pt, z["cam_final"], z["pt_final"])


@pytest.mark.parametrize("name", ["tiny_cauchy", "tiny_huber", "tiny_none", "twoview"])
def test_oracle_matches_independent_numpy(oracle, name):
    """Different derivative method (complex step) and different linear algebra (dense lstsq, no Schur)."""
    prob, z = load_golden(name)
    loss = int(z["loss"])
    ref, s = oracle.solve(prob, oracle.options(loss=loss))
    q = py_oracle.solve(prob, loss_kind=loss)
    assert q["n_iters"] == s["n_iters"] and q["stop_reason"] == s["stop_reason"]
    assert max_rel(q["cost"], s["cost"]) < 1e-9
    assert q["accepted"] == s["accepted"]
    assert np.allclose(q["cam"], ref.cam, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("loss", [0, 1, 2])
def test_autodiff_jacobians_match_complex_step(oracle, loss):
    prob = scene.make_scene(5, 60, 3, seed=5, outlier_frac=0.2, rot_sigma=0.3, pos_sigma=0.1)
    prob.cam[0, :3] = 0.0            # Ceres' small-angle branch (theta^2 <= epsilon)
    prob.cam[1, :3] = [1e-9, -2e-9, 1e-9]
    L = oracle.linearize(prob, 1e4, oracle.options(loss=loss))
    Jc, Jp = py_oracle.jacobian_complex_step(prob.cam, prob.pt, prob.obs_cam, prob.obs_pt, prob.obs_u, prob.obs_v, prob.K)
    r = py_oracle.residuals(prob.cam, prob.pt, prob.obs_cam, prob.obs_pt, prob.obs_u, prob.obs_v, prob.K)
    rho, rho1 = py_oracle.loss(loss, 1.0, (r * r).sum(1))
    w = np.sqrt(rho1)
    assert np.abs(L.jac_cam - w[:, None, None] * Jc).max() <= 1e-9 * np.abs(Jc).max()
    assert np.abs(L.jac_pt - w[:, None, None] * Jp).max() <= 1e-9 * np.abs(Jp).max()
    assert np.abs(L.residuals - w[:, None] * r).max() <= 1e-9 * max(1.0, np.abs(r).max())
    assert abs(L.cost - 0.5 * rho.sum()) <= 1e-12 * abs(L.cost)


def test_rotation_convention_matches_cv2(oracle):
    """full_ba packs cv::Rodrigues(R_wc) (slam_core.cpp:769); the residual must see R_wc^T (X - c)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for _ in range(20):
        w = rng.normal(size=3) * rng.choice([1e-3, 0.1, 1.0, 2.5])
        R, _ = cv2.Rodrigues(w.reshape(3, 1))
        assert np.allclose(scene.rodrigues(w)[0], R, atol=1e-13)
        w2, _ = cv2.Rodrigues(R)
        assert np.allclose(scene.rotation_to_angle_axis(R)[0], w2.ravel(), atol=1e-9)
        c, X = rng.normal(size=3), rng.normal(size=3) + [0, 0, 10]
        p = R.T @ (X - c)
        u, v = 700 * p[0] / p[2] + 600, 710 * p[1] / p[2] + 200
        prob = HostProblem(np.r_[w, c][None], X[None], [0], [0], [u + 0.25], [v - 0.5], (700, 710, 600, 200))
        L = oracle.linearize(prob, 1e4, oracle.options(loss=0))
        assert np.allclose(L.residuals[0], [-0.25, 0.5], atol=1e-9)


def test_zero_noise_scene_is_a_fixed_point(oracle):
    prob, cam_gt, pt_gt = scene.make_scene(6, 100, 4, seed=3, pixel_sigma=0.0, rot_sigma=0.0, pos_sigma=0.0, pt_sigma=0.0,
                                           return_gt=True)
    ref, s = oracle.solve(prob)
    # residuals are O(1e-13) px (projection round-off): either the gradient test or the function test stops at once
    assert s["initial_cost"] < 1e-18 and s["n_iters"] <= 1 and s["termination"] == 0 and s["n_successful"] == 0
    assert np.array_equal(ref.cam, prob.cam) and np.array_equal(ref.pt, prob.pt)


def test_recovers_ground_truth_without_noise(oracle):
    prob, cam_gt, pt_gt = scene.make_scene(6, 150, 4, seed=4, pixel_sigma=0.0, rot_sigma=0.003, pos_sigma=0.02, pt_sigma=0.05,
                                           return_gt=True)
    ref, s = oracle.solve(prob, oracle.options(loss=0, function_tol=1e-14, max_iters=60))
    assert s["final_cost"] < 1e-10 * s["initial_cost"]
    # two fixed cameras pin the gauge: free cameras return to ground truth
    assert np.abs(ref.cam - cam_gt).max() < 1e-5


def test_edge_cases(oracle):
    prob = scene.make_scene(4, 30, 3, seed=9)
    # no observations at all
    empty = HostProblem(prob.cam, prob.pt, [], [], [], [], prob.K, prob.cam_fixed)
    ref, s = oracle.solve(empty)
    assert s["status"] == 0 and s["n_iters"] == 0 and s["final_cost"] == 0.0
    # an unobserved camera / point stays untouched and does not enter |x|
    cam = np.vstack([prob.cam, [[0.1, 0.2, 0.3, 50, 60, 70]]])
    pt = np.vstack([prob.pt, [[7, 8, 9]]])
    big = HostProblem(cam, pt, prob.obs_cam, prob.obs_pt, prob.obs_u, prob.obs_v, prob.K, np.r_[prob.cam_fixed, 0])
    r1, s1 = oracle.solve(prob)
    r2, s2 = oracle.solve(big)
    # (OpenMP chunking changes with n_pt, so sums differ in the last bits)
    assert max_rel(s1["cost"], s2["cost"]) < 1e-11 and np.array_equal(r2.cam[-1], cam[-1]) and np.array_equal(r2.pt[-1], pt[-1])
    assert np.allclose(r1.cam, r2.cam[:-1], rtol=1e-9, atol=1e-12)
    # all points fixed: structure stays, only free cameras move (pose-only flavour)
    fixed = HostProblem(prob.cam, prob.pt, prob.obs_cam, prob.obs_pt, prob.obs_u, prob.obs_v, prob.K, prob.cam_fixed,
                        np.ones(prob.n_pt, np.uint8))
    r3, s3 = oracle.solve(fixed)
    assert np.array_equal(r3.pt, prob.pt) and s3["final_cost"] <= s3["initial_cost"]
    # out-of-range index is rejected, inputs untouched
    bad = HostProblem(prob.cam, prob.pt, prob.obs_cam.copy(), prob.obs_pt, prob.obs_u, prob.obs_v, prob.K)
    bad.obs_cam[0] = 99
    before = bad.cam.copy()
    _, s4 = oracle.solve(bad)
    assert s4["status"] == -1 and np.array_equal(bad.cam, before)


def test_pose_only_golden(oracle):
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "pose_only.npz"))
    cam, s = oracle.pose_only(z["cam0"], z["X"], z["uv"], tuple(z["K"]))
    assert s["n_iters"] == int(z["n_iters"]) and max_rel(s["cost"], z["cost"]) < 1e-9
    assert np.allclose(cam, z["cam_final"], rtol=1e-6, atol=1e-9)
    # and it is the same thing as full BA with every point held constant
    n = z["X"].shape[0]
    prob = HostProblem(z["cam0"][None], z["X"], np.zeros(n, int), np.arange(n), z["uv"][:, 0], z["uv"][:, 1], tuple(z["K"]),
                       None, np.ones(n, np.uint8))
    ref, s2 = oracle.solve(prob)
    assert max_rel(s2["cost"], s["cost"]) < 1e-12 and np.allclose(ref.cam[0], cam, rtol=1e-10, atol=1e-13)


def test_cull_points_matches_numpy(oracle):
    """post_ba_map_point_culling (slam_core.cpp:993-1035): mean pixel error / cheirality / observation count."""
    prob = scene.make_scene(8, 200, lambda rng, n: 2 + rng.poisson(1.5, size=n), seed=17, outlier_frac=0.1)
    prob.pt[5] = prob.cam[prob.obs_cam[prob.obs_pt == 5][0], 3:6] - [0, 0, 5.0]   # behind its camera
    bad, err = oracle.cull_points(prob, 3, 1.0)
    u, v, depth = scene.project(prob.cam, prob.pt, prob.obs_cam, prob.obs_pt, prob.K)
    e = np.hypot(u - prob.obs_u, v - prob.obs_v)
    cnt = np.bincount(prob.obs_pt, minlength=prob.n_pt)
    mean = np.bincount(prob.obs_pt, weights=e, minlength=prob.n_pt) / np.maximum(cnt, 1)
    behind = np.bincount(prob.obs_pt, weights=(depth <= 0), minlength=prob.n_pt) > 0
    want = behind | (cnt < 3) | (mean > 1.0)
    assert np.array_equal(bad.astype(bool), want) and bad[5] == 1
    assert np.allclose(err[~behind], mean[~behind], rtol=1e-12)


def test_jacobians_match_opencv_chain_rule(oracle):
    """Third-party pin of row a1 (residual + 2x6 / 2x3 blocks): OpenCV's own analytic Jacobians.

    The reference residual is  K * R(-w) (X - c)  (slam_core.cpp:705-726).  With rvec = -w, tvec = -R(rvec) c this is
    cv2.projectPoints(X, rvec, tvec, K); OpenCV returns d(proj)/d(rvec), d(proj)/d(tvec), and cv2.Rodrigues returns dR/d(rvec).
    Chain rule to the reference's parameters (w additive, c = camera centre):
        d/dw = -( J_r + J_t * d(tvec)/d(rvec) ),  d(tvec)/d(rvec_i) = -(dR/d rvec_i) c,   d/dc = -J_t R,   d/dX = J_t R."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    K = scene.KITTI_K
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]])
    n = 40
    cam = np.c_[rng.normal(0, 0.4, (n, 3)), rng.normal(0, 2.0, (n, 3))]
    Rwc = scene.rodrigues(cam[:, :3])
    pc = np.c_[rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(4, 30, n)]
    X = np.einsum("nij,nj->ni", Rwc, pc) + cam[:, 3:]
    u, v, _ = scene.project(cam, X, np.arange(n), np.arange(n), K)
    prob = HostProblem(cam, X, np.arange(n), np.arange(n), u + 0.3, v - 0.2, K)
    L = oracle.linearize(prob, 1e4, oracle.options(loss=0))
    for i in range(n):
        rvec = -cam[i, :3]
        R, dR = cv2.Rodrigues(rvec.reshape(3, 1))          # dR: 3 x 9, row k = d vec(R) / d rvec_k
        tvec = -R @ cam[i, 3:]
        img, jac = cv2.projectPoints(X[i].reshape(1, 1, 3), rvec.reshape(3, 1), tvec.reshape(3, 1), Km, None)
        assert np.allclose(img.ravel(), [u[i], v[i]], atol=1e-9)
        Jr, Jt = jac[:, 0:3], jac[:, 3:6]
        dt_dr = np.stack([-(dR[k].reshape(3, 3) @ cam[i, 3:]) for k in range(3)], axis=1)     # 3x3: column k = d tvec / d rvec_k
        Jw = -(Jr + Jt @ dt_dr)
        Jc = -Jt @ R
        JX = Jt @ R
        got_c, got_p = L.jac_cam[i], L.jac_pt[i]
        scale = np.abs(got_c).max()
        assert np.abs(got_c[:, :3] - Jw).max() < 1e-9 * scale, (i, np.abs(got_c[:, :3] - Jw).max())
        assert np.abs(got_c[:, 3:] - Jc).max() < 1e-9 * scale
        assert np.abs(got_p - JX).max() < 1e-9 * scale


def test_pose_only_optimum_matches_opencv_pnp_refine(oracle):
    """Without a robust loss pose_only_ba minimises the plain reprojection error: the optimum must coincide with
    OpenCV's LM refinement of the same objective (cv2.solvePnPRefineLM), whatever path either solver takes."""
    cv2 = pytest.importorskip("cv2")
    cam0, X, uv, _ = scene.pose_only_scene(400, seed=11, outlier_frac=0.0)
    K = scene.KITTI_K
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]])
    got, s = oracle.pose_only(cam0, X, uv, K, oracle.options(loss=0, function_tol=1e-15, parameter_tol=1e-14, max_iters=100))
    assert s["termination"] == 0
    R0 = scene.rodrigues(-cam0[:3])[0]
    rvec, tvec = (-cam0[:3]).reshape(3, 1).copy(), (-R0 @ cam0[3:]).reshape(3, 1).copy()
    rvec, tvec = cv2.solvePnPRefineLM(X.reshape(-1, 1, 3), uv.reshape(-1, 1, 2), Km, None, rvec, tvec,
                                      criteria=(cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 200, 1e-15))
    R = cv2.Rodrigues(rvec)[0]
    c_cv = -R.T @ tvec.ravel()
    w_cv = -rvec.ravel()
    assert np.allclose(got[:3], w_cv, atol=2e-7) and np.allclose(got[3:], c_cv, atol=2e-6), (got, w_cv, c_cv)
