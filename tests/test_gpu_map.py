"""Persistent device-resident map: window selection / packing on the device must give the same local BA
as packing the same window on the host (slam_core.cpp:750-819) and handing it to glba_solve — and as the oracle."""
import numpy as np
import pytest

import gl_slam_b200 as g
from gl_slam_b200 import scene
from gl_slam_b200._abi import HostProblem
from helpers import check_trajectory, well_conditioned_points

pytestmark = pytest.mark.gpu


def slam_scene(n_cam=40, n_pt=6000, seed=5):
    return scene.make_scene(n_cam, n_pt, lambda rng, n: 3 + rng.poisson(2.0, size=n), seed=seed, rot_sigma=0.003, pos_sigma=0.03,
                            outlier_frac=0.03)


def replay(ctx, prob, upto=None, rng=None):
    """Feed the map the way update_map_and_keyframe_data does: one keyframe at a time, new points when first observed,
    that keyframe's observations in arbitrary order."""
    m = g.DeviceMap(ctx, prob.K)
    upto = prob.n_cam if upto is None else upto
    first_seen = np.full(prob.n_pt, prob.n_cam, np.int64)
    np.minimum.at(first_seen, prob.obs_pt, prob.obs_cam)
    assert np.all(np.diff(first_seen[first_seen < prob.n_cam]) >= 0), "scene must number points in creation order"
    n_added = 0
    for k in range(upto):
        assert m.add_keyframes(prob.cam[k]) == k
        new = np.nonzero(first_seen == k)[0]
        if len(new):
            assert new[0] == n_added and m.add_points(prob.pt[new]) == n_added
            n_added += len(new)
        sel = np.nonzero(prob.obs_cam == k)[0]
        if rng is not None:
            sel = rng.permutation(sel)
        m.add_observations(prob.obs_cam[sel], prob.obs_pt[sel], np.stack([prob.obs_u[sel], prob.obs_v[sel]], axis=1))
    return m, n_added


def host_window(prob, cam, pt, first, window, n_known_pt, bad=None, min_obs=1):
    in_win = (prob.obs_cam >= first) & (prob.obs_cam < first + window) & (prob.obs_pt < n_known_pt)
    if bad is not None:
        in_win &= ~bad[prob.obs_pt]
    keep = np.bincount(prob.obs_pt[in_win], minlength=prob.n_pt) >= min_obs
    sel = in_win & keep[prob.obs_pt]
    new_id = np.cumsum(keep) - 1
    fixed = np.zeros(window, np.uint8)
    fixed[:2] = 1
    sub = HostProblem(cam[first:first + window], pt[keep], prob.obs_cam[sel] - first, new_id[prob.obs_pt[sel]], prob.obs_u[sel],
                      prob.obs_v[sel], prob.K, fixed)
    return sub, keep


def test_window_schedule_matches_host_packing(ctx):
    prob = slam_scene()
    m, n_added = replay(ctx, prob, rng=np.random.default_rng(1))
    assert m.size() == (prob.n_cam, n_added, int(np.sum(prob.obs_pt < n_added)))
    cam, pt = prob.cam.copy(), prob.pt.copy()
    ctx2 = g.Context()
    for first in range(0, prob.n_cam - 10 + 1, 7):
        sub, keep = host_window(prob, cam, pt, first, 10, n_added)
        ref, sh = ctx2.solve(sub)
        sm = m.solve_window(first, 10)
        assert (sm["n_pt"], sm["n_obs"]) == (sub.n_pt, sub.n_obs)
        check_trajectory(sm, sh, rtol=1e-9)
        cam[first:first + 10] = ref.cam
        pt[keep] = ref.pt
        assert np.allclose(m.read_keyframes(), cam, rtol=1e-9, atol=1e-12), first
        assert np.allclose(m.read_points()[0], pt[:n_added], rtol=1e-9, atol=1e-12), first
    ctx2.close()
    m.close()


@pytest.mark.parametrize("min_obs", [1, 2])
def test_window_matches_oracle(ctx, oracle, min_obs):
    prob = slam_scene(n_cam=14, n_pt=1500, seed=9)
    m, n_added = replay(ctx, prob)
    bad = np.zeros(prob.n_pt, bool)
    bad[::11] = True
    m.set_bad(np.nonzero(bad[:n_added])[0])
    sub, keep = host_window(prob, prob.cam, prob.pt, 3, 10, n_added, bad=bad, min_obs=min_obs)
    ref, so = oracle.solve(sub, oracle.options())
    sm = m.solve_window(3, 10, min_obs=min_obs)
    assert (sm["n_pt"], sm["n_obs"]) == (sub.n_pt, sub.n_obs)
    cam = m.read_keyframes()
    pts, bad_dev = m.read_points()
    assert np.array_equal(bad_dev.astype(bool), bad[:n_added])
    if min_obs == 1:
        # single-observation points leave the late iterations (radius > 1e9) chaotic at the 1e-16 level (see
        # test_gpu_parity.test_sliding_window_schedule): gate the well-defined part against the oracle; the strict
        # check of this path is test_window_schedule_matches_host_packing
        assert abs(sm["n_iters"] - so["n_iters"]) <= 3
        assert np.allclose(sm["cost"][:6], so["cost"][:6], rtol=1e-8)
        assert abs(sm["final_cost"] - so["final_cost"]) <= 0.1 * so["final_cost"]
        m.close()
        return
    check_trajectory(sm, so, rtol=1e-9)
    assert np.allclose(cam[3:13], ref.cam, rtol=1e-6, atol=1e-9)
    assert np.array_equal(cam[:3], prob.cam[:3]) and np.array_equal(cam[13:], prob.cam[13:])      # outside the window: untouched
    ok = well_conditioned_points(sub, ref.pt)      # outlier-hit two-view points are not determined to 1e-6 by the data
    assert ok.mean() > 0.9
    assert np.allclose(pts[keep[:n_added]][ok], ref.pt[ok], rtol=1e-6, atol=1e-7)
    assert np.array_equal(pts[~keep[:n_added]], prob.pt[:n_added][~keep[:n_added]])                # bad / unobserved: untouched
    m.close()


def test_growth_host_edits_and_errors(ctx):
    m = g.DeviceMap(ctx, scene.KITTI_K)
    rng = np.random.default_rng(3)
    cams, pts, obs = [], [], []
    for step in range(60):                         # many small appends: every array outgrows its capacity several times
        c = rng.normal(size=(rng.integers(1, 4), 6))
        p = rng.normal(size=(rng.integers(1, 300), 3))
        assert m.add_keyframes(c) == sum(len(x) for x in cams)
        assert m.add_points(p) == sum(len(x) for x in pts)
        cams.append(c)
        pts.append(p)
        nk, npnt, _ = m.size()
        o = np.stack([rng.integers(0, nk, 200), rng.integers(0, npnt, 200)], axis=1)
        uv = rng.normal(size=(200, 2))
        m.add_observations(o[:, 0], o[:, 1], uv)
        obs.append(o)
    cams, pts = np.concatenate(cams), np.concatenate(pts)
    assert m.size() == (len(cams), len(pts), 60 * 200)
    assert np.array_equal(m.read_keyframes(), cams)
    assert np.array_equal(m.read_points()[0], pts)
    m.write_keyframes(5, cams[5:8] + 1.0)
    m.write_points(100, pts[100:140] * 2.0)
    assert np.array_equal(m.read_keyframes(5, 3), cams[5:8] + 1.0)
    assert np.array_equal(m.read_points(100, 40)[0], pts[100:140] * 2.0)
    with pytest.raises(g.GlbaError):
        m.add_observations([len(cams)], [0], [[0.0, 0.0]])        # keyframe not in the map
    with pytest.raises(g.GlbaError):
        m.add_observations([0], [len(pts)], [[0.0, 0.0]])         # point not in the map
    with pytest.raises(g.GlbaError):
        m.solve_window(len(cams) - 5, 10)                         # window runs past the newest keyframe
    with pytest.raises(g.GlbaError):
        m.read_points(len(pts) - 1, 2)
    assert m.size() == (len(cams), len(pts), 60 * 200)            # failed calls left the map alone
    m.close()


def test_empty_window(ctx):
    """A window nobody observes (or whose points are all bad) is a no-op, not an error."""
    m = g.DeviceMap(ctx, scene.KITTI_K)
    m.add_keyframes(np.zeros((4, 6)))
    s = m.solve_window(0, 4)
    assert (s["n_pt"], s["n_obs"], s["n_iters"]) == (0, 0, 0)
    m.add_points([[0.0, 0.0, 5.0]])
    m.add_observations([2, 3], [0, 0], [[600.0, 180.0], [601.0, 181.0]])
    m.set_bad([0])
    s = m.solve_window(0, 4)
    assert (s["n_pt"], s["n_obs"]) == (0, 0)
    m.close()


def test_map_cull_points(ctx):
    """post_ba_map_point_culling on the resident map against the same rule in numpy (slam_core.cpp:977-1038)."""
    prob = scene.make_scene(30, 3000, lambda rng, n: 3 + rng.poisson(2.0, size=n), seed=12, rot_sigma=0.0, pos_sigma=0.0, pt_sigma=0.0,
                            outlier_frac=0.02)                       # a converged map: 0.5 px noise, a few outliers
    rng = np.random.default_rng(4)
    wrong = rng.choice(prob.n_pt, 200, replace=False)
    prob.pt[wrong[:150]] += rng.normal(0.0, 0.5, (150, 3))            # large reprojection error
    prob.pt[wrong[150:]] -= np.array([0.0, 0.0, 80.0])                # behind their cameras
    m, n_added = replay(ctx, prob)
    pre_bad = np.zeros(n_added, bool)
    pre_bad[::17] = True
    m.set_bad(np.nonzero(pre_bad)[0])
    first_kf, last_kf, min_obs, max_err = 5, 20, 4, 1.0
    ncand, culled = m.cull_points(first_kf, last_kf, min_obs=min_obs, max_mean_err=max_err)
    # reference rule
    sel = prob.obs_pt < n_added
    oc, op = prob.obs_cam[sel], prob.obs_pt[sel]
    u, v, z = scene.project(prob.cam, prob.pt, oc, op, prob.K)
    err = np.hypot(u - prob.obs_u[sel], v - prob.obs_v[sel])
    first = np.full(n_added, 10 ** 9)
    np.minimum.at(first, op, oc)
    cnt = np.bincount(op, minlength=n_added)
    behind = np.zeros(n_added, bool)
    np.logical_or.at(behind, op, z <= 0.0)
    mean = np.bincount(op, weights=err, minlength=n_added) / np.maximum(cnt, 1)
    cand = ~pre_bad & (first >= first_kf) & (first <= last_kf)
    want = cand & (behind | (cnt < min_obs) | (mean > max_err))
    near = cand & ~behind & (np.abs(mean - max_err) < 1e-9)           # ties at the threshold would be rounding-dependent
    assert not near.any()
    assert ncand == int(cand.sum())
    assert sorted(culled.tolist()) == np.nonzero(want)[0].tolist()
    assert want.sum() > 50 and (cand & ~want).sum() > 50
    _, bad_dev = m.read_points()
    assert np.array_equal(bad_dev.astype(bool), pre_bad | want)
    # a second pass finds nothing new among the same candidates; an empty range is a no-op
    ncand2, culled2 = m.cull_points(first_kf, last_kf, min_obs=min_obs, max_mean_err=max_err)
    assert ncand2 == int((cand & ~want).sum()) and len(culled2) == 0
    assert m.cull_points(25, 20)[0] == 0
    m.close()


def test_context_closed_before_its_map():
    """Closing a context destroys the maps created on it first; closing such a map afterwards is a no-op, not a crash."""
    c = g.Context()
    m = g.DeviceMap(c, scene.KITTI_K)
    m.add_keyframes(np.zeros((2, 6)))
    c.close()
    m.close()
    assert not m._h


def _project_to_so3_ref(A):
    """ProjectToSO3 (slam_core.cpp:885-897) with numpy's SVD (singular values descending, as cv::SVD)."""
    U, _, Vt = np.linalg.svd(A)
    R = U @ Vt
    if np.linalg.det(R) < 0:
        U = U.copy()
        U[:, 2] *= -1.0
        R = U @ Vt
    return R


@pytest.mark.parametrize("case", ["drifted_rotation", "reflection"])
def test_propagate_matches_numpy_restatement(ctx, case):
    """glba_map_propagate (device) against a numpy restatement of ComputeDeltaPose_SO3 / post_ba_map_update_for_new_keyframes
    (slam_core.cpp:885-973): a `before` rotation with numerical drift, and one with det < 0 (the U.col(2) flip)."""
    rng = np.random.default_rng(5)
    prob = scene.make_scene(8, 200, 3, seed=12)
    dm = g.DeviceMap(ctx, prob.K)
    dm.add_keyframes(prob.cam)
    dm.add_points(prob.pt)
    kf_last = 5
    Ra = scene.rodrigues(prob.cam[kf_last, :3])[0]
    ta = prob.cam[kf_last, 3:]
    Rb = scene.rodrigues(prob.cam[kf_last, :3] + np.array([0.01, -0.02, 0.005]))[0] + 1e-7 * rng.normal(size=(3, 3))
    if case == "reflection":
        Rb = Rb @ np.diag([1.05, 1.0, -0.95])             # det < 0 with distinct singular values: the flip is well defined
    tb = ta + np.array([0.2, -0.1, 0.05])
    kf_ids, pt_ids = np.array([6, 7]), np.array([3, 50, 199])
    dR, dt = dm.propagate(Rb, tb, kf_last, kf_ids, pt_ids)
    dR_ref = _project_to_so3_ref(_project_to_so3_ref(Ra) @ _project_to_so3_ref(Rb).T)
    dt_ref = ta - dR_ref @ tb
    assert np.abs(dR - dR_ref).max() < 1e-12 and np.abs(dt - dt_ref).max() < 1e-11
    cams = dm.read_keyframes()
    pts, _ = dm.read_points()
    for i in range(8):
        if i in kf_ids:
            R_new = dR_ref @ scene.rodrigues(prob.cam[i, :3])[0]
            assert np.abs(scene.rodrigues(cams[i, :3])[0] - R_new).max() < 1e-11
            assert np.abs(cams[i, 3:] - (dR_ref @ prob.cam[i, 3:] + dt_ref)).max() < 1e-11
        else:
            assert np.array_equal(cams[i], prob.cam[i])
    moved = np.zeros(prob.n_pt, bool)
    moved[pt_ids] = True
    assert np.abs(pts[moved] - (prob.pt[moved] @ dR_ref.T + dt_ref)).max() < 1e-11
    assert np.array_equal(pts[~moved], prob.pt[~moved])
    dm.close()
