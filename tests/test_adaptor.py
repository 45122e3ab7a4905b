"""The C++ host adaptor (include/glba_slam.hpp): slam_types mirrors + full_ba / pose_only_ba over the C ABI."""
import os
import subprocess

import numpy as np
import pytest

import gl_slam_b200 as g
from gl_slam_b200 import scene
from gl_slam_b200._abi import HostProblem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("adaptor") / "adaptor_test"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "adaptor_test.cpp"),
                           "-o", str(out), g.LIB_PATH, "-Wl,-rpath," + os.path.dirname(g.LIB_PATH), "-lpthread"])
    return str(out)


def write_scene(path, prob, window, run_window, first, bad=None):
    R = scene.rodrigues(prob.cam[:, :3])
    with open(path, "w") as f:
        f.write(f"{prob.n_cam} {prob.n_pt} {prob.n_obs} {prob.K[0]!r} {prob.K[1]!r} {prob.K[2]!r} {prob.K[3]!r} {window} {run_window} {first}\n")
        for i in range(prob.n_cam):
            f.write(" ".join(repr(float(x)) for x in list(R[i].ravel()) + list(prob.cam[i, 3:])) + "\n")
        for j in range(prob.n_pt):
            f.write(" ".join(repr(float(x)) for x in prob.pt[j]) + f" {int(bad[j]) if bad is not None else 0}\n")
        for k in range(prob.n_obs):
            f.write(f"{int(prob.obs_cam[k])} {int(prob.obs_pt[k])} {float(prob.obs_u[k])!r} {float(prob.obs_v[k])!r}\n")


def window_scene():
    """12 keyframes with ids 5..16; the BA window is the last 10 (ids 7..16, run_window = 16)."""
    return scene.make_scene(12, 300, 4, seed=77, rot_sigma=0.004, pos_sigma=0.03, outlier_frac=0.03, n_fixed=0)


def expected_window(prob, first_cam, window, bad):
    """What full_ba must hand to the solver: window cameras, their non-bad points, in-window observations."""
    in_win = (prob.obs_cam >= first_cam) & (prob.obs_cam < first_cam + window)
    seen = np.zeros(prob.n_pt, bool)
    seen[prob.obs_pt[in_win]] = True
    keep_pt = seen & ~bad
    new_id = -np.ones(prob.n_pt, int)
    new_id[keep_pt] = np.arange(keep_pt.sum())
    sel = in_win & keep_pt[prob.obs_pt]
    fixed = np.zeros(window, np.uint8)
    fixed[:2] = 1
    sub = HostProblem(prob.cam[first_cam:first_cam + window], prob.pt[keep_pt], prob.obs_cam[sel] - first_cam, new_id[prob.obs_pt[sel]],
                      prob.obs_u[sel], prob.obs_v[sel], prob.K, fixed)
    return sub, keep_pt


def test_pack_window_restates_reference_packing(exe, tmp_path):
    prob = window_scene()
    bad = np.zeros(prob.n_pt, bool)
    bad[[3, 50]] = True
    path = tmp_path / "scene.txt"
    write_scene(path, prob, window=10, run_window=16, first=5, bad=bad)
    out = subprocess.check_output([exe, "pack", str(path)]).decode().split("\n")
    n_cam, n_pt, n_obs, first = map(int, out[0].split())
    sub, keep = expected_window(prob, 2, 10, bad)
    assert (n_cam, n_pt, n_obs, first) == (10, sub.n_pt, sub.n_obs, 7)
    cams = np.array(out[1].split(), float).reshape(10, 6)
    assert np.allclose(cams, sub.cam, atol=1e-12)                 # cv::Rodrigues(R) -> angle-axis, t
    ids = np.array(out[2].split(), int)
    assert np.array_equal(ids, 1000 + np.flatnonzero(keep))       # is_bad points skipped (slam_core.cpp:791)
    obs = np.array(out[3].split(), float).reshape(-1, 4)
    assert np.array_equal(obs[:, 0].astype(int), sub.obs_cam) and np.array_equal(obs[:, 1].astype(int), sub.obs_pt)
    assert np.array_equal(obs[:, 2], sub.obs_u) and np.array_equal(obs[:, 3], sub.obs_v)
    assert out[4].split() == ["1", "1"] + ["0"] * 8               # cameras 0 and 1 constant (slam_core.cpp:831-833)
    assert float(out[5]) < 1e-14                                  # Rodrigues round trip
    err, consumed = out[6].split()                                # post_ba_map_update_for_new_keyframes (slam_core.cpp:916-973)
    assert float(err) < 1e-7 and consumed == "1"


@pytest.mark.gpu
def test_full_ba_adaptor_matches_c_abi(exe, tmp_path, ctx, oracle):
    prob = window_scene()
    bad = np.zeros(prob.n_pt, bool)
    bad[[3, 50]] = True
    path = tmp_path / "scene.txt"
    write_scene(path, prob, window=10, run_window=16, first=5, bad=bad)
    out = subprocess.check_output([exe, "solve", str(path)]).decode().split("\n")
    head = out[0].split()
    assert head[0] == "ok"
    sub, keep = expected_window(prob, 2, 10, bad)
    ref, so = oracle.solve(sub)
    assert int(head[1]) == so["n_iters"] and abs(float(head[3]) - so["final_cost"]) <= 1e-9 * so["final_cost"]
    got = np.array([l.split() for l in out[1:11]], float)
    assert np.allclose(got[:, :9].reshape(10, 3, 3), scene.rodrigues(ref.cam[:, :3]), atol=1e-7)
    assert np.allclose(got[:, 9:], ref.cam[:, 3:], rtol=1e-6, atol=1e-8)
    pts = np.array(out[11].split(), float).reshape(-1, 3)
    assert np.array_equal(pts[~keep], prob.pt[~keep])             # points outside the problem are untouched
    from helpers import check_state
    check_state(sub, ref.cam, pts[keep], ref.cam, ref.pt)
    # keyframes before the window are untouched: not printed, but culling ran over keyframes [run_window-window, run_window-4]
    culled = int(out[12])
    flags = np.array(out[13].split(), int)
    assert culled == flags.sum() - 2 and flags[3] == 1 and flags[50] == 1
    # post-BA propagation ran inside full_ba's critical section (PostBa): keyframe 17 follows keyframe 16, the late point moved
    perr, consumed, _ = out[14].split()
    assert float(perr) < 1e-9 and consumed == "1"


@pytest.mark.gpu
def test_full_ba_resident_matches_full_ba(exe, tmp_path):
    """ResidentMap + full_ba_resident (the map mirrored into HBM keyframe by keyframe, window packed on the device) must
    leave the host Map exactly where full_ba (host packing) leaves it."""
    prob = window_scene()
    bad = np.zeros(prob.n_pt, bool)
    bad[[3, 50]] = True
    path = tmp_path / "scene.txt"
    write_scene(path, prob, window=10, run_window=16, first=5, bad=bad)
    a = subprocess.check_output([exe, "solve", str(path)]).decode().split("\n")
    b = subprocess.check_output([exe, "resident", str(path)]).decode().split("\n")
    assert a[0].split()[0] == b[0].split()[0] == "ok"
    assert a[0].split()[1] == b[0].split()[1]                                         # iterations
    for la, lb in zip(a[:12], b[:12]):
        assert np.allclose(np.array(la.split()[1 if la.startswith("ok") else 0:], float),
                           np.array(lb.split()[1 if lb.startswith("ok") else 0:], float), rtol=1e-9, atol=1e-12)
    assert a[12:14] == b[12:14]                                                        # culling result
    perr, consumed, mirror = b[14].split()                                             # propagation: host map and device mirror agree
    assert float(perr) < 1e-9 and consumed == "1" and float(mirror) < 1e-9


@pytest.mark.gpu
def test_g2o_bundle_adjustment_adaptor(exe, tmp_path, oracle):
    """glslam::bundleAdjustment (the archived g2o entry point, Old/mult_img_recoverpose_single_ba:251-326) against the
    oracle's g2o restatement on the same world-to-camera poses."""
    from gl_slam_b200._abi import MODE_G2O
    prob = window_scene()
    path = tmp_path / "scene.txt"
    write_scene(path, prob, window=10, run_window=16, first=5)
    out = subprocess.check_output([exe, "g2o", str(path)]).decode().split("\n")
    head = out[0].split()
    assert head[0] == "ok"
    gp = scene.as_g2o(prob)
    gp.cam_fixed = np.zeros(prob.n_cam, np.uint8)
    gp.cam_fixed[0] = 1
    gp.K = (prob.K[0], prob.K[0], prob.K[2], prob.K[3])
    ref, so = oracle.solve(gp, oracle.options(mode=MODE_G2O, loss=0, max_iters=12))
    assert (int(head[1]), int(head[2])) == (so["n_iters"], so["n_successful"])
    assert abs(float(head[4]) - so["final_cost"]) <= 1e-9 * so["final_cost"]
    got = np.array([l.split() for l in out[1:1 + prob.n_cam]], float)
    assert np.allclose(got[:, :9].reshape(-1, 3, 3), scene.rodrigues(ref.cam[:, :3]), atol=1e-7)
    assert np.allclose(got[:, 9:], ref.cam[:, 3:], rtol=1e-6, atol=1e-7)
    pts = np.array(out[1 + prob.n_cam].split(), float).reshape(-1, 3)
    assert np.allclose(pts, ref.pt, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_pose_only_adaptor(exe, tmp_path, oracle):
    cam0, X, uv, _ = scene.pose_only_scene(400, seed=5)
    R = scene.rodrigues(cam0[:3])[0]
    path = tmp_path / "pose.txt"
    with open(path, "w") as f:
        f.write(f"{X.shape[0]} " + " ".join(repr(float(k)) for k in scene.KITTI_K) + "\n")
        f.write(" ".join(repr(float(x)) for x in list(R.ravel()) + list(cam0[3:])) + "\n")
        for i in range(X.shape[0]):
            f.write(" ".join(repr(float(x)) for x in list(X[i]) + list(uv[i])) + "\n")
    out = subprocess.check_output([exe, "pose", str(path)]).decode().split("\n")
    want, s = oracle.pose_only(cam0, X, uv, scene.KITTI_K)
    head = out[0].split()
    assert head[0] == "ok" and int(head[1]) == s["n_iters"]
    got = np.array(out[1].split(), float)
    assert np.allclose(got[:9].reshape(3, 3), scene.rodrigues(want[:3])[0], atol=1e-8)
    assert np.allclose(got[9:], want[3:], rtol=1e-6, atol=1e-9)
