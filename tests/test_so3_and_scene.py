"""CPU tests of two pieces of round-2 host logic: the SO(3) projection shared by the adaptor and the device propagation
kernel (include/glba_so3.hpp, GL-SLAM slam_core.cpp:885-912) against numpy's SVD, and the street-grid scene (SURVEY §8d's C4)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from gl_slam_b200 import scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so3(tmp_path_factory):
    d = tmp_path_factory.mktemp("so3")
    src = d / "so3.cpp"
    src.write_text('#include "glba_so3.hpp"\n'
                   'extern "C" void proj(const double* A, double* R) { glba_so3::project_to_so3(A, R); }\n'
                   'extern "C" void delta(const double* Rb, const double* tb, const double* Ra, const double* ta, double* dR, double* dt) '
                   '{ glba_so3::compute_delta_pose_so3(Rb, tb, Ra, ta, dR, dt); }\n')
    so = d / "libso3.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(so)])
    return C.CDLL(str(so))


def _ref_project(A):
    """ProjectToSO3 as the reference writes it (slam_core.cpp:885-897) with numpy's SVD (singular values descending, as cv::SVD)."""
    U, _, Vt = np.linalg.svd(A)
    R = U @ Vt
    if np.linalg.det(R) < 0:
        U = U.copy()
        U[:, 2] *= -1.0
        R = U @ Vt
    return R


def test_project_to_so3_matches_svd(so3):
    rng = np.random.default_rng(0)
    for t in range(600):
        if t % 3 == 0:                    # a rotation with numerical drift: what the function is for
            A = np.linalg.qr(rng.normal(size=(3, 3)))[0] + 1e-6 * rng.normal(size=(3, 3))
            tol = 1e-13
        elif t % 3 == 1:                  # det < 0 with distinct singular values: the U.col(2) flip (:890-895)
            A = np.linalg.qr(rng.normal(size=(3, 3)))[0] @ np.diag([1.1, 1.0, -0.9]) @ np.linalg.qr(rng.normal(size=(3, 3)))[0]
            tol = 1e-12
        else:                             # anything
            A = rng.normal(size=(3, 3))
            tol = 1e-8
        A = np.ascontiguousarray(A)
        R = np.zeros((3, 3))
        so3.proj(A.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p))
        assert np.abs(R - _ref_project(A)).max() < tol, (t, np.linalg.svd(A)[1])
        assert abs(np.linalg.det(R) - 1.0) < 1e-9 and np.abs(R @ R.T - np.eye(3)).max() < 1e-9


def test_compute_delta_pose(so3):
    rng = np.random.default_rng(1)
    Rb = scene.rodrigues(rng.normal(0, 0.3, 3))[0] + 1e-8 * rng.normal(size=(3, 3))
    Ra = scene.rodrigues(rng.normal(0, 0.3, 3))[0]
    tb, ta = rng.normal(size=3), rng.normal(size=3)
    dR, dt = np.zeros((3, 3)), np.zeros(3)
    so3.delta(*(np.ascontiguousarray(x).ctypes.data_as(C.c_void_p) for x in (Rb, tb, Ra, ta)), dR.ctypes.data_as(C.c_void_p), dt.ctypes.data_as(C.c_void_p))
    dR_ref = _ref_project(_ref_project(Ra) @ _ref_project(Rb).T)
    assert np.abs(dR - dR_ref).max() < 1e-13 and np.abs(dt - (ta - dR_ref @ tb)).max() < 1e-12


def test_street_grid_scene_is_what_the_survey_describes():
    """60-wide streets driven in serpentine order: every observation inside the image at a sane depth, every keyframe observed,
    banded covisibility plus tracks that come back hundreds of keyframe ids later, point ids in creation order."""
    prob, cam_gt, pt_gt = scene.make_street_grid(n_rows=6, n_cols=60, n_pt=20000, track_len=lambda rng, n: 2 + rng.poisson(3.0, size=n), seed=4,
                                                 return_gt=True, pixel_sigma=0.0)
    assert prob.n_cam == 360 and prob.cam_fixed[:2].all() and not prob.cam_fixed[2:].any()
    u, v, depth = scene.project(cam_gt, pt_gt, prob.obs_cam, prob.obs_pt, prob.K)
    assert np.abs(u - prob.obs_u).max() < 1e-9 and np.abs(v - prob.obs_v).max() < 1e-9      # noise-free: observations ARE the projections
    assert depth.min() > 5.0 and depth.max() < 45.0
    assert u.min() >= 0 and u.max() <= scene.IMG_W and v.min() >= 0 and v.max() <= scene.IMG_H
    assert np.bincount(prob.obs_cam, minlength=prob.n_cam).min() > 0
    tl = np.bincount(prob.obs_pt, minlength=prob.n_pt)
    assert tl.min() >= 2 and 4.5 < tl.mean() < 5.5
    assert np.all(np.diff(prob.obs_pt) >= 0)                                                  # track-contiguous
    first = np.cumsum(tl) - tl
    last = first + tl - 1
    span = prob.obs_cam[last] - prob.obs_cam[first]
    assert 0.05 < (span > 20).mean() < 0.5 and span.max() > 60                                # revisits from the next street
    first_cam = prob.obs_cam[first]
    assert np.all(np.diff(first_cam) >= 0)                                                    # creation order
    # weak-scaling shards: the same cameras, points of the rank's own streets only
    a = scene.make_street_grid(n_rows=6, n_cols=60, n_pt=5000, track_len=4, seed=4, shard=0, row_range=(0, 3))
    b = scene.make_street_grid(n_rows=6, n_cols=60, n_pt=5000, track_len=4, seed=4, shard=1, row_range=(3, 6))
    assert np.array_equal(a.cam, b.cam) and a.obs_cam.max() < 4 * 60 and b.obs_cam.min() >= 2 * 60
