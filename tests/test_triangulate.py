"""Two-view triangulation + filter (slam_core.cpp:173-256) against golden vectors made with the reference's own
dependency, cv2.triangulatePoints (tools/make_golden_triangulate.py)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN


def _load():
    return np.load(os.path.join(GOLDEN, "triangulate.npz"))


def test_golden_is_consistent_with_cv2_if_present():
    """Re-derives the fixture where cv2 exists (this container): guards the committed vectors against drift."""
    cv2 = pytest.importorskip("cv2")
    z = _load()
    K = z["K"]
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]])
    X4 = cv2.triangulatePoints(Km @ np.c_[z["R1"], z["t1"]], Km @ np.c_[z["R2"], z["t2"]], z["p0"].T.copy(), z["p1"].T.copy())
    ok = np.abs(X4[3]) > 1e-9
    assert np.allclose((X4[:3] / X4[3]).T[ok], (z["X4"][:3] / z["X4"][3]).T[ok], rtol=1e-9, atol=1e-9)
    # DLT solution really is the null vector of the 4x4 system (independent check of what "triangulatePoints" means)
    i = 7
    P0, P1 = Km @ np.c_[z["R1"], z["t1"]], Km @ np.c_[z["R2"], z["t2"]]
    A = np.stack([z["p0"][i, 0] * P0[2] - P0[0], z["p0"][i, 1] * P0[2] - P0[1], z["p1"][i, 0] * P1[2] - P1[0], z["p1"][i, 1] * P1[2] - P1[1]])
    v = np.linalg.svd(A)[2][-1]
    assert np.allclose(v[:3] / v[3], z["X4"][:3, i] / z["X4"][3, i], rtol=1e-8)


@pytest.mark.gpu
def test_triangulate_filter_matches_reference_dependency(ctx):
    z = _load()
    X, keep = ctx.triangulate_filter(z["R1"], z["t1"], z["R2"], z["t2"], tuple(z["K"]), z["p0"], z["p1"], float(z["dist_thr"]),
                                     float(z["reproj_thr"]))
    want_keep = z["keep"]
    # decisions: identical except observations within 1e-6 px / 1e-6 m of a threshold (none in this fixture)
    assert np.array_equal(keep, want_keep), (int(keep.sum()), int(want_keep.sum()))
    assert keep.sum() > 1000 and (~keep).sum() > 90
    got, ref = X[want_keep], z["X"][want_keep]
    err = np.linalg.norm(got - ref, axis=1) / np.linalg.norm(ref, axis=1)
    assert err.max() < 1e-9, err.max()
    # every finite reference point (kept or not) agrees too
    fin = np.isfinite(z["X"]).all(axis=1)
    err = np.linalg.norm(X[fin] - z["X"][fin], axis=1) / np.maximum(np.linalg.norm(z["X"][fin], axis=1), 1.0)
    assert np.median(err) < 1e-12 and err.max() < 1e-6


@pytest.mark.gpu
def test_triangulate_edge_cases(ctx):
    z = _load()
    X, keep = ctx.triangulate_filter(z["R1"], z["t1"], z["R2"], z["t2"], tuple(z["K"]), np.zeros((0, 2)), np.zeros((0, 2)), 100.0, 2.0)
    assert X.shape == (0, 3) and keep.shape == (0,)
    # identical cameras: every match is degenerate (no baseline) -> finite output or rejection, never a crash
    X, keep = ctx.triangulate_filter(z["R1"], z["t1"], z["R1"], z["t1"], tuple(z["K"]), z["p0"][:50], z["p0"][:50], 100.0, 2.0)
    assert keep.shape == (50,)
