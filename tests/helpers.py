import glob
import os

import numpy as np

from gl_slam_b200._abi import HostProblem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")) if "pose_only" not in p and "triangulate" not in p)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    prob = HostProblem(z["cam"], z["pt"], z["obs_cam"], z["obs_pt"], z["obs_u"], z["obs_v"], tuple(z["K"]), z["cam_fixed"], None)
    return prob, z


def max_rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def rel_to_max(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)) if a.size else 0.0


def check_trajectory(got, want, rtol=1e-9):
    """north_star parity gate: same iteration count, per-iteration cost within 1e-9 relative."""
    assert got["n_iters"] == int(want["n_iters"]), (got["n_iters"], int(want["n_iters"]))
    w = np.asarray(want["cost"], float)
    g = np.asarray(got["cost"], float)
    assert g.shape == w.shape
    assert max_rel(g, w) <= rtol, max_rel(g, w)
    assert list(got["accepted"]) == list(np.asarray(want["accepted"]).astype(int)), "accept/reject sequence differs"


def well_conditioned_points(prob, pt_ref, max_depth=150.0):
    """Points whose reference solution stays within max_depth of every observing camera: the ones whose
    coordinates are determined by the data (outlier-hit short tracks legitimately drift to infinity)."""
    d = np.linalg.norm(pt_ref[prob.obs_pt] - prob.cam[prob.obs_cam, 3:6], axis=1)
    far = np.zeros(prob.n_pt, bool)
    np.logical_or.at(far, prob.obs_pt, d > max_depth)
    return ~far


def check_state(prob, got_cam, got_pt, ref_cam, ref_pt, rtol=1e-6):
    """north_star parity gate: final poses / points within 1e-6 relative."""
    assert np.allclose(got_cam, ref_cam, rtol=rtol, atol=rtol * 1e-2), float(np.abs(got_cam - ref_cam).max())
    ok = well_conditioned_points(prob, ref_pt)
    assert ok.mean() > 0.9
    err = np.linalg.norm(got_pt[ok] - ref_pt[ok], axis=1) / np.maximum(np.linalg.norm(ref_pt[ok], axis=1), 1.0)
    assert err.max() <= rtol, float(err.max())
