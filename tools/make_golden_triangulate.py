"""tests/golden/triangulate.npz — inputs and the outputs of the REFERENCE'S OWN dependency for
slam_core::triangulate_and_filter_3d_points (slam_core.cpp:173-256): cv2.triangulatePoints is the very call at :194
(importable in this container; it cannot travel to the GPU box, hence the committed fixture).  The filter
(:206-251) is restated in numpy on top of cv2's homogeneous points.  Run: python tools/make_golden_triangulate.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gl_slam_b200 import scene  # noqa: E402


def reference_filter(X4, R1, t1, R2, t2, K, p0, p1, dist_thr, reproj_thr):
    fx, fy, cx, cy = K
    n = X4.shape[1]
    keep = np.zeros(n, bool)
    X = np.full((n, 3), np.nan)
    for i in range(n):
        w = X4[3, i]
        if abs(w) < 1e-9:
            continue
        X[i] = X4[:3, i] / w
        c1 = (R1 @ X4[:3, i] + t1 * w) / w
        c2 = (R2 @ X4[:3, i] + t2 * w) / w
        if c1[2] <= 0 or c1[2] > dist_thr or c2[2] <= 0 or c2[2] > dist_thr:
            continue
        e1 = np.hypot(fx * c1[0] / c1[2] + cx - p0[i, 0], fy * c1[1] / c1[2] + cy - p0[i, 1])
        e2 = np.hypot(fx * c2[0] / c2[2] + cx - p1[i, 0], fy * c2[1] / c2[2] + cy - p1[i, 1])
        if e1 > reproj_thr or e2 > reproj_thr:
            continue
        keep[i] = True
    return X, keep


def main():
    rng = np.random.default_rng(42)
    K = scene.KITTI_K
    fx, fy, cx, cy = K
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    n = 1500
    # two camera-to-world poses ~1 m apart, converted to world-to-camera as the tracking thread does before the call
    cam = np.array([[0.01, -0.02, 0.005, 0.1, 0.0, 0.0], [0.02, 0.03, -0.01, 0.25, -0.05, 1.05]])
    Rwc = scene.rodrigues(cam[:, :3])
    R1, R2 = Rwc[0].T, Rwc[1].T
    t1, t2 = -R1 @ cam[0, 3:], -R2 @ cam[1, 3:]
    z = rng.uniform(3, 120, n)
    P = np.c_[(rng.uniform(0, 1241, n) - cx) / fx * z, (rng.uniform(0, 376, n) - cy) / fy * z, z]
    P[:40, 2] *= -1                                  # behind the cameras
    Xw = (Rwc[0] @ P.T).T + cam[0, 3:]
    def proj(R, t):
        c = (R @ Xw.T).T + t
        return np.c_[fx * c[:, 0] / c[:, 2] + cx, fy * c[:, 1] / c[:, 2] + cy]
    p0 = proj(R1, t1) + rng.normal(0, 0.4, (n, 2))
    p1 = proj(R2, t2) + rng.normal(0, 0.4, (n, 2))
    p1[100:160] += rng.uniform(-30, 30, (60, 2))     # mismatches
    P0 = Km @ np.c_[R1, t1]
    P1 = Km @ np.c_[R2, t2]
    X4 = cv2.triangulatePoints(P0, P1, p0.T.copy(), p1.T.copy())
    dist_thr, reproj_thr = 100.0, 2.0
    X, keep = reference_filter(X4, R1, t1, R2, t2, K, p0, p1, dist_thr, reproj_thr)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "triangulate.npz"), R1=R1, t1=t1, R2=R2, t2=t2, K=np.array(K), p0=p0, p1=p1,
                        X4=X4, X=X, keep=keep, dist_thr=dist_thr, reproj_thr=reproj_thr, cv2_version=cv2.__version__)
    print(f"triangulate: {n} matches, cv2 {cv2.__version__}: kept {int(keep.sum())}")


if __name__ == "__main__":
    main()
