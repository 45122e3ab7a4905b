"""Seeded synthetic BA scenes for the configs in BASELINE.json / BASELINE.md §4 (C1..C5).

All scenes use the intrinsics GL-SLAM reads from KITTI calib.txt (slam_core.cpp:38-57): fx=fy=718.856,
cx=607.1928, cy=185.2157, 1241x376.  Cameras are camera-to-world [angle-axis(R_wc), centre], exactly the
parameter block full_ba packs (slam_core.cpp:768-776).  Observations are emitted track-contiguous
(sorted by point, then camera).  Pure numpy, vectorised: C5 (30 M observations) builds in seconds.
"""
import numpy as np

from ._abi import HostProblem

KITTI_K = (718.856, 718.856, 607.1928, 185.2157)
IMG_W, IMG_H = 1241.0, 376.0


def rodrigues(w):
    """angle-axis (n,3) -> rotation matrices (n,3,3) (cv::Rodrigues semantics)."""
    w = np.asarray(w, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(w, axis=1)
    small = th < 1e-12
    k = w / np.where(small, 1.0, th)[:, None]
    K = np.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    R = np.eye(3)[None] + s * K + (1 - c) * (K @ K)
    R[small] = np.eye(3)
    return R


def rotation_to_angle_axis(R):
    """rotation matrices (n,3,3) -> angle-axis (n,3); valid for angles < pi."""
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    v = np.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], axis=1) * 0.5
    s = np.linalg.norm(v, axis=1)
    c = np.clip((np.trace(R, axis1=1, axis2=2) - 1.0) * 0.5, -1.0, 1.0)
    th = np.arctan2(s, c)
    scale = np.where(s < 1e-12, 1.0, th / np.where(s < 1e-12, 1.0, s))
    return v * scale[:, None]


def project(cam, pt, obs_cam, obs_pt, K):
    """Pinhole projection of the reference residual (slam_core.cpp:705-726): returns (u, v, depth)."""
    fx, fy, cx, cy = K
    R = rodrigues(cam[:, :3])                    # R_wc
    q = pt[obs_pt] - cam[obs_cam, 3:6]
    p = np.einsum("nji,nj->ni", R[obs_cam], q)   # R_wc^T q
    return fx * p[:, 0] / p[:, 2] + cx, fy * p[:, 1] / p[:, 2] + cy, p[:, 2]


def _trajectory(n_cam, step, yaw_per_frame, loop):
    """Camera-to-world poses on a planar arc: camera looks along +z, yaw about +y (KITTI convention)."""
    if loop:
        yaw_per_frame = 2.0 * np.pi / n_cam
    yaw = np.arange(n_cam) * yaw_per_frame
    fwd = np.stack([np.sin(yaw), np.zeros(n_cam), np.cos(yaw)], axis=1)
    c = np.concatenate([np.zeros((1, 3)), np.cumsum(fwd[:-1] * step, axis=0)], axis=0)
    w = np.stack([np.zeros(n_cam), yaw, np.zeros(n_cam)], axis=1)
    # keep |w| < pi so angle-axis stays in the principal branch
    w[:, 1] = (w[:, 1] + np.pi) % (2 * np.pi) - np.pi
    return np.concatenate([w, c], axis=1)


def make_scene(n_cam, n_pt, track_len, seed, *, step=0.8, yaw_per_frame=np.deg2rad(0.5), loop=False,
               pixel_sigma=0.5, outlier_frac=0.0, outlier_px=(10.0, 50.0), rot_sigma=0.02, pos_sigma=0.05,
               pt_sigma=0.10, depth=(5.0, 50.0), n_fixed=2, K=KITTI_K, return_gt=False, min_parallax_deg=1.0,
               creation_order=True, shard=0, start_range=None):
    """track_len: int, or callable(rng, n_pt) -> int array (clipped to [2, n_cam]).

    Returns a HostProblem whose cam/pt are the *initial guess* (ground truth + Gaussian noise; the first
    n_fixed cameras stay at ground truth and are flagged fixed, slam_core.cpp:831-833).
    """
    # cameras (trajectory + initial-guess noise) depend on `seed` only, so every rank of a sharded run builds the
    # same replicated camera set; points / observations additionally depend on `shard`
    rng_cam = np.random.default_rng([seed, 0])
    rng = np.random.default_rng([seed, 1, shard])
    fx, fy, cx, cy = K
    cam_gt = _trajectory(n_cam, step, yaw_per_frame, loop)
    if callable(track_len):
        tl = np.asarray(track_len(rng, n_pt), dtype=np.int64)
    else:
        tl = np.full(n_pt, int(track_len), dtype=np.int64)
    tl = np.clip(tl, 2 if n_cam >= 2 else 1, n_cam)
    if start_range is not None:
        start = rng.integers(start_range[0], start_range[1], size=n_pt)
    elif loop:
        start = rng.integers(0, n_cam, size=n_pt)
    else:
        start = (rng.random(n_pt) * (n_cam - tl + 1)).astype(np.int64)
    if creation_order:
        # GL-SLAM hands out map-point ids in creation order (next_point_id++ while the newest keyframe is inserted,
        # slam_core.cpp:386-405), so point ids are sorted by their first-observing keyframe
        order0 = np.argsort(start, kind="stable")
        start, tl = start[order0], tl[order0]
    n_obs = int(tl.sum())
    obs_pt = np.repeat(np.arange(n_pt, dtype=np.int64), tl)
    first = np.cumsum(tl) - tl
    within = np.arange(n_obs, dtype=np.int64) - np.repeat(first, tl)
    obs_cam = (np.repeat(start, tl) + within) % n_cam
    # anchor each point in the frustum of the middle camera of its track; like a real front-end's
    # triangulation filter, re-draw points whose parallax between the first and last camera of the
    # track is below min_parallax_deg (forward motion leaves the depth of points near the focus of
    # expansion unobservable, which makes any BA chaotic at the 1e-16 level).
    anchor = (start + tl // 2) % n_cam
    last = (start + tl - 1) % n_cam
    Rg = rodrigues(cam_gt[:, :3])
    pt_gt = np.zeros((n_pt, 3))
    todo = np.arange(n_pt)
    cos_min = np.cos(np.deg2rad(min_parallax_deg))
    for _round in range(40):
        m = todo.shape[0]
        if m == 0:
            break
        # keep the point in front of the last camera of its track as well (the rig moves towards it)
        zmin = depth[0] + step * (tl[todo] - tl[todo] // 2)
        z = zmin + rng.random(m) * np.maximum(depth[1] - zmin, 1.0)
        ua = rng.uniform(0.05 * IMG_W, 0.95 * IMG_W, size=m)
        va = rng.uniform(0.05 * IMG_H, 0.95 * IMG_H, size=m)
        p_cam = np.stack([(ua - cx) / fx * z, (va - cy) / fy * z, z], axis=1)
        X = np.einsum("nij,nj->ni", Rg[anchor[todo]], p_cam) + cam_gt[anchor[todo], 3:6]
        pt_gt[todo] = X
        r0 = X - cam_gt[start[todo], 3:6]
        r1 = X - cam_gt[last[todo], 3:6]
        cosang = (r0 * r1).sum(1) / (np.linalg.norm(r0, axis=1) * np.linalg.norm(r1, axis=1))
        todo = todo[cosang > cos_min] if min_parallax_deg > 0 else todo[:0]
    # sort each track by camera index (wrap-around tracks are not monotone)
    order = np.lexsort((obs_cam, obs_pt))
    obs_cam, obs_pt = obs_cam[order], obs_pt[order]
    u, v, dep = project(cam_gt, pt_gt, obs_cam, obs_pt, K)
    u = u + rng.normal(0.0, pixel_sigma, size=n_obs)
    v = v + rng.normal(0.0, pixel_sigma, size=n_obs)
    if outlier_frac > 0:
        bad = rng.random(n_obs) < outlier_frac
        mag = rng.uniform(outlier_px[0], outlier_px[1], size=n_obs)
        ang = rng.uniform(0, 2 * np.pi, size=n_obs)
        u = np.where(bad, u + mag * np.cos(ang), u)
        v = np.where(bad, v + mag * np.sin(ang), v)
    cam0 = cam_gt.copy()
    cam0[n_fixed:, :3] += rng_cam.normal(0.0, rot_sigma, size=(n_cam - n_fixed, 3)) if n_cam > n_fixed else 0.0
    cam0[n_fixed:, 3:] += rng_cam.normal(0.0, pos_sigma, size=(n_cam - n_fixed, 3)) if n_cam > n_fixed else 0.0
    pt0 = pt_gt + rng.normal(0.0, pt_sigma, size=(n_pt, 3))
    cam_fixed = np.zeros(n_cam, dtype=np.uint8)
    cam_fixed[:n_fixed] = 1
    prob = HostProblem(cam0, pt0, obs_cam.astype(np.int32), obs_pt.astype(np.int32), u, v, K, cam_fixed, None)
    if return_gt:
        return prob, cam_gt, pt_gt
    return prob


def make_street_grid(n_rows, n_cols, n_pt, track_len, seed, *, step=0.8, gap=7.5, revisit_frac=0.35, depth=(10.0, 40.0), pixel_sigma=0.5,
                     outlier_frac=0.0, outlier_px=(10.0, 50.0), rot_sigma=0.002, pos_sigma=0.03, pt_sigma=0.10, n_fixed=2, K=KITTI_K,
                     min_parallax_deg=1.0, shard=0, row_range=None, return_gt=False):
    """SURVEY §8d's C4: cameras on an n_rows x n_cols STREET GRID driven in serpentine order (row r is a street along +Z
    at X = r * gap; even rows are driven towards +Z, odd rows back towards -Z, so keyframe ids run 0 .. n_rows*n_cols-1
    along the path).  Covisibility is banded (consecutive keyframes of a street see a point ahead of them) PLUS revisits:
    a fraction `revisit_frac` of the points lies between two neighbouring streets and is also seen, from the other side,
    by keyframes of the next street — ids ~2*(n_cols - c) apart, so its track is a run of consecutive cameras plus a
    second run far away in id: loop-closure structure in every tile of the map.
    Every observation is inside the 1241 x 376 image with depth in `depth` by construction.  Point ids are in creation
    order (sorted by first-observing keyframe, as GL-SLAM numbers map points).  row_range=(r0, r1) restricts the POINTS
    to primary streets r0 <= r < r1 (weak-scaling shards); cameras are always the whole grid."""
    rng_cam = np.random.default_rng([seed, 0])
    rng = np.random.default_rng([seed, 1, shard])
    fx, fy, cx, cy = K
    n_cam = n_rows * n_cols
    r_idx = np.repeat(np.arange(n_rows), n_cols)
    c_idx = np.tile(np.arange(n_cols), n_rows)
    sgn = np.where(r_idx % 2 == 0, 1.0, -1.0)                               # driving / viewing direction along Z
    z_cam = np.where(r_idx % 2 == 0, c_idx, n_cols - 1 - c_idx) * step
    cam_gt = np.zeros((n_cam, 6))
    cam_gt[:, 1] = np.where(r_idx % 2 == 0, 0.0, np.pi - 1e-3) + rng_cam.normal(0.0, 0.01, n_cam)      # yaw about +Y
    cam_gt[:, 3] = r_idx * gap
    cam_gt[:, 5] = z_cam
    tl = np.asarray(track_len(rng, n_pt), dtype=np.int64) if callable(track_len) else np.full(n_pt, int(track_len), dtype=np.int64)
    tl = np.clip(tl, 2, 2 * n_cols)
    r_lo, r_hi = (0, n_rows) if row_range is None else row_range
    # per point: primary street, second street (revisit), run lengths, position; redraw what does not fit
    r0 = np.zeros(n_pt, np.int64); r1 = np.zeros(n_pt, np.int64)
    l1 = np.zeros(n_pt, np.int64); l2 = np.zeros(n_pt, np.int64)
    c1 = np.zeros(n_pt, np.int64); c2 = np.zeros(n_pt, np.int64)            # first column (in id order) of each run
    X = np.zeros((n_pt, 3))
    todo = np.arange(n_pt)
    cos_min = np.cos(np.deg2rad(min_parallax_deg))
    d_near = (0.6 * depth[0], depth[0])                                      # closest allowed depth: one-street points / revisited points
    d_far = (0.625 * depth[1], 0.75 * depth[1])
    for _round in range(200):
        m = todo.shape[0]
        if m == 0:
            break
        pr = rng.integers(r_lo, r_hi, size=m)
        rev = (rng.random(m) < revisit_frac) & (n_rows > 1)
        nb = np.where(pr + 1 < n_rows, pr + 1, pr - 1)                      # the neighbouring street that revisits
        dmin = np.where(rev, d_near[1], d_near[0]); dmax = np.where(rev, d_far[1], d_far[0])
        slots = ((dmax - dmin) / step).astype(np.int64)
        a = np.where(rev, np.maximum(1, tl[todo] // 2), tl[todo])
        b = np.where(rev, tl[todo] - a, 0)
        a = np.minimum(a, np.minimum(slots, max(1, n_cols - 3))); b = np.minimum(b, np.minimum(slots, max(1, n_cols - 3)))
        s0 = np.where(pr % 2 == 0, 1.0, -1.0)
        zp = rng.uniform(-0.7 * d_far[0], (n_cols - 1) * step + 0.7 * d_far[0], size=m)      # beyond the street ends too: every keyframe sees points
        # street pr (viewing direction s0) sees the point from Z_cam = zp - s0 * d; the neighbour (direction -s0) from
        # Z_cam = zp + s0 * d'.  Runs are consecutive grid positions.
        d_last = dmin + rng.random(m) * np.maximum((slots - a) * step, 0.0)  # depth of the primary run's LAST (closest) camera
        k_last = np.round((zp - s0 * d_last) / step).astype(np.int64)        # its grid slot along Z
        k_first = k_last - (s0 * (a - 1)).astype(np.int64)                   # farthest camera (driven earlier)
        ok = (np.minimum(k_first, k_last) >= 0) & (np.maximum(k_first, k_last) <= n_cols - 1)
        zp = k_last * step + s0 * d_last                                     # snap the point: depths are exact
        d2_near = dmin + rng.random(m) * np.maximum((slots - b) * step, 0.0)
        k2_near = np.round((zp + s0 * d2_near) / step).astype(np.int64)      # neighbour run: closest camera ...
        k2_far = k2_near + (s0 * (b - 1)).astype(np.int64)                   # ... and farthest (driven first on that street)
        d2 = s0 * (k2_near * step - zp)
        ok &= ~rev | ((np.minimum(k2_far, k2_near) >= 0) & (np.maximum(k2_far, k2_near) <= n_cols - 1) &
                      (d2 >= dmin - 1e-9) & (d2 + (b - 1) * step <= dmax + 1e-9))
        # lateral position: between the two streets for revisited points, else a fraction of the closest depth to either side
        side = np.where(nb > pr, 1.0, -1.0)
        ratio = rng.uniform(0.15, 0.72, size=m) * np.where(rng.random(m) < 0.5, 1.0, -1.0)
        xp = np.where(rev, pr * gap + side * gap * rng.uniform(0.3, 0.7, size=m), pr * gap + ratio * d_last)
        yp = rng.uniform(-0.22, 0.2, size=m) * np.where(rev, np.minimum(d_last, d2), d_last)
        # parallax between the first and the last camera of the track (forward motion towards a point near the axis has none)
        col_first = np.where(pr % 2 == 0, k_first, n_cols - 1 - k_first)
        col_last = np.where(pr % 2 == 0, k_last, n_cols - 1 - k_last)
        col2_near = np.where(nb % 2 == 0, k2_near, n_cols - 1 - k2_near)
        col2_far = np.where(nb % 2 == 0, k2_far, n_cols - 1 - k2_far)
        i_first = pr * n_cols + np.clip(col_first, 0, n_cols - 1)
        i_last = np.where(rev, nb * n_cols + np.clip(col2_near, 0, n_cols - 1), pr * n_cols + np.clip(col_last, 0, n_cols - 1))
        P = np.stack([xp, yp, zp], axis=1)
        ra = P - cam_gt[i_first, 3:6]
        rb = P - cam_gt[i_last, 3:6]
        cosang = (ra * rb).sum(1) / (np.linalg.norm(ra, axis=1) * np.linalg.norm(rb, axis=1))
        ok &= cosang <= cos_min
        good = todo[ok]
        r0[good] = pr[ok]; r1[good] = nb[ok]; l1[good] = a[ok]; l2[good] = b[ok]
        c1[good] = np.minimum(col_first, col_last)[ok]
        c2[good] = np.minimum(col2_far, col2_near)[ok]
        X[good] = P[ok]
        todo = todo[~ok]
    if todo.shape[0]:
        raise RuntimeError("street grid: could not place %d points" % todo.shape[0])
    tl = l1 + l2
    # creation order: ids sorted by the first-observing keyframe
    first_cam = np.minimum(r0 * n_cols + c1, np.where(l2 > 0, r1 * n_cols + c2, n_cam))
    order0 = np.argsort(first_cam, kind="stable")
    r0, r1, l1, l2, c1, c2, X, tl = r0[order0], r1[order0], l1[order0], l2[order0], c1[order0], c2[order0], X[order0], tl[order0]
    n_obs = int(tl.sum())
    obs_pt = np.repeat(np.arange(n_pt, dtype=np.int64), tl)
    first = np.cumsum(tl) - tl
    within = np.arange(n_obs, dtype=np.int64) - np.repeat(first, tl)
    in_first = within < np.repeat(l1, tl)
    obs_cam = np.where(in_first, np.repeat(r0 * n_cols + c1, tl) + within, np.repeat(r1 * n_cols + c2, tl) + within - np.repeat(l1, tl))
    order = np.lexsort((obs_cam, obs_pt))
    obs_cam, obs_pt = obs_cam[order], obs_pt[order]
    u, v, dep = project(cam_gt, X, obs_cam, obs_pt, K)
    assert dep.min() > 1.0 and u.min() > -20 and u.max() < IMG_W + 20 and v.min() > -20 and v.max() < IMG_H + 20, \
        (dep.min(), u.min(), u.max(), v.min(), v.max())
    u = u + rng.normal(0.0, pixel_sigma, size=n_obs)
    v = v + rng.normal(0.0, pixel_sigma, size=n_obs)
    if outlier_frac > 0:
        bad = rng.random(n_obs) < outlier_frac
        mag = rng.uniform(outlier_px[0], outlier_px[1], size=n_obs)
        ang = rng.uniform(0, 2 * np.pi, size=n_obs)
        u = np.where(bad, u + mag * np.cos(ang), u)
        v = np.where(bad, v + mag * np.sin(ang), v)
    cam0 = cam_gt.copy()
    cam0[n_fixed:, :3] += rng_cam.normal(0.0, rot_sigma, size=(n_cam - n_fixed, 3))
    cam0[n_fixed:, 3:] += rng_cam.normal(0.0, pos_sigma, size=(n_cam - n_fixed, 3))
    pt0 = X + rng.normal(0.0, pt_sigma, size=(n_pt, 3))
    cam_fixed = np.zeros(n_cam, dtype=np.uint8)
    cam_fixed[:n_fixed] = 1
    prob = HostProblem(cam0, pt0, obs_cam.astype(np.int32), obs_pt.astype(np.int32), u, v, K, cam_fixed, None)
    if return_gt:
        return prob, cam_gt, X
    return prob


def _poisson_tracks(base, lam):
    return lambda rng, n: base + rng.poisson(lam, size=n)


def config(name, scale=1.0, **overrides):
    """Named configs of BASELINE.md §4.  `scale` shrinks n_pt (and, for C4/C5, n_cam) for quick tests."""
    name = name.upper()
    s = float(scale)
    if name == "C1":      # two-view pair, ~500 points, both cameras fixed in the live path
        kw = dict(n_cam=2, n_pt=max(8, int(500 * s)), track_len=2, seed=1, step=1.0, yaw_per_frame=np.deg2rad(0.5),
                  rot_sigma=0.0, pos_sigma=0.0, pt_sigma=0.3, n_fixed=2)
    elif name == "C2":    # local window: 10 keyframes, 5 000 points, 20 000 observations, 5 % outliers
        kw = dict(n_cam=10, n_pt=max(16, int(5000 * s)), track_len=4, seed=2, outlier_frac=0.05,
                  rot_sigma=0.02, pos_sigma=0.05, pt_sigma=0.10)
    elif name == "C3":    # 200 frames, 200 k points, mean track 5 -> 1 M observations (as one problem)
        kw = dict(n_cam=200, n_pt=max(64, int(200000 * s)), track_len=_poisson_tracks(2, 3.0), seed=3,
                  rot_sigma=0.005, pos_sigma=0.05, pt_sigma=0.10)
    elif name == "C4":    # 1 800 cameras on a 60 x 30 street grid (banded covisibility + revisits), 1 M points, ~5 M observations
        rows = max(2, int(round(30 * min(1.0, s * 4))))
        kw = dict(n_rows=rows, n_cols=60, n_pt=max(64, int(1000000 * s)), track_len=_poisson_tracks(2, 3.0), seed=4,
                  rot_sigma=0.002, pos_sigma=0.03, pt_sigma=0.10)
        kw.update(overrides)
        return make_street_grid(**kw)
    elif name == "C4LOOP":    # round-1 stand-in for C4: 1 800 cameras on ONE closed loop (perfectly banded covisibility)
        kw = dict(n_cam=max(8, int(1800 * min(1.0, s * 4))), n_pt=max(64, int(1000000 * s)),
                  track_len=_poisson_tracks(2, 3.0), seed=4, loop=True, rot_sigma=0.002, pos_sigma=0.03, pt_sigma=0.10)
    elif name == "C5":    # 10 k cameras, 4 M points, ~30 M observations, 10 % outliers, Huber
        kw = dict(n_cam=max(8, int(10000 * min(1.0, s * 4))), n_pt=max(64, int(4000000 * s)),
                  track_len=_poisson_tracks(2, 5.5), seed=5, loop=True, outlier_frac=0.10,
                  rot_sigma=0.002, pos_sigma=0.03, pt_sigma=0.10)
    else:
        raise KeyError(name)
    kw.update(overrides)
    return make_scene(**kw)


def config_weak(name, world, rank, scale=1.0):
    """Weak-scaling family: an N-times larger loop map (N x cameras, N x points) of which rank r generates and owns
    the arc whose tracks start at cameras [r*n_cam1, (r+1)*n_cam1) — the contiguous block shard_by_point would hand it.
    Cameras are identical on every rank.  world=1 is the plain config."""
    name = name.upper()
    if name not in ("C4", "C5"):
        raise KeyError("weak scaling is defined for the loop maps C4 / C5")
    if name == "C4":      # N x 30 streets of 60 keyframes; rank r owns the points whose primary street lies in its block of 30
        rows1 = max(2, int(round(30 * min(1.0, scale * 4))))
        return make_street_grid(n_rows=rows1 * world, n_cols=60, n_pt=max(64, int(1000000 * scale)), track_len=_poisson_tracks(2, 3.0), seed=4,
                                rot_sigma=0.002, pos_sigma=0.03, pt_sigma=0.10, shard=rank,
                                row_range=(rank * rows1, (rank + 1) * rows1) if world > 1 else None)
    base = dict(C4=(1800, 1000000, 3.0, 4, 0.0), C5=(10000, 4000000, 5.5, 5, 0.10))[name]
    n_cam1 = max(8, int(base[0] * min(1.0, scale * 4)))
    n_pt1 = max(64, int(base[1] * scale))
    return make_scene(n_cam=n_cam1 * world, n_pt=n_pt1, track_len=_poisson_tracks(2, base[2]), seed=base[3], loop=True,
                      outlier_frac=base[4], rot_sigma=0.002, pos_sigma=0.03, pt_sigma=0.10, shard=rank,
                      start_range=(rank * n_cam1, (rank + 1) * n_cam1) if world > 1 else None)


def pose_only_scene(n, seed, K=KITTI_K, pixel_sigma=0.5, outlier_frac=0.05, rot_sigma=0.01, pos_sigma=0.05):
    """One frame against n fixed 3-D points (tracking-thread input, thread_pool.cpp:149-199)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    cam_gt = np.array([[0.01, 0.05, -0.02, 0.3, -0.1, 5.0]])
    z = rng.uniform(5, 50, size=n)
    ua = rng.uniform(0.02 * IMG_W, 0.98 * IMG_W, size=n)
    va = rng.uniform(0.02 * IMG_H, 0.98 * IMG_H, size=n)
    p_cam = np.stack([(ua - cx) / fx * z, (va - cy) / fy * z, z], axis=1)
    R = rodrigues(cam_gt[:, :3])[0]
    X = p_cam @ R.T + cam_gt[0, 3:]
    uv = np.stack([ua, va], axis=1) + rng.normal(0, pixel_sigma, size=(n, 2))
    bad = rng.random(n) < outlier_frac
    uv[bad] += rng.uniform(-40, 40, size=(int(bad.sum()), 2))
    cam0 = cam_gt[0].copy()
    cam0[:3] += rng.normal(0, rot_sigma, 3)
    cam0[3:] += rng.normal(0, pos_sigma, 3)
    return cam0, X, uv, cam_gt[0]


def shard_by_point(prob, world, rank):
    """SURVEY §8e partitioning: whole point tracks per rank, balanced by observation count.

    Points are dealt out in contiguous blocks whose observation counts are as equal as possible, so
    every rank keeps a track-contiguous slice; cameras are replicated.  Returns (HostProblem, pt_index)
    where pt_index maps the shard's local points back to the global ones.
    """
    n_pt = prob.n_pt
    counts = np.bincount(prob.obs_pt, minlength=n_pt)
    csum = np.concatenate([[0], np.cumsum(counts)])
    total = csum[-1]
    bounds = [int(np.searchsorted(csum, total * r / world, side="left")) for r in range(world + 1)]
    bounds[0], bounds[-1] = 0, n_pt
    lo, hi = bounds[rank], bounds[rank + 1]
    sel = (prob.obs_pt >= lo) & (prob.obs_pt < hi)
    pt_index = np.arange(lo, hi)
    sub = HostProblem(prob.cam, prob.pt[lo:hi], prob.obs_cam[sel], prob.obs_pt[sel] - lo, prob.obs_u[sel],
                      prob.obs_v[sel], prob.K, prob.cam_fixed,
                      None if prob.pt_fixed is None else prob.pt_fixed[lo:hi],
                      None if getattr(prob, "pt_info", None) is None else prob.pt_info[lo:hi])
    return sub, pt_index


def to_world_to_camera(cam):
    """[angle-axis of R_wc | centre] (the live Ceres path, slam_core.cpp:703-713) -> [angle-axis of R_cw | t] with
    p = R_cw X + t (g2o VertexSE3Expmap, docs/old_unorganized/4image_pnp_ba.txt:350-357).  R_cw = R_wc', t = -R_cw c."""
    cam = np.asarray(cam, float).reshape(-1, 6)
    R_cw = np.transpose(rodrigues(cam[:, :3]), (0, 2, 1))
    return np.concatenate([-cam[:, :3], -np.einsum("nij,nj->ni", R_cw, cam[:, 3:])], axis=1)


def to_camera_to_world(cam):
    """Inverse of to_world_to_camera."""
    cam = np.asarray(cam, float).reshape(-1, 6)
    R_cw = rodrigues(cam[:, :3])
    return np.concatenate([-cam[:, :3], -np.einsum("nji,nj->ni", R_cw, cam[:, 3:])], axis=1)


def as_g2o(prob):
    """The same scene with poses in the GLBA_MODE_G2O convention."""
    q = prob.copy()
    q.cam = to_world_to_camera(prob.cam)
    return q
