"""Independent numpy restatement of the BA path — TEST INFRASTRUCTURE ONLY (small problems).

Purpose: pin oracle/glba_oracle.cpp against a second implementation that shares no code and no
derivation with it: Jacobians by complex-step differentiation of the residual (not dual numbers,
not the closed form the CUDA kernels use), and the LM step from the *dense, un-eliminated*
augmented least-squares problem  min |J y - r|^2 + |D y|^2  (numpy lstsq, no Schur complement).
Follows the same reference call sites: slam_core.cpp:699-733 (residual), :814 (CauchyLoss),
:831-833 (fixed cameras), :842-849 (Ceres LM options).
"""
import numpy as np

EPS = np.finfo(np.float64).eps
DBL_MAX = np.finfo(np.float64).max


def _rotate(aa, pt):
    """ceres::AngleAxisRotatePoint semantics on (n,3) arrays; works for complex dtype."""
    th2 = (aa * aa).sum(axis=1)
    big = th2.real > EPS
    th = np.sqrt(np.where(big, th2, 1.0))
    c, s = np.cos(th), np.sin(th)
    w = aa / th[:, None]
    wxp = np.cross(w, pt)
    tmp = (w * pt).sum(axis=1) * (1.0 - c)
    full = pt * c[:, None] + wxp * s[:, None] + w * tmp[:, None]
    small = pt + np.cross(aa, pt)
    return np.where(big[:, None], full, small)


def residuals(cam, pt, obs_cam, obs_pt, u, v, K):
    fx, fy, cx, cy = K
    c = cam[obs_cam]
    q = pt[obs_pt] - c[:, 3:6]
    p = _rotate(-c[:, :3], q)
    return np.stack([fx * p[:, 0] / p[:, 2] + cx - u, fy * p[:, 1] / p[:, 2] + cy - v], axis=1)


def loss(kind, a, s):
    if kind == 2:      # Cauchy
        b = a * a
        return b * np.log1p(s / b) if False else b * np.log(1.0 + s / b), 1.0 / (1.0 + s / b)
    if kind == 1:      # Huber
        b = a * a
        r = np.sqrt(np.where(s > b, s, 1.0))
        return np.where(s > b, 2 * a * r - b, s), np.where(s > b, a / r, 1.0)
    return s, np.ones_like(s)


def jacobian_complex_step(cam, pt, obs_cam, obs_pt, u, v, K, h=1e-30):
    """Per-observation 2x6 and 2x3 blocks by complex-step differentiation (exact to rounding)."""
    n = obs_cam.shape[0]
    Jc = np.zeros((n, 2, 6))
    Jp = np.zeros((n, 2, 3))
    camc = cam.astype(np.complex128)
    ptc = pt.astype(np.complex128)
    for a in range(6):
        cc = camc.copy()
        cc[:, a] += 1j * h
        Jc[:, :, a] = residuals(cc, ptc, obs_cam, obs_pt, u, v, K).imag / h
    for a in range(3):
        pp = ptc.copy()
        pp[:, a] += 1j * h
        Jp[:, :, a] = residuals(camc, pp, obs_cam, obs_pt, u, v, K).imag / h
    return Jc, Jp


def cost_of(cam, pt, prob, kind, a):
    r = residuals(cam, pt, prob.obs_cam, prob.obs_pt, prob.obs_u, prob.obs_v, prob.K)
    rho, _ = loss(kind, a, (r * r).sum(axis=1))
    c = 0.5 * rho.sum()
    return c if np.isfinite(c) else DBL_MAX


def solve(prob, loss_kind=2, loss_scale=1.0, max_iters=30, function_tol=1e-6, gradient_tol=1e-10,
          parameter_tol=1e-8, initial_radius=1e4, max_radius=1e16, min_radius=1e-32, min_relative_decrease=1e-3,
          min_diag=1e-6, max_diag=1e32):
    cam, pt = prob.cam.copy(), prob.pt.copy()
    oc, op = prob.obs_cam.astype(np.int64), prob.obs_pt.astype(np.int64)
    n_cam, n_pt, n = cam.shape[0], pt.shape[0], oc.shape[0]
    cam_seen = np.bincount(oc, minlength=n_cam) > 0
    pt_seen = np.bincount(op, minlength=n_pt) > 0
    cam_free = cam_seen & ~(prob.cam_fixed.astype(bool) if prob.cam_fixed is not None else np.zeros(n_cam, bool))
    pt_free = pt_seen & ~(prob.pt_fixed.astype(bool) if prob.pt_fixed is not None else np.zeros(n_pt, bool))
    # column index of each parameter in the dense Jacobian (-1 = constant)
    col_c = np.full((n_cam, 6), -1)
    col_p = np.full((n_pt, 3), -1)
    ncol = 0
    for i in np.flatnonzero(cam_free):
        col_c[i] = np.arange(ncol, ncol + 6); ncol += 6
    for j in np.flatnonzero(pt_free):
        col_p[j] = np.arange(ncol, ncol + 3); ncol += 3

    def pack(c, x):
        out = np.zeros(ncol)
        out[col_c[cam_free].ravel()] = c[cam_free].ravel()
        out[col_p[pt_free].ravel()] = x[pt_free].ravel()
        return out

    def linearize(c, x):
        r = residuals(c, x, oc, op, prob.obs_u, prob.obs_v, prob.K)
        s = (r * r).sum(axis=1)
        rho, rho1 = loss(loss_kind, loss_scale, s)
        w = np.sqrt(rho1)
        Jc, Jp = jacobian_complex_step(c, x, oc, op, prob.obs_u, prob.obs_v, prob.K)
        J = np.zeros((2 * n, ncol))
        rows = np.arange(n) * 2
        for row in range(2):
            for a in range(6):
                cols = col_c[oc, a]; m = cols >= 0
                J[rows[m] + row, cols[m]] = (w * Jc[:, row, a])[m]
            for a in range(3):
                cols = col_p[op, a]; m = cols >= 0
                J[rows[m] + row, cols[m]] = (w * Jp[:, row, a])[m]
        return 0.5 * rho.sum(), (w[:, None] * r).ravel(), J

    cost, r, J = linearize(cam, pt)
    g = J.T @ r
    gmax = np.abs(g).max() if ncol else 0.0
    scale = 1.0 / (1.0 + np.sqrt((J * J).sum(axis=0)))
    J = J * scale
    x_norm = np.linalg.norm(pack(cam, pt))
    radius, dec = initial_radius, 2.0
    out = dict(cost=[cost], cost_candidate=[cost], radius=[radius], accepted=[0], stop_reason=0)
    it = 0
    while True:
        if it >= max_iters: out["stop_reason"] = 1; break
        if gmax <= gradient_tol: out["stop_reason"] = 2; break
        if radius <= min_radius: out["stop_reason"] = 5; break
        it += 1
        diag = np.clip((J * J).sum(axis=0), min_diag, max_diag)
        D = np.sqrt(diag / radius)
        A = np.vstack([J, np.diag(D)])
        b = np.concatenate([r, np.zeros(ncol)])
        y = np.linalg.lstsq(A, b, rcond=None)[0]
        step = -y
        m = J @ step
        model_change = -(m @ (r + m / 2.0))
        if not model_change > 0:
            radius /= dec; dec *= 2
            out["cost"].append(cost); out["cost_candidate"].append(cost); out["radius"].append(radius); out["accepted"].append(0)
            continue
        delta = step * scale
        cam_c, pt_c = cam.copy(), pt.copy()
        cam_c[cam_free] += delta[col_c[cam_free].ravel()].reshape(-1, 6)
        pt_c[pt_free] += delta[col_p[pt_free].ravel()].reshape(-1, 3)
        cand = cost_of(cam_c, pt_c, prob, loss_kind, loss_scale)
        out["cost_candidate"].append(cand)
        step_norm = np.linalg.norm(pack(cam, pt) - pack(cam_c, pt_c))
        if step_norm <= parameter_tol * (x_norm + parameter_tol):
            out["cost"].append(cost); out["radius"].append(radius); out["accepted"].append(0); out["stop_reason"] = 3; break
        if abs(cost - cand) <= function_tol * cost:
            out["cost"].append(cost); out["radius"].append(radius); out["accepted"].append(0); out["stop_reason"] = 4; break
        rel = (cost - cand) / model_change
        if rel > min_relative_decrease:
            cam, pt = cam_c, pt_c
            x_norm = np.linalg.norm(pack(cam, pt))
            cost, r, J = linearize(cam, pt)
            gmax = np.abs(J.T @ r).max()
            J = J * scale
            radius = min(max_radius, radius / max(1.0 / 3.0, 1.0 - (2.0 * rel - 1.0) ** 3))
            dec = 2.0
            out["accepted"].append(1)
        else:
            radius /= dec; dec *= 2
            out["accepted"].append(0)
        out["cost"].append(cost); out["radius"].append(radius)
    out["n_iters"] = it
    out["cam"], out["pt"] = cam, pt
    return out
