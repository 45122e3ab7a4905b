"""Independent numpy restatement of the g2o-formulation BA (GLBA_MODE_G2O) — TEST INFRASTRUCTURE ONLY (small problems).

Pins oracle/glba_oracle.cpp::solve_g2o against a second implementation that shares no code with it:
  * the pose update T <- exp([dw, dv]) T through scipy.linalg.expm of the 4x4 twist (not the closed-form V matrix),
  * Jacobians by complex-step differentiation of e(d) = obs - proj((I + [dw]x) p + dv) at d = 0 (not the closed form of
    EdgeProjectXYZ2UV::linearizeOplus),
  * the step from the dense, un-eliminated normal equations (H + lambda I) d = b (no Schur complement),
  * angle-axis <-> matrix through scipy.spatial.transform.Rotation.
Follows docs/old_unorganized/4image_pnp_ba.txt:321-430 / Old/mult_img_recoverpose_single_ba:258-314 and g2o's published
OptimizationAlgorithmLevenberg (lambda0 = tau max diag H; rho = (chi2 - chi2') / (d'(lambda d + b) + 1e-3); accept on
rho > 0 with lambda *= clamp(1 - (2 rho - 1)^3, 1/3, 2/3); else lambda *= nu, nu *= 2; <= max_trials trials).
"""
import numpy as np
from scipy.linalg import expm
from scipy.spatial.transform import Rotation

from .py_oracle import loss as loss_fn


def _project(p, K):
    fx, fy, cx, cy = K
    return np.stack([fx * p[:, 0] / p[:, 2] + cx, fy * p[:, 1] / p[:, 2] + cy], axis=1)


def _errors(R, t, pt, prob):
    p = np.einsum("nij,nj->ni", R[prob.obs_cam], pt[prob.obs_pt]) + t[prob.obs_cam]
    return np.stack([prob.obs_u, prob.obs_v], axis=1) - _project(p, prob.K), p


def _jacobians(R, p, K, h=1e-30):
    n = p.shape[0]
    Jc, Jp = np.zeros((n, 2, 6)), np.zeros((n, 2, 3))
    pc = p.astype(np.complex128)
    for a in range(6):
        d = np.zeros(6, np.complex128)
        d[a] = 1j * h
        q = pc + np.cross(np.broadcast_to(d[:3], pc.shape), pc) + d[3:]
        Jc[:, :, a] = -_project(q, K).imag / h
    for a in range(3):                       # d p / d X = R
        q = pc + 1j * h * R[:, :, a]
        Jp[:, :, a] = -_project(q, K).imag / h
    return Jc, Jp


def solve(prob, loss_kind=2, loss_a=1.0, max_iters=30, tau=1e-5, max_trials=10):
    """prob.cam = [angle-axis of R_cw | t] (world-to-camera).  Returns dict(cost, accepted, lam, n_iters, cam, pt)."""
    R = Rotation.from_rotvec(prob.cam[:, :3]).as_matrix()
    t = prob.cam[:, 3:].copy()
    pt = prob.pt.copy()
    fixed = np.zeros(prob.n_cam, bool) if prob.cam_fixed is None else prob.cam_fixed.astype(bool)
    seen_c = np.zeros(prob.n_cam, bool); seen_c[prob.obs_cam] = True
    seen_p = np.zeros(prob.n_pt, bool); seen_p[prob.obs_pt] = True
    free_c = np.nonzero(~fixed & seen_c)[0]
    free_p = np.nonzero(seen_p)[0]
    slot_c = -np.ones(prob.n_cam, int); slot_c[free_c] = np.arange(len(free_c))
    slot_p = -np.ones(prob.n_pt, int); slot_p[free_p] = np.arange(len(free_p))
    nc, nv = 6 * len(free_c), 6 * len(free_c) + 3 * len(free_p)
    n = prob.n_obs

    info = np.ones(prob.n_pt) if getattr(prob, "pt_info", None) is None else prob.pt_info
    om = info[prob.obs_pt]                   # edge information = om * I

    def linearize(R, t, pt):
        e, p = _errors(R, t, pt, prob)
        rho, w = loss_fn(loss_kind, loss_a, om * (e * e).sum(axis=1))
        w = w * om
        Jc, Jp = _jacobians(R[prob.obs_cam], p, prob.K)
        J = np.zeros((2 * n, nv))
        for k in range(n):
            sc, sp = slot_c[prob.obs_cam[k]], slot_p[prob.obs_pt[k]]
            if sc >= 0:
                J[2 * k:2 * k + 2, 6 * sc:6 * sc + 6] = Jc[k]
            J[2 * k:2 * k + 2, nc + 3 * sp:nc + 3 * sp + 3] = Jp[k]
        W = np.repeat(w, 2)
        H = J.T @ (W[:, None] * J)
        b = -J.T @ (W * e.ravel())
        return 0.5 * rho.sum(), H, b

    def chi(R, t, pt):
        e, _ = _errors(R, t, pt, prob)
        return 0.5 * loss_fn(loss_kind, loss_a, om * (e * e).sum(axis=1))[0].sum()

    cost, H, b = linearize(R, t, pt)
    lam, nu = tau * np.max(np.diag(H)), 2.0
    out = dict(cost=[cost], accepted=[0], lam=[lam], rho=[0.0])
    n_succ = 0
    stop = "max_iters"
    while n_succ < max_iters:
        n_rej, rho = 0, 0.0
        while True:
            d = np.linalg.solve(H + lam * np.eye(nv), b)
            Rn, tn, ptn = R.copy(), t.copy(), pt.copy()
            for s, i in enumerate(free_c):
                tw = np.zeros((4, 4))
                w = d[6 * s:6 * s + 3]
                tw[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
                tw[:3, 3] = d[6 * s + 3:6 * s + 6]
                E = expm(tw)
                Rn[i] = E[:3, :3] @ R[i]
                tn[i] = E[:3, :3] @ t[i] + E[:3, 3]
            ptn[free_p] += d[nc:].reshape(-1, 3)
            cand = chi(Rn, tn, ptn)
            scale = d @ (lam * d + b)
            rho = (cost - cand) / (0.5 * scale + 0.5e-3) if np.isfinite(cand) else -1.0
            if rho > 0 and np.isfinite(cand):
                R, t, pt = Rn, tn, ptn
                cost, H, b = linearize(R, t, pt)
                lam *= max(1.0 / 3.0, min(1.0 - (2.0 * rho - 1.0) ** 3, 2.0 / 3.0))
                nu = 2.0
                n_succ += 1
                out["accepted"].append(1)
            else:
                lam *= nu
                nu *= 2.0
                n_rej += 1
                out["accepted"].append(0)
            out["cost"].append(cost); out["lam"].append(lam); out["rho"].append(rho)
            if not (rho < 0 and n_rej < max_trials):
                break
        if n_rej >= max_trials or rho == 0.0:
            stop = "trials"
            break
    cam = np.concatenate([Rotation.from_matrix(R).as_rotvec(), t], axis=1)
    cam[fixed | ~seen_c] = prob.cam[fixed | ~seen_c]
    return dict(out, n_iters=len(out["cost"]) - 1, n_successful=n_succ, stop=stop, cam=cam, pt=pt)
