"""Numpy experiment behind DESIGN.md 7b: does a stronger preconditioner than block-Jacobi pay in the mode bench.py quotes?

Builds the reduced camera matrix S of a street-grid map with the full C4 camera count (1 800) and fewer points, at the
initial point and after two LM iterations (where the bench's PCG hits its 40-iteration cap), and counts PCG iterations to
1e-1 / 1e-2 / 1e-3 / 1e-6 for block-Jacobi and for a two-level additive preconditioner: block-Jacobi + P A_c^-1 P' with
piecewise-constant aggregates of m consecutive cameras (6 coarse unknowns per aggregate, exact coarse solve).
CPU only; uses the oracle to linearise.  usage: python tools/precond_experiment.py [n_points=100000]"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gl_slam_b200 import scene  # noqa: E402
from oracle import oracle  # noqa: E402


def build(prob, radius):
    L = oracle.linearize(prob, radius, per_obs=True)
    n_cam, n_pt, n = prob.n_cam, prob.n_pt, prob.n_obs
    Jc = L.jac_cam.reshape(n, 2, 6); Jp = L.jac_pt.reshape(n, 2, 3); r = L.residuals.reshape(n, 2)
    free = np.where(prob.cam_fixed == 0)[0]; slot = -np.ones(n_cam, int); slot[free] = np.arange(len(free)); nf = len(free)
    sc = slot[prob.obs_cam]; keep = sc >= 0
    rows = (2 * np.arange(n)[:, None, None] + np.arange(2)[None, :, None])
    colc = (6 * sc[:, None, None] + np.arange(6)[None, None, :])
    m = keep[:, None, None] & np.ones((1, 2, 6), bool)
    JC = sp.csr_matrix((Jc[m], (np.broadcast_to(rows, Jc.shape)[m], np.broadcast_to(colc, Jc.shape)[m])), shape=(2 * n, 6 * nf))
    colp = (3 * prob.obs_pt[:, None, None] + np.arange(3)[None, None, :])
    JP = sp.csr_matrix((Jp.ravel(), (np.broadcast_to(rows, Jp.shape).ravel(), np.broadcast_to(colp, Jp.shape).ravel())), shape=(2 * n, 3 * n_pt))
    rr = r.ravel()
    Hcc = (JC.T @ JC).tocsr(); Hpp = (JP.T @ JP).tocsr(); E = (JC.T @ JP).tocsr()
    Hcc = Hcc + sp.diags(Hcc.diagonal() / radius); Hpp = Hpp + sp.diags(Hpp.diagonal() / radius)
    C = np.zeros((n_pt, 3, 3)); Hc = Hpp.tocoo(); C[Hc.row // 3, Hc.row % 3, Hc.col % 3] = Hc.data
    Cinv = sp.bsr_matrix((np.linalg.inv(C), np.arange(n_pt), np.arange(n_pt + 1)), shape=(3 * n_pt, 3 * n_pt))
    S = (Hcc - E @ Cinv @ E.T).toarray()
    rhs = JC.T @ rr - E @ (Cinv @ (JP.T @ rr))
    return S, rhs, nf


def pcg(S, b, Minv, tol, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; rz0 = rz
    for it in range(1, maxit + 1):
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q; z = Minv(r); rz1 = r @ z
        if np.sqrt(abs(rz1)) <= tol * np.sqrt(rz0):
            return it
        p = z + (rz1 / rz) * p; rz = rz1
    return maxit


def block_jacobi(S):
    n = S.shape[0] // 6
    inv = np.stack([np.linalg.inv(S[6 * i:6 * i + 6, 6 * i:6 * i + 6]) for i in range(n)])
    return lambda r: np.einsum('nij,nj->ni', inv, r.reshape(n, 6)).ravel()


def main():
    npt = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    tols = (1e-1, 1e-2, 1e-3, 1e-6)
    prob = scene.make_street_grid(30, 60, npt, track_len=lambda rng, n: 2 + rng.poisson(3.0, size=n), seed=7, rot_sigma=0.002, pos_sigma=0.03)
    print("cameras", prob.n_cam, "observations", prob.n_obs, flush=True)
    later, so = oracle.solve(prob, oracle.options(max_iters=2))
    for label, state, radius in (("initial point", prob, 1e4), ("after two LM iterations", later, 9e4)):
        t = time.time(); S, rhs, nf = build(state, radius)
        B = block_jacobi(S)
        print(f"{label}, radius {radius:g} (S built in {time.time() - t:.1f} s): block-Jacobi", {tol: pcg(S, rhs, B, tol) for tol in tols}, flush=True)
        for m in (4, 10, 15, 30):
            nagg = (nf + m - 1) // m
            P = np.zeros((6 * nf, 6 * nagg))
            for i in range(nf):
                for d in range(6):
                    P[6 * i + d, 6 * (i // m) + d] = 1.0
            fc = sla.cho_factor(P.T @ S @ P)
            two = lambda r: B(r) + P @ sla.cho_solve(fc, P.T @ r)
            print(f"   two-level, aggregates of {m} cameras (coarse dimension {6 * nagg}):", {tol: pcg(S, rhs, two, tol) for tol in tols}, flush=True)


if __name__ == "__main__":
    main()
