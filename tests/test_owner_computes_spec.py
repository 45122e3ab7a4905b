"""Executable specification of the round-2 multi-GPU exchange (DESIGN §6 "owner-computes cameras") — numpy only.

Today every rank keeps every camera-sized vector and all-reduces all cameras' partial sums.  The planned scheme:
  * a rank is ACTIVE on the cameras its own tracks observe; cameras seen by one rank never leave it;
  * cameras seen by several ranks (SHARED) are packed into a short buffer and sum-reduced; nothing else is exchanged;
  * every scalar sum counts a shared camera once: on its OWNER, the lowest rank that observes it;
  * the PCG is the single-reduction (Chronopoulos-Gear) variant, so ONE collective per iteration carries
    [ w = S u on the shared cameras | gamma = r.u | delta = u.w ], using that u.w is linear in the per-rank partial products.
This file simulates R ranks in one process, with the per-rank partial blocks taken from the oracle's linearisation of each
shard, and checks that the scheme reproduces the replicated solve: same reduced system on every active camera, same PCG
solution to 1e-12, and counts the doubles that cross the "network" in both schemes."""
import numpy as np
import pytest

from gl_slam_b200 import scene


def shard_blocks(oracle, prob, world):
    """Per rank: observed-camera mask and the local pieces of the reduced system in dense form (small problems only)."""
    out = []
    for r in range(world):
        sub, _ = scene.shard_by_point(prob, world, r)
        L = oracle.linearize(sub, 1e4, per_obs=True)
        n = 6 * prob.n_cam
        Jc, Jp, res = L.jac_cam, L.jac_pt, L.residuals
        # local contribution to S and rhs:  B^r - sum_j W_j C_j^-1 W_j',  b^r - sum_j W_j C_j^-1 g_j  (points are local: exact)
        B = np.zeros((n, n)); b = np.zeros(n)
        C = np.zeros((sub.n_pt, 3, 3)); g = np.zeros((sub.n_pt, 3))
        W = {}
        for k in range(sub.n_obs):
            i, j = sub.obs_cam[k], sub.obs_pt[k]
            B[6 * i:6 * i + 6, 6 * i:6 * i + 6] += Jc[k].T @ Jc[k]
            b[6 * i:6 * i + 6] += Jc[k].T @ res[k]
            C[j] += Jp[k].T @ Jp[k]; g[j] += Jp[k].T @ res[k]
            W[(i, j)] = W.get((i, j), 0) + Jc[k].T @ Jp[k]
        lam_p = 1e-4                                            # any fixed damping: the spec is about the exchange, not LM
        S, rhs = B.copy(), b.copy()
        for j in range(sub.n_pt):
            Ci = np.linalg.inv(C[j] + lam_p * np.eye(3))
            cams = [i for (i, jj) in W if jj == j]
            for i in cams:
                rhs[6 * i:6 * i + 6] -= W[(i, j)] @ Ci @ g[j]
                for k2 in cams:
                    S[6 * i:6 * i + 6, 6 * k2:6 * k2 + 6] -= W[(i, j)] @ Ci @ W[(k2, j)].T
        touched = np.zeros(prob.n_cam, bool); touched[sub.obs_cam] = True
        out.append(dict(S=S, rhs=rhs, touched=touched))
    return out


def pcg_replicated(S, b, Minv, tol, max_it):
    x = np.zeros_like(b); r = b.copy(); z = Minv @ r; p = z.copy(); rz = r @ z; rz0 = rz
    for it in range(max_it):
        q = S @ p; a = rz / (p @ q); x += a * p; r -= a * q; z = Minv @ r; rz1 = r @ z
        if np.sqrt(rz1) <= tol * np.sqrt(rz0):
            return x, it + 1
        p = z + (rz1 / rz) * p; rz = rz1
    return x, max_it


@pytest.mark.parametrize("world", [2, 3])
def test_owner_computes_exchange_reproduces_replicated_solve(oracle, world):
    prob = scene.make_scene(n_cam=30, n_pt=450, track_len=lambda rng, n: 2 + rng.poisson(1.5, size=n), seed=17, rot_sigma=0.003,
                            pos_sigma=0.02)
    free = ~prob.cam_fixed.astype(bool)
    ranks = shard_blocks(oracle, prob, world)
    n = 6 * prob.n_cam
    fmask = np.repeat(free, 6)
    lam_c = 1e-3
    # ---- replicated scheme (today): all-reduce everything, every rank solves the whole camera system
    S_all = sum(r["S"] for r in ranks) + lam_c * np.eye(n)
    b_all = sum(r["rhs"] for r in ranks)
    Sf, bf = S_all[np.ix_(fmask, fmask)], b_all[fmask]
    Minv = np.zeros_like(Sf)
    for i in range(Sf.shape[0] // 6):
        Minv[6 * i:6 * i + 6, 6 * i:6 * i + 6] = np.linalg.inv(Sf[6 * i:6 * i + 6, 6 * i:6 * i + 6])
    x_ref, it_ref = pcg_replicated(Sf, bf, Minv, 1e-13, 500)
    assert np.allclose(Sf @ x_ref, bf, rtol=0, atol=1e-9 * np.abs(bf).max())
    traffic_replicated = world * (54 * prob.n_cam) + it_ref * world * 6 * prob.n_cam       # doubles contributed to all-reduces

    # ---- owner-computes scheme
    touch = np.stack([r["touched"] for r in ranks])                   # [world, n_cam]
    nranks = touch.sum(axis=0)
    shared = nranks >= 2
    owner = np.where(nranks > 0, np.argmax(touch, axis=0), -1)        # lowest observing rank
    assert shared.sum() < prob.n_cam                                   # the scene is banded: most cameras are private
    act = [touch[r] & free for r in range(world)]
    wgt = [act[r] & (owner == r) for r in range(world)]              # counts a camera once in scalar sums
    assert np.array_equal(np.sum(wgt, axis=0).astype(bool), free & (nranks > 0))
    sh6 = np.repeat(shared, 6)

    def exchange(vecs):                                                # sum-reduce of the shared cameras' entries only
        tot = sum(v[sh6] for v in vecs)
        for v in vecs:
            v[sh6] = tot
        return vecs

    # finalised diagonal blocks / rhs: private cameras are already complete locally, shared ones after the exchange
    Sdiag = [np.stack([r["S"][6 * i:6 * i + 6, 6 * i:6 * i + 6] for i in range(prob.n_cam)]).reshape(-1) for r in ranks]
    sh36 = np.repeat(shared, 36)
    tot = sum(d[sh36] for d in Sdiag)
    for d in Sdiag:
        d[sh36] = tot
    rhs = exchange([r["rhs"].copy() for r in ranks])
    for r in range(world):
        for i in np.nonzero(act[r])[0]:
            blk = Sdiag[r][36 * i:36 * i + 36].reshape(6, 6) + lam_c * np.eye(6)
            assert np.allclose(blk, S_all[6 * i:6 * i + 6, 6 * i:6 * i + 6], rtol=1e-13, atol=1e-9)
            assert np.allclose(rhs[r][6 * i:6 * i + 6], b_all[6 * i:6 * i + 6], rtol=1e-12, atol=1e-9)
    Mi = []
    for r in range(world):
        M = np.zeros((n, n))
        for i in np.nonzero(act[r])[0]:
            M[6 * i:6 * i + 6, 6 * i:6 * i + 6] = np.linalg.inv(Sdiag[r][36 * i:36 * i + 36].reshape(6, 6) + lam_c * np.eye(6))
        Mi.append(M)
    a6 = [np.repeat(a, 6) for a in act]
    w6 = [np.repeat(w, 6) for w in wgt]

    def local_matvec(r, u):
        """S^r u restricted to rank r's active cameras; the damping term is added once, by the owner."""
        y = ranks[r]["S"] @ (u * a6[r])
        y += lam_c * u * w6[r]
        return y * a6[r]

    # Chronopoulos-Gear PCG, one exchange per iteration: [w on shared cameras | gamma | delta]
    x = [np.zeros(n) for _ in range(world)]
    rr = [rhs[r] * a6[r] for r in range(world)]
    u = [Mi[r] @ rr[r] for r in range(world)]
    w = exchange([local_matvec(r, u[r]) for r in range(world)])
    # NB: after the exchange a shared camera's w holds the full sum on every rank that is active on it
    gamma = sum((rr[r] * w6[r]) @ u[r] for r in range(world))
    # delta = u.w: linear in the per-rank partial products, so it can be formed BEFORE the exchange completes
    delta = sum(u[r] @ local_matvec(r, u[r]) for r in range(world))
    gamma0 = gamma
    alpha, beta = gamma / delta, 0.0
    p = [np.zeros(n) for _ in range(world)]
    s = [np.zeros(n) for _ in range(world)]
    collectives, doubles = 1, 0
    for it in range(500):
        for r in range(world):
            p[r] = u[r] + beta * p[r]; s[r] = w[r] + beta * s[r]
            x[r] += alpha * p[r]; rr[r] -= alpha * s[r]
            u[r] = Mi[r] @ rr[r]
        part = [local_matvec(r, u[r]) for r in range(world)]
        g_new = sum((rr[r] * w6[r]) @ u[r] for r in range(world))         # rides in the same message
        delta = sum(u[r] @ part[r] for r in range(world))                 # idem (partial products)
        w = exchange(part)
        collectives += 1; doubles += world * (6 * int(shared.sum()) + 2)
        if np.sqrt(g_new) <= 1e-13 * np.sqrt(gamma0):
            break
        beta = g_new / gamma
        alpha = g_new / (delta - beta * g_new / alpha)
        gamma = g_new
    # every rank holds the solution on its active cameras; together they cover every free, observed camera exactly
    x_oc = np.zeros(n)
    x_oc_ref = x_ref_full(x_ref, fmask, n)
    for r in range(world):
        x_oc[w6[r]] = x[r][w6[r]]
        sel = a6[r] & sh6
        assert np.allclose(x[r][sel], x_oc_ref[sel], rtol=1e-6, atol=1e-9 * np.abs(x_ref).max())     # shared cameras agree on every rank
    # both iterations stop at the same preconditioned-residual tolerance; their solutions then agree to cond(S) x 1e-13
    assert np.allclose(x_oc[fmask], x_ref, rtol=1e-6, atol=1e-9 * np.abs(x_ref).max())
    assert np.abs(Sf @ x_oc[fmask] - bf).max() <= 1e-9 * np.abs(bf).max()
    assert abs((it + 1) - it_ref) <= 3                                   # same Krylov space, one reduction per iteration
    traffic_owner = world * 54 * int(shared.sum()) + doubles
    assert traffic_owner < 0.6 * traffic_replicated                      # 30 cameras; the ratio shrinks with the map (C4 x 8: ~1 %)


def x_ref_full(x_ref, fmask, n):
    full = np.zeros(n)
    full[fmask] = x_ref
    return full
