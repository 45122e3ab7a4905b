"""Executable numpy specification of the assembled reduced camera matrix (gl_slam_b200/csrc/glba_sparse.cuh), CPU only.

What the CUDA path does, restated on the oracle's per-observation Jacobians:
  * an *instance* is an unordered pair of observations of one point whose cameras are both free; the off-diagonal block
    S_ab (a < b) is minus the sum over its instances of  W_a C_j^-1 W_b'  with  W = Jc' Jp  (6x3);
  * the diagonal block is B_a + D_a - sum over the camera's own observations;
  * a row list over both triangles (lower-triangle entries read the stored block transposed) gives S x.
The sums must reproduce the dense Schur complement of the same linearisation.  Also specified here: the split of the rows
over the CTAs of the one-row-per-warp PCG kernel (contiguous ranges, bounded rows, smallest feasible bound on the entries).
The GPU tests (tests/test_gpu_scale.py::test_explicit_*) hold the real kernels to the oracle and to the matrix-free product."""
import numpy as np
import pytest

from gl_slam_b200 import scene


def _blocks(prob, L, radius):
    n, n_cam, n_pt = prob.n_obs, prob.n_cam, prob.n_pt
    Jc = L.jac_cam.reshape(n, 2, 6); Jp = L.jac_pt.reshape(n, 2, 3)
    free = prob.cam_fixed == 0
    B = np.zeros((n_cam, 6, 6)); np.add.at(B, prob.obs_cam, np.einsum('nij,nik->njk', Jc, Jc))
    C = np.zeros((n_pt, 3, 3)); np.add.at(C, prob.obs_pt, np.einsum('nij,nik->njk', Jp, Jp))
    B += np.einsum('ni,ij->nij', np.einsum('nii->ni', B) / radius, np.eye(6))
    C += np.einsum('ni,ij->nij', np.einsum('nii->ni', C) / radius, np.eye(3))
    Ci = np.linalg.inv(C)
    W = np.einsum('nij,nik->njk', Jc, Jp)
    return free, B, Ci, W


def _dense_schur(prob, free, B, Ci, W):
    idx = np.where(free)[0]; slot = -np.ones(prob.n_cam, int); slot[idx] = np.arange(len(idx))
    S = np.zeros((6 * len(idx), 6 * len(idx)))
    for a in idx:
        S[6 * slot[a]:6 * slot[a] + 6, 6 * slot[a]:6 * slot[a] + 6] = B[a]
    order = np.argsort(prob.obs_pt, kind='stable'); starts = np.searchsorted(prob.obs_pt[order], np.arange(prob.n_pt + 1))
    for j in range(prob.n_pt):
        ks = order[starts[j]:starts[j + 1]]
        for oa in ks:
            a = slot[prob.obs_cam[oa]]
            if a < 0:
                continue
            for ob in ks:
                b = slot[prob.obs_cam[ob]]
                if b >= 0:
                    S[6 * a:6 * a + 6, 6 * b:6 * b + 6] -= W[oa] @ Ci[j] @ W[ob].T
    return S, slot


def test_instances_reproduce_the_dense_schur_complement():
    from oracle import oracle
    prob = scene.make_street_grid(3, 8, 400, track_len=lambda rng, n: 2 + rng.poisson(2.0, size=n), seed=5, rot_sigma=0.002, pos_sigma=0.03)
    radius = 1e4
    L = oracle.linearize(prob, radius, per_obs=True)
    free, B, Ci, W = _blocks(prob, L, radius)
    S, slot = _dense_schur(prob, free, B, Ci, W)
    # --- the structure the device builds: instances sorted by the block key a * n_cam + b (a < b), point order within a block
    order = np.argsort(prob.obs_pt, kind='stable'); starts = np.searchsorted(prob.obs_pt[order], np.arange(prob.n_pt + 1))
    keys, inst = [], []
    for j in range(prob.n_pt):
        ks = [o for o in order[starts[j]:starts[j + 1]] if free[prob.obs_cam[o]]]
        for x in range(len(ks)):
            for y in range(x + 1, len(ks)):
                oa, ob = ks[x], ks[y]
                a, b = prob.obs_cam[oa], prob.obs_cam[ob]
                assert a != b                      # (a point seen twice by one camera keeps the matrix-free product)
                if a > b:
                    oa, ob, a, b = ob, oa, b, a
                keys.append(a * prob.n_cam + b); inst.append((oa, ob, j))
    keys = np.array(keys); perm = np.argsort(keys, kind='stable')
    ukey, first, count = np.unique(keys[perm], return_index=True, return_counts=True)
    blocks = np.zeros((len(ukey), 6, 6))
    for p, (f, c) in enumerate(zip(first, count)):
        for i in perm[f:f + c]:
            oa, ob, j = inst[i]
            blocks[p] -= W[oa] @ Ci[j] @ W[ob].T
    # diagonal: B_a + D_a minus the camera's own observations
    diag = B.copy()
    np.subtract.at(diag, prob.obs_cam, np.einsum('nij,njk,nlk->nil', W, Ci[prob.obs_pt], W))
    # --- row lists over both triangles; product against the dense matrix
    rows = [[] for _ in range(prob.n_cam)]
    for p, k in enumerate(ukey):
        a, b = divmod(int(k), prob.n_cam)
        rows[a].append((b, p, False)); rows[b].append((a, p, True))
    rng = np.random.default_rng(0)
    x = np.zeros((prob.n_cam, 6)); x[free] = rng.standard_normal((free.sum(), 6))
    y = np.zeros_like(x)
    for a in np.where(free)[0]:
        y[a] = diag[a] @ x[a]
        for b, p, tr in sorted(rows[a]):
            y[a] += (blocks[p].T if tr else blocks[p]) @ x[b]
    ref = (S @ x[free].ravel()).reshape(-1, 6)
    assert np.abs(y[free] - ref).max() <= 1e-11 * np.abs(ref).max()
    # every non-zero off-diagonal block of the dense matrix is one of the assembled blocks, and nothing else is
    nz = {(a, b) for a in np.where(free)[0] for b in np.where(free)[0] if a < b and np.abs(S[6 * slot[a]:6 * slot[a] + 6, 6 * slot[b]:6 * slot[b] + 6]).max() > 0}
    assert nz == {divmod(int(k), prob.n_cam) for k in ukey}


def _split(row_len, rows_per, max_ctas):
    """glba.cu::ensure_explicit: smallest bound on the entries per CTA for which a greedy split into contiguous ranges of at most
    rows_per rows needs at most max_ctas CTAs (bisection)."""
    def greedy(bound):
        out, r = [0], 0
        while r < len(row_len):
            cnt = ent = 0
            while r < len(row_len) and cnt < rows_per and (cnt == 0 or ent + row_len[r] <= bound):
                ent += row_len[r]; cnt += 1; r += 1
            out.append(r)
        return out
    lo, hi = max(1, int(max(row_len))), max(1, int(sum(row_len)))
    while lo < hi:
        mid = (lo + hi) // 2
        if len(greedy(mid)) - 1 <= max_ctas:
            hi = mid
        else:
            lo = mid + 1
    return lo, greedy(lo)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_balanced_row_ranges(seed):
    rng = np.random.default_rng(seed)
    n = 1800
    row_len = np.where(rng.random(n) < 0.3, rng.integers(60, 80, n), rng.integers(20, 34, n))      # revisited streets: 2.5x the entries
    row_len[:2] = 0                                                                                 # fixed cameras: empty rows
    bound, cr = _split(row_len, 16, 148)
    assert cr[0] == 0 and cr[-1] == n and all(b > a for a, b in zip(cr[:-1], cr[1:])) and len(cr) - 1 <= 148
    ent = [int(row_len[a:b].sum()) for a, b in zip(cr[:-1], cr[1:])]
    assert max(b - a for a, b in zip(cr[:-1], cr[1:])) <= 16 and max(ent) <= bound
    # 16 consecutive rows per CTA (the first version) is far less even
    naive = max(int(row_len[a:a + 16].sum()) for a in range(0, n, 16))
    assert bound <= 780 and bound < 0.75 * naive, (bound, naive)
    # no smaller bound is feasible
    if bound > int(row_len.max()):
        lo2, cr2 = bound - 1, None
        r = cnt_ctas = 0
        while r < n:
            cnt = e = 0
            while r < n and cnt < 16 and (cnt == 0 or e + row_len[r] <= lo2):
                e += row_len[r]; cnt += 1; r += 1
            cnt_ctas += 1
        assert cnt_ctas > 148
