"""Pins the oracle (and the CUDA path) to REAL libceres trajectories, when someone has produced them.

Ceres is absent from the build container, so parity is "unpinned" until tools/ceres_crosscheck.cpp has
been run on a box that has it and its JSON output committed as tests/golden/ceres/<name>.json.  With no such file these
tests skip and say so; with one they enforce the north-star gates against Ceres itself: same iteration count, every
per-iteration cost within 1e-9 relative, final poses / points within 1e-6.
"""
import glob
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN as GOLDEN_DIR, load_golden

CERES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "ceres", "*.json")))
LOSS = {"none": 0, "huber": 1, "cauchy": 2}


def _load(path):
    with open(path) as f:
        c = json.load(f)
    name = os.path.splitext(os.path.basename(path))[0]
    prob, gold = load_golden(name)
    assert LOSS[c["loss"]] == int(gold["loss"]), "cross-check was run with a different loss than the fixture"
    its = c["iterations"]
    return prob, int(gold["loss"]), c, np.array([i["cost"] for i in its]), np.array([i["trust_region_radius"] for i in its])


def _check(c, cost_ceres, s, cam, pt):
    # Ceres lists iteration 0 (the initial evaluation) plus one entry per trust-region iteration
    assert s["n_iters"] == len(cost_ceres) - 1
    got = np.asarray(s["cost"][: len(cost_ceres)])
    assert np.max(np.abs(got - cost_ceres) / np.abs(cost_ceres)) < 1e-9
    assert np.allclose(cam.ravel(), np.asarray(c["cam"]), rtol=1e-6, atol=1e-9)
    assert np.allclose(pt.ravel(), np.asarray(c["pt"]), rtol=1e-6, atol=1e-8)


def test_ceres_outputs_present_or_unpinned():
    if not CERES:
        pytest.skip("parity UNPINNED: no tests/golden/ceres/*.json (run tools/ceres_crosscheck.cpp where Ceres is installed)")


@pytest.mark.parametrize("path", CERES or [None])
def test_oracle_matches_ceres(oracle, path):
    if path is None:
        pytest.skip("no Ceres trajectories committed")
    prob, loss, c, cost_ceres, _ = _load(path)
    ref, s = oracle.solve(prob, oracle.options(loss=loss))
    _check(c, cost_ceres, s, ref.cam, ref.pt)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CERES or [None])
def test_gpu_matches_ceres(ctx, path):
    if path is None:
        pytest.skip("no Ceres trajectories committed")
    import gl_slam_b200 as g
    prob, loss, c, cost_ceres, _ = _load(path)
    got, s = ctx.solve(prob, g.options(loss=loss))
    _check(c, cost_ceres, s, got.cam, got.pt)
