"""CPU tests of the multi-GPU host logic: point-track partitioning + the camera-block exchange.

world_size-2 gloo run: every rank linearises ITS shard (whole tracks, all cameras) with the oracle, the additive
camera quantities are all-reduced exactly as the CUDA path all-reduces them over NCCL, and must equal the
single-rank linearisation of the whole map; per-point quantities must simply partition.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gl_slam_b200 import scene


def test_shard_by_point_partitions_whole_tracks():
    prob = scene.make_scene(20, 3000, lambda rng, n: 2 + rng.poisson(3.0, size=n), seed=5)
    for world in (2, 3, 8):
        seen_pts, n_obs = [], 0
        sizes = []
        for r in range(world):
            sub, idx = scene.shard_by_point(prob, world, r)
            assert np.array_equal(sub.cam, prob.cam) and np.array_equal(sub.cam_fixed, prob.cam_fixed)   # cameras replicated
            assert np.array_equal(sub.pt, prob.pt[idx])
            # whole tracks: every observation of a shard point is in the shard, none of any other point
            want = np.isin(prob.obs_pt, idx)
            assert sub.n_obs == want.sum()
            assert np.array_equal(idx[sub.obs_pt], prob.obs_pt[want]) and np.array_equal(sub.obs_cam, prob.obs_cam[want])
            assert np.all(np.diff(sub.obs_pt) >= 0)           # still track-contiguous: no device sort needed
            seen_pts.append(idx); n_obs += sub.n_obs; sizes.append(sub.n_obs)
        allpts = np.concatenate(seen_pts)
        assert np.array_equal(np.sort(allpts), np.arange(prob.n_pt)) and n_obs == prob.n_obs
        assert max(sizes) - min(sizes) <= 2 * 20              # balanced by observation count (within a track or two)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle
    prob = scene.make_scene(12, 800, lambda rng, n: 2 + rng.poisson(2.0, size=n), seed=9, outlier_frac=0.05)
    sub, idx = scene.shard_by_point(prob, world, rank)
    L = oracle.linearize(sub, 1e4, per_obs=False)
    # the exchange step: cost + per-camera gradient and Hessian blocks (27 independent numbers per camera)
    buf = torch.from_numpy(np.concatenate([[L.cost], L.grad_cam.ravel(), L.hess_cam.ravel()]))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    gmax = torch.tensor([np.abs(L.grad_pt).max()])
    dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        full = oracle.linearize(prob, 1e4, per_obs=False)
        ref = np.concatenate([[full.cost], full.grad_cam.ravel(), full.hess_cam.ravel()])
        out["cam_err"] = float(np.abs(buf.numpy() - ref).max() / np.abs(ref).max())
        out["gmax_err"] = float(abs(gmax.item() - np.abs(full.grad_pt).max()))
        out["pt_err"] = float(np.abs(L.grad_pt - full.grad_pt[idx]).max() + np.abs(L.hess_pt - full.hess_pt[idx]).max())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_exchange_reproduces_single_rank_linearisation():
    from oracle import oracle
    oracle.build()
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        assert out["cam_err"] < 1e-13, out["cam_err"]
        assert out["gmax_err"] == 0.0
        assert out["pt_err"] == 0.0      # per-point blocks are local to the rank that owns the track
