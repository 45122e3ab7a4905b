"""Multi-rank parity check (launch with torchrun, one rank per GPU): a sharded solve must reproduce the single-rank
oracle trajectory; cameras identical on every rank; each rank returns its own points; the gradient norms every rank
reports must agree with each other and with the single-GPU solve (a rank-local norm would let ranks take different
gradient-tolerance decisions and hang the next collective).

`run_cases(ctx, rank, world, local, dev)` is also what bench.py runs before its timed region at N > 1 (the
`sharded_parity` key of its line), so the driver's multi-GPU runs carry the evidence."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gl_slam_b200 as g  # noqa: E402
from gl_slam_b200 import _abi, scene  # noqa: E402


def cases():
    return {
        "window_dense": (dict(n_cam=10, n_pt=2000, track_len=4, seed=2, outlier_frac=0.05, rot_sigma=0.005, pos_sigma=0.03), {}),
        "map_pcg": (dict(n_cam=24, n_pt=3000, track_len=lambda rng, n: 3 + rng.poisson(3.0, size=n), seed=41, rot_sigma=0.003, pos_sigma=0.03), {}),
        "huber": (dict(n_cam=16, n_pt=1500, track_len=4, seed=8, outlier_frac=0.1, rot_sigma=0.004, pos_sigma=0.03), dict(loss=1)),
        # a street grid: every shard boundary cuts through revisited streets, many cameras are seen by two ranks
        "street_grid_pcg": ("street", dict(n_rows=6, n_cols=12, n_pt=6000, track_len=lambda rng, n: 2 + rng.poisson(3.0, size=n), seed=4,
                                           rot_sigma=0.002, pos_sigma=0.03), dict(max_iters=6)),
        # the archived g2o formulation, sharded: lambda0 needs the max Hessian diagonal over ALL ranks' points
        "g2o_pcg": (dict(n_cam=24, n_pt=3000, track_len=lambda rng, n: 3 + rng.poisson(3.0, size=n), seed=43, rot_sigma=0.003, pos_sigma=0.03),
                    dict(mode=_abi.MODE_G2O, loss=1, max_iters=8)),
        "g2o_dense_info": (dict(n_cam=10, n_pt=2000, track_len=4, seed=5, outlier_frac=0.05, rot_sigma=0.005, pos_sigma=0.03),
                           dict(mode=_abi.MODE_G2O, loss=1, loss_scale=3.0, max_iters=8)),
    }


def run_cases(ctx, rank, world, local, dev):
    """Returns (ok, per-case dict) on rank 0 ((ok, {}) elsewhere).  Collective: every rank must call it."""
    results = {}
    ok = True
    single = g.Context(device=local) if rank == 0 else None
    for name, spec in cases().items():
        if spec[0] == "street":
            prob, okw = scene.make_street_grid(**spec[1]), spec[2]
        else:
            prob, okw = scene.make_scene(**spec[0]), spec[1]
        if okw.get("mode") == _abi.MODE_G2O:
            prob = scene.as_g2o(prob)
        if name.endswith("_info"):
            prob.pt_info = 1.0 / np.maximum(prob.pt[:, 2], 0.1) ** 2
        sub, idx = scene.shard_by_point(prob, world, rank)
        got, s = ctx.solve(sub, g.options(**okw))
        # cameras must be bit-identical across ranks (all-reduced sums are)
        cam_t = torch.from_numpy(got.cam).to(dev)
        cam0 = cam_t.clone()
        dist.broadcast(cam0, 0)
        same = bool(torch.equal(cam_t, cam0))
        # ... and so must every number the trust-region decisions are taken on
        trace = torch.tensor(list(s["cost"]) + list(s["gradient_max_norm"]) + list(s["radius"]) + [s["n_iters"], s["termination"], s["stop_reason"]],
                             dtype=torch.float64, device=dev)
        trace0 = trace.clone()
        dist.broadcast(trace0, 0)
        same_trace = bool(torch.equal(trace, trace0))
        pts = torch.zeros(prob.n_pt, 3, dtype=torch.float64, device=dev)
        pts[torch.from_numpy(idx).to(dev)] = torch.from_numpy(got.pt).to(dev)
        dist.all_reduce(pts)
        if rank == 0:
            from oracle import oracle
            ref, so = oracle.solve(prob, oracle.options(**okw))
            one, s1 = single.solve(prob, g.options(**okw))
            n = min(len(s["cost"]), len(so["cost"]))
            rel = max(abs(a - b) / abs(b) for a, b in zip(s["cost"][:n], so["cost"][:n]))
            g1 = np.array(s1["gradient_max_norm"][:s1["n_iters"] + 1]); gs = np.array(s["gradient_max_norm"][:s["n_iters"] + 1])
            grel = float(np.max(np.abs(gs - g1) / np.maximum(np.abs(g1), 1e-300))) if gs.shape == g1.shape else float("inf")
            centres = scene.to_camera_to_world(prob.cam)[:, 3:6] if okw.get("mode") == _abi.MODE_G2O else prob.cam[:, 3:6]
            d = np.linalg.norm(ref.pt[prob.obs_pt] - centres[prob.obs_cam], axis=1)
            far = np.zeros(prob.n_pt, bool); np.logical_or.at(far, prob.obs_pt, d > 150.0)
            perr = float((np.linalg.norm(pts.cpu().numpy() - ref.pt, axis=1) / np.maximum(np.linalg.norm(ref.pt, axis=1), 1.0))[~far].max())
            case_ok = bool(s["n_iters"] == so["n_iters"] and rel < 1e-9 and np.allclose(got.cam, ref.cam, rtol=1e-6, atol=1e-8) and perr < 1e-6
                           and grel < 1e-6 and s["n_iters"] == s1["n_iters"])
            results[name] = dict(ok=case_ok, iters=[s["n_iters"], so["n_iters"]], cost_rel=rel, cam_err=float(np.abs(got.cam - ref.cam).max()), pt_err=perr,
                                 cams_identical=same, gradient_norm_rel_vs_single_gpu=grel, cg=max(s["cg_iters"]))
            ok &= case_ok
        flag = torch.tensor([1 if (same and same_trace) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok &= bool(flag.item())
        if rank == 0:
            results[name]["identical_on_all_ranks"] = bool(flag.item())
    if single is not None:
        single.close()
    okt = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(okt, 0)
    return bool(okt.item()), results


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idt = torch.zeros(_abi.GLBA_NCCL_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(g.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    ctx = g.Context(device=local, rank=rank, world=world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
    ok, results = run_cases(ctx, rank, world, local, dev)
    if rank == 0:
        print(json.dumps({"world": world, "ok": bool(ok), "cases": results}))
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
