#!/usr/bin/env python
"""BASELINE.json configs[4] ("C5"): project + resample + splat on N GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/run_c5.py [--points 2000000] [--views 16] [--size 1024] [--check]

Pipeline per outer iteration:
  1. every rank projects + resamples ITS contiguous shard of the cloud (ShardedUniformProjection:
     projection without communication, resample with one all-gather of xyz+normal per iteration);
  2. one all-gather of the resulting iso-points -> the replicated point set every rank splats;
  3. every rank rasterises ITS views (synthetic orthographic cameras at the kernel boundary: the EWA
     parameter math of the reference stays PyTorch and is not part of this path), blends RGBA and
     back-propagates synthetic occupancy / depth gradients to the shared 3-D points;
  4. one all-reduce (sum) of the per-point gradients.
Rank 0 prints one JSON line with device-timed stage durations (max over ranks).  --check also runs
the splat stage for ALL views on every rank and verifies that the all-reduced gradient equals it.
"""
import argparse
import json
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isopoints_b200 import splat  # noqa: E402
from isopoints_b200.dist import (ShardedUniformProjection, all_gather_varlen, all_reduce_point_grads,  # noqa: E402
                                 shard_range, shard_views)
from tests.helpers import Siren, SirenSDF, SphereSDF  # noqa: E402


def view_rotation(v, n_views, dev):
    a = 2 * math.pi * v / n_views
    b = 0.35 * math.sin(3 * a)
    ca, sa, cb, sb = math.cos(a), math.sin(a), math.cos(b), math.sin(b)
    ry = torch.tensor([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]], device=dev)
    rx = torch.tensor([[1, 0, 0], [0, cb, -sb], [0, sb, cb]], device=dev)
    return rx @ ry


def splat_views(pts_world, views, n_views, S, K, occ_grads, z_grads, rgb):
    """Rasterise + blend + backward for `views`; returns (rgba list, grad wrt pts_world)."""
    dev = pts_world.device
    P = pts_world.shape[0]
    p = pts_world.detach().requires_grad_(True)
    scr = []
    for v in views:
        q = p @ view_rotation(v, n_views, dev).T
        scr.append(torch.stack([q[:, 0] * 0.45, q[:, 1] * 0.45, q[:, 2] + 3.0], dim=1))
    scr = torch.cat(scr, 0)
    nv = len(views)
    sig = 1.5 * 2.0 / S
    ell = torch.tensor([1 / sig ** 2, 0.0, 1 / sig ** 2], device=dev).expand(nv * P, 3).contiguous()
    radii = torch.full((nv * P, 2), sig, device=dev)
    cutoff = torch.ones(1, device=dev)
    first = torch.arange(nv, device=dev, dtype=torch.int64) * P
    num = torch.full((nv,), P, device=dev, dtype=torch.int64)
    idx, zbuf, qv, occ = splat.EllipticalRasterizer.apply(scr, ell, cutoff.expand(nv * P), radii, first, num, 0.05, S, K,
                                                          64 if S > 512 else 32, 0, 10.0)
    rgba = splat.blend_rgba(idx, qv, occ, None, rgb.repeat(nv, 1))
    og = torch.stack([occ_grads[v] for v in views])
    zg = torch.stack([z_grads[v] for v in views])
    ((occ * og).sum() + (zbuf * zg).sum()).backward()
    return rgba, p.grad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=2_000_000)
    ap.add_argument("--views", type=int, default=16)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--sdf", default="siren", choices=["siren", "opaque", "sphere"],
                    help="siren: reference decoder structure (fused SDF kernel); opaque: same weights through autograd")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.backends.cuda.matmul.allow_tf32 = False
    S, K, V = args.size, 8, args.views

    g = torch.Generator().manual_seed(7)
    b, e = shard_range(args.points, rank, world)
    # every rank draws the same stream and keeps its slice: identical to a single-rank run
    x = ((torch.rand(args.points, 3, generator=g) - 0.5) * 2)[b:e].to(dev)[None]
    net = {"siren": lambda: Siren(256, 7, 30.0, seed=0), "opaque": lambda: SirenSDF(seed=0),
           "sphere": SphereSDF}[args.sdf]().to(dev)
    proj = ShardedUniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    gg = torch.Generator().manual_seed(9)
    occ_grads = [(torch.randn(S, S, generator=gg) * (torch.rand(S, S, generator=gg) < 0.1)).to(dev) for _ in range(V)]
    z_grads = [torch.randn(S, S, K, generator=gg).to(dev) for _ in range(V)]

    def ev():
        t = torch.cuda.Event(enable_timing=True)
        t.record()
        return t

    stages = {}
    for it in range(2):                   # iteration 0 warms up (allocator, NCCL channels)
        dist.barrier(); torch.cuda.synchronize()
        t0 = ev()
        out = proj.project_points(x, net, skip_upsampling=True)
        t1 = ev()
        loc = out["levelset_points"][0][out["mask"][0]]
        iso, counts = all_gather_varlen(loc)
        t2 = ev()
        rgb = (iso * 0.5 + 0.5).clamp(0, 1).contiguous()
        rgba, grad = splat_views(iso, shard_views(V, rank, world), V, S, K, occ_grads, z_grads, rgb)
        t3 = ev()
        all_reduce_point_grads(grad)
        t4 = ev()
        torch.cuda.synchronize()
        stages = {"project_resample_ms": t0.elapsed_time(t1), "allgather_points_ms": t1.elapsed_time(t2),
                  "splat_fwd_blend_bwd_ms": t2.elapsed_time(t3), "allreduce_grads_ms": t3.elapsed_time(t4),
                  "total_ms": t0.elapsed_time(t4)}
    red = torch.tensor([stages[k] for k in sorted(stages)], dtype=torch.float64, device=dev)
    dist.all_reduce(red, op=dist.ReduceOp.MAX)
    stages = dict(zip(sorted(stages), red.tolist()))
    ok = None
    if args.check:
        _, full = splat_views(iso, list(range(V)), V, S, K, occ_grads, z_grads, rgb)
        scale = full.abs().amax(0).clamp_min(1e-20)
        ok = bool(torch.allclose(grad / scale, full / scale, rtol=1e-4, atol=2e-5))
    if rank == 0:
        n_iso = int(iso.shape[0])
        print(json.dumps({"config": "C5: %d points, %s SDF, project+resample+splat %dx%d^2, %d GPUs"
                                    % (args.points, args.sdf, V, S, world), "n_gpus": world, "iso_points": n_iso,
                          "iso_points_per_s": args.points / (stages["total_ms"] * 1e-3), **stages,
                          "allgather_bytes": n_iso * 12, "allreduce_bytes": n_iso * 12,
                          "grad_check_vs_all_views_on_one_rank": ok}), flush=True)
    dist.destroy_process_group()
    if ok is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
