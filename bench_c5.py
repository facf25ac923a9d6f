#!/usr/bin/env python
"""BASELINE.json configs[4] ("C5"): 2 000 000 points, project + resample + splat 16 x 1024^2 on N GPUs of one box.

Library use (bench.py adds the returned record to its JSON line under "c5" at every N, so the driver's 1/2/4/8
scaling run carries the STRONG scaling of this fixed 2 M-point problem):

    import bench_c5; rec = bench_c5.run(rank, world, dev, steps=2, check=True)

Stand-alone (one JSON line from rank 0):

    python bench_c5.py                                                     # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 bench_c5.py [--points 2000000] [--views 16] [--size 1024] [--check]

Pipeline per outer iteration:
  1. every rank projects + resamples ITS contiguous shard of the cloud (ShardedUniformProjection: projection
     without communication, resample with one all-gather of xyz + normal per sample iteration);
  2. one all-gather of the resulting iso-points -> the replicated point set every rank splats;
  3. every rank rasterises ITS views (synthetic cameras at the kernel boundary: per-view rotation, orthographic
     scale, splats of sigma = 1.5 px), blends RGBA and back-propagates synthetic occupancy / depth gradients to
     the shared 3-D points;
  4. one all-reduce (sum) of the per-point gradients.
Stage durations are CUDA-event times, max over ranks.  `check`: the splat stage is also run for ALL views on every
rank and the all-reduced gradient must equal it (rtol 1e-4).
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def view_rotation(v, n_views, dev):
    a = 2 * math.pi * v / n_views
    b = 0.35 * math.sin(3 * a)
    ca, sa, cb, sb = math.cos(a), math.sin(a), math.cos(b), math.sin(b)
    ry = torch.tensor([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]], device=dev)
    rx = torch.tensor([[1, 0, 0], [0, cb, -sb], [0, sb, cb]], device=dev)
    return rx @ ry


def splat_views(pts_world, views, n_views, S, K, occ_grads, z_grads, rgb):
    """Rasterise + blend + backward for `views`; returns (rgba, grad wrt pts_world)."""
    from isopoints_b200 import splat
    dev = pts_world.device
    P = pts_world.shape[0]
    p = pts_world.detach().requires_grad_(True)
    nv = len(views)
    if nv == 0 or P == 0:
        return None, torch.zeros_like(pts_world)
    scr = []
    for v in views:
        q = p @ view_rotation(v, n_views, dev).T
        scr.append(torch.stack([q[:, 0] * 0.45, q[:, 1] * 0.45, q[:, 2] + 3.0], dim=1))
    scr = torch.cat(scr, 0)
    sig = 1.5 * 2.0 / S
    ell = torch.tensor([1 / sig ** 2, 0.0, 1 / sig ** 2], device=dev).expand(nv * P, 3).contiguous()
    radii = torch.full((nv * P, 2), sig, device=dev)
    cutoff = torch.ones(1, device=dev)
    first = torch.arange(nv, device=dev, dtype=torch.int64) * P
    num = torch.full((nv,), P, device=dev, dtype=torch.int64)
    # raster + RGBA blend in one pass over the pixels (the renderer's path, splat.SplatRender)
    idx, zbuf, qv, occ, rgba = splat.SplatRender.apply(scr, ell, cutoff.expand(nv * P), radii, first, num, 0.05, S, K,
                                                       64 if S > 512 else 32, 10.0, None, rgb.repeat(nv, 1),
                                                       splat.NORM_WEIGHT_EPS)
    og = torch.stack([occ_grads[v] for v in views])
    zg = torch.stack([z_grads[v] for v in views])
    torch.autograd.backward([occ, zbuf], [og, zg])   # the incoming image-space gradients, no synthetic loss kernels
    return rgba, p.grad


def run(rank, world, dev, points=2_000_000, views=16, size=1024, sdf="siren", steps=5, check=False):
    """One record (python dict; identical on every rank).  `world` > 1 needs an initialised process group."""
    import torch.distributed as dist
    from isopoints_b200.dist import (ShardedUniformProjection, all_gather_varlen, all_reduce_point_grads,
                                     shard_range, shard_views)
    from isopoints_b200.levelset_sampling import UniformProjection
    from tests.helpers import SphereSDF, pinned_siren
    S, K, V = size, 8, views
    g = torch.Generator().manual_seed(7)
    b, e = shard_range(points, rank, world)
    # every rank draws the same stream and keeps its slice: identical to a single-rank run
    x = ((torch.rand(points, 3, generator=g) - 0.5) * 2)[b:e].to(dev)[None]
    net = {"siren": lambda: pinned_siren(0), "opaque": lambda: pinned_siren(0).as_opaque(),
           "sphere": SphereSDF}[sdf]().to(dev)
    kw = dict(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    proj = ShardedUniformProjection(**kw) if world > 1 else UniformProjection(**kw)
    gg = torch.Generator().manual_seed(9)
    occ_grads = [(torch.randn(S, S, generator=gg) * (torch.rand(S, S, generator=gg) < 0.1)).to(dev) for _ in range(V)]
    z_grads = [torch.randn(S, S, K, generator=gg).to(dev) for _ in range(V)]
    my_views = shard_views(V, rank, world)

    def ev():
        t = torch.cuda.Event(enable_timing=True)
        t.record()
        return t

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    acc = None
    iso = grad = rgb = None
    for it in range(steps + 1):            # iteration 0 warms up (allocator, NCCL channels, clocks)
        sync()
        t0 = ev()
        out = proj.project_points(x, net, skip_upsampling=True)
        t1 = ev()
        loc = out["levelset_points"][0][out["mask"][0]]
        iso = all_gather_varlen(loc)[0] if world > 1 else loc
        t2 = ev()
        rgb = (iso * 0.5 + 0.5).clamp(0, 1).contiguous()
        rgba, grad = splat_views(iso, my_views, V, S, K, occ_grads, z_grads, rgb)
        t3 = ev()
        if world > 1:
            all_reduce_point_grads(grad)
        t4 = ev()
        torch.cuda.synchronize()
        cur = torch.tensor([t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3), t3.elapsed_time(t4),
                            t0.elapsed_time(t4)], dtype=torch.float64, device=dev)
        if it > 0:
            acc = cur if acc is None else acc + cur
    acc = acc / steps
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.MAX)
    pr, ag, sp, ar, tot = acc.tolist()
    ok = None
    if check:
        _, full = splat_views(iso, list(range(V)), V, S, K, occ_grads, z_grads, rgb)
        scale = full.abs().amax(0).clamp_min(1e-20)
        okt = torch.tensor([1 if torch.allclose(grad / scale, full / scale, rtol=1e-4, atol=2e-5) else 0], device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(int(okt.item()))
    n_iso = int(iso.shape[0])
    return {"config": "C5: %d points, %s SDF (SURVEY 8d pinned Siren), project+resample+splat %dx%d^2, %d GPU(s); "
                      "strong scaling of one fixed problem" % (points, sdf, V, S, world),
            "n_gpus": world, "steps": steps, "points": points, "iso_points": n_iso,
            "iso_points_per_s": points / (tot * 1e-3), "total_ms": tot, "project_resample_ms": pr,
            "allgather_points_ms": ag, "splat_fwd_blend_bwd_ms": sp, "allreduce_grads_ms": ar,
            "views_per_rank": len(my_views), "splat_ms_per_view": sp / max(len(shard_views(V, 0, world)), 1),
            "allgather_bytes": n_iso * 12, "allreduce_bytes": n_iso * 12,
            "grad_check_vs_all_views_on_one_rank": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=2_000_000)
    ap.add_argument("--views", type=int, default=16)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--sdf", default="siren", choices=["siren", "opaque", "sphere"])
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    torch.backends.cuda.matmul.allow_tf32 = False
    rec = run(rank, world, dev, args.points, args.views, args.size, args.sdf, args.steps, args.check)
    if rank == 0:
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rec["grad_check_vs_all_views_on_one_rank"] is False:
        sys.exit(1)


if __name__ == "__main__":
    main()
