"""Point-to-ray search of the in-surface sampler (Model.sample_offsurface_using_isopoints,
DSS/models/combined_modeling.py:325-352) at a training-step shape: 2048 pixel rays against 20 000 visible
iso-points, both searches (front / back) of one view.  Runnable alone:  python bench_rays.py [--steps K]
Prints one JSON object: ray-point pairs / second of ``isob200_ray_nearest_point`` and of the reference's dense
PyTorch formulation (two (R,M) matrices + topk) on the same device."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R, M = 2048, 20_000


def _dense(cam_pos, ray0, pts):
    """combined_modeling.py:331-347 as written (PyTorch on the GPU)."""
    pC = pts - cam_pos.view(1, 3)
    ray_sq = (pC[None, :, :] * ray0[:, None, :]).sum(-1) ** 2
    dist_to_ray = (pC ** 2).sum(-1).unsqueeze(0) - ray_sq
    _, nn_idx = torch.topk(dist_to_ray, k=1, dim=1, largest=False)
    return torch.gather(ray_sq, 1, nn_idx).view(-1), nn_idx.view(-1)


def run(dev, steps=20):
    from isopoints_b200 import offsurface
    g = torch.Generator().manual_seed(0)
    d = torch.randn(M, 3, generator=g)
    front = (0.6 * d / d.norm(dim=-1, keepdim=True)).to(dev)
    back = (-front).contiguous()
    cam = torch.tensor([0.5, -1.0, 2.3], device=dev)
    tgt = ((torch.rand(R, 3, generator=g) - 0.5) * 0.8).to(dev)
    rays = torch.nn.functional.normalize(tgt - cam, dim=-1).contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"workload": "%d rays x %d points, 2 searches (front / back), L2 flushed between steps" % (R, M)}
    fns = {"ray_nearest_point_kernel": lambda: (offsurface.closest_point_to_rays(cam, rays, back),
                                                offsurface.closest_point_to_rays(cam, rays, front)),
           "dense_pytorch": lambda: (_dense(cam, rays, back), _dense(cam, rays, front))}
    res = {}
    for name, fn in fns.items():
        for _ in range(3):
            res[name] = fn()
        torch.cuda.synchronize()
        ms = 0.0
        for k in range(steps):
            flush.fill_(k & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res[name] = fn()
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        ms /= steps
        out[name] = {"ms_per_step": ms, "pairs_per_s": 2.0 * R * M / (ms * 1e-3),
                     "gflops": 2.0 * R * M * 17 / (ms * 1e-3) / 1e9}
    same = [(res["ray_nearest_point_kernel"][k][1] == res["dense_pytorch"][k][1]).float().mean().item() for k in (0, 1)]
    out["same_point_as_dense_frac"] = same
    out["speedup"] = out["dense_pytorch"]["ms_per_step"] / out["ray_nearest_point_kernel"]["ms_per_step"]
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(run(torch.device("cuda", 0), a.steps)))
