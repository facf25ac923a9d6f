"""Ray marching onto the level set (SphereTracing.project_points) on the C2 network:
200 000 rays, random-init SIREN 8x256 (last bias shifted so the zero set fills the unit sphere), 10 marching
iterations.  Runnable alone:  python bench_trace.py [--steps K]
Prints one JSON object: rays / second with the fused forward-only tcgen05 kernel and with the same weights
behind an opaque nn.Module (autograd value + gradient per iteration, as the reference does)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS, ITERS = 200_000, 10


def _model(opaque):
    """The SURVEY-pinned SIREN (tests/helpers.pinned_siren) with its head bias shifted so that the zero set
    crosses the unit ball (a marching target); fused structure or the same weights behind an opaque module."""
    from tests.helpers import pinned_siren
    net = pinned_siren(0)
    with torch.no_grad():
        x = (torch.rand(4000, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 2
        net.net[-1].bias -= net(x).sdf.mean()
    return net.as_opaque() if opaque else net


def run(dev, steps=5):
    from isopoints_b200 import _ext, siren
    from isopoints_b200.levelset_sampling import SphereTracing
    from tests.helpers import make_rays
    torch.backends.cuda.matmul.allow_tf32 = False
    ray0, dirs = make_rays(RAYS, seed=0, target_radius=0.7)
    ray0, dirs = ray0.to(dev), dirs.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tracer = SphereTracing(proj_max_iters=ITERS, proj_tolerance=5e-5)
    out = {"workload": "%d rays, SIREN 8x256, %d marching iterations, L2 flushed between steps" % (RAYS, ITERS)}
    lib = _ext.lib()
    for name, opaque in (("fused_forward_only", False), ("opaque_module_autograd", True)):
        net = _model(opaque).to(dev)
        for _ in range(3):
            res = tracer.project_points(ray0, dirs, net)
        torch.cuda.synchronize()
        ms = 0.0
        l0 = lib.isob200_launch_count()
        for k in range(steps):
            flush.fill_(k & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = tracer.project_points(ray0, dirs, net)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        ms /= steps
        out[name] = {"ms_per_step": ms, "rays_per_s": RAYS / (ms * 1e-3), "hit_frac": float(res["mask"].float().mean()),
                     "gpu_launches_per_step": (lib.isob200_launch_count() - l0) / steps}
    out["speedup"] = out["opaque_module_autograd"]["ms_per_step"] / out["fused_forward_only"]["ms_per_step"]
    out["ray_tracing"] = run_ray_tracing(dev, steps, flush)
    return out


def run_ray_tracing(dev, steps, flush, cams=4, pixels=50_000):
    """IDR RayTracing.forward (levelset_sampling.py:810-1167; isopoints_b200/ray_tracing.py), eval mode, on the
    same network: 4 cameras x 50 000 pixel rays, 10 two-sided marching iterations, 100-sample search + 8 secant
    steps on the rays left over.  SDF through the forward-only fused kernel (siren.sdf_fn) vs the same weights
    evaluated by PyTorch under no_grad, as the reference's call site does (implicit_modeling.py:431)."""
    from isopoints_b200 import _ext, siren
    from isopoints_b200.ray_tracing import RayTracing
    from tests.helpers import make_camera_rays
    cam, dirs = make_camera_rays(cams, pixels, seed=1)
    cam, dirs = cam.to(dev), dirs.to(dev)
    object_mask = torch.ones(cams * pixels, dtype=torch.bool, device=dev)
    tracer = RayTracing().eval()
    lib = _ext.lib()
    rec = {"workload": "%d x %d rays, eval mode, reference defaults (10 marching iterations, n_steps 100, 8 secant "
                       "steps)" % (cams, pixels)}
    for name, opaque in (("fused_forward_only", False), ("torch_no_grad", True)):
        net = _model(opaque).to(dev).eval()
        if opaque:
            def sdf(x, net=net):
                with torch.no_grad():
                    return net(x).sdf.squeeze(-1)
        else:
            sdf = siren.sdf_fn(net)
        rows0 = siren.STATS["rows"]
        for _ in range(2):
            pts, mask, dist = tracer(sdf, cam, object_mask, dirs)
        torch.cuda.synchronize()
        ms = 0.0
        l0 = lib.isob200_launch_count()
        rows0 = siren.STATS["rows"]
        siren.MASKED_ROWS = []          # device counters of the march's masked evaluations (no host read-back there)
        for k in range(steps):
            flush.fill_(k & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            pts, mask, dist = tracer(sdf, cam, object_mask, dirs)
            b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        ms /= steps
        masked_rows = int(torch.stack(siren.MASKED_ROWS).sum().item()) if siren.MASKED_ROWS else 0
        siren.MASKED_ROWS = None
        rec[name] = {"ms_per_step": ms, "rays_per_s": cams * pixels / (ms * 1e-3), "hit_frac": float(mask.float().mean()),
                     "gpu_launches_per_step": (lib.isob200_launch_count() - l0) / steps,
                     "sdf_evaluations_per_step": (siren.STATS["rows"] - rows0 + masked_rows) / steps}
    rec["speedup"] = rec["torch_no_grad"]["ms_per_step"] / rec["fused_forward_only"]["ms_per_step"]
    return rec


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(run(torch.device("cuda", 0), a.steps)))
