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
    from tests.helpers import Siren, SirenSDF
    net = (SirenSDF if opaque else Siren)(256, 7, 30.0, seed=0)
    ref = Siren(256, 7, 30.0, seed=0)
    with torch.no_grad():
        x = (torch.rand(4000, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 2
        shift = ref(x).sdf.mean()
        (net.lin[-1] if opaque else net.net[-1]).bias -= shift
    return net


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
    return out


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    print(json.dumps(run(torch.device("cuda", 0), a.steps)))
