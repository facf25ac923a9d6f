#!/usr/bin/env python
"""Side-by-side device timing of the reference's OWN CUDA kernels (oracle/_ref, recompiled unmodified for
sm_100a) and this library on the BASELINE configs C3 (FRNN 500 k, K=16, r=0.05) and C4 (splat 8x300 k, 512^2,
K=8).  Evidence only (written to gpurun_out/ref_cuda_timing.json); not part of bench.py."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from isopoints_b200 import frnn, splat  # noqa: E402
from oracle import ref_native  # noqa: E402
from tests.helpers import make_splat_inputs  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = "cuda"
    out = {}
    g = torch.Generator().manual_seed(0)
    for name, p in (("box", torch.rand(1, 500_000, 3, generator=g)),
                    ("sphere", torch.nn.functional.normalize(torch.randn(1, 500_000, 3, generator=g), dim=-1))):
        p = p.to(dev)
        lens = torch.tensor([500_000], device=dev)
        r = torch.tensor([0.05], device=dev)
        ours = timeit(lambda: frnn.frnn_grid_points(p, p, lens, lens, K=16, r=r))
        ref = timeit(lambda: ref_native.frnn_grid_points_cuda(p, p, lens, lens, 16, r), n=3, warm=1)
        out["frnn_c3_" + name] = {"ours_ms": ours, "reference_cuda_ms": ref, "speedup": ref / ours,
                                  "queries_per_s_ours": 5e5 / (ours * 1e-3)}
    V, PV, S, K = 8, 300_000, 512, 8
    inp = make_splat_inputs(V, PV, S, seed=0, sigma_px=1.5, aniso=False, behind_frac=0.0)
    t = {k: torch.as_tensor(v, device=dev) for k, v in inp.items()}
    C = ref_native.dss_C()
    args = (t["points"], t["ellipse"], t["cutoff"], t["radii"], t["first_idx"], t["num_points"], 0.05, S, K, 32)
    ours = timeit(lambda: splat._C.splat_points(*args, 0))
    ref = timeit(lambda: C.splat_points(*args, PV), n=3, warm=1)
    out["splat_fwd_c4"] = {"ours_ms": ours, "reference_cuda_ms": ref, "speedup": ref / ours}
    idx = splat._C.splat_points(*args, 0)[0]
    gg = torch.Generator().manual_seed(0)
    occ_grad = (torch.randn(V, S, S, generator=gg) * (torch.rand(V, S, S, generator=gg) < 0.1)).to(dev)
    zbuf_grad = torch.randn(V, S, S, K, generator=gg).to(dev)

    def ours_bwd():
        pts = t["points"].detach().requires_grad_(True)
        o = splat.EllipticalRasterizer.apply(pts, t["ellipse"], t["cutoff"], t["radii"], t["first_idx"],
                                             t["num_points"], 0.05, S, K, 32, 0, 10.0)
        ((o[3] * occ_grad).sum() + (o[1] * zbuf_grad).sum()).backward()

    ours_fb = timeit(ours_bwd)
    ref_b = timeit(lambda: ref_native.splat_backward_fast_cuda(t["points"], t["radii"], idx, t["first_idx"],
                                                               t["num_points"], occ_grad, zbuf_grad, 10.0), n=3, warm=1)
    out["splat_c4_backward"] = {"ours_fwd_plus_bwd_ms": ours_fb, "ours_bwd_only_ms_approx": ours_fb - ours,
                                "reference_cuda_bwd_ms": ref_b, "speedup_bwd": ref_b / max(ours_fb - ours, 1e-6)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_timing.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
