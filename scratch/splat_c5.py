"""Kernel breakdown of the splat stage at C5 shape (16 views x 2M points, 1024^2) and C4 for comparison."""
import argparse, json, sys
import torch
sys.path.insert(0, ".")
import bench_splat
from bench import _peaks
ap = argparse.Namespace(steps=3)
dev = torch.device("cuda", 0)
peaks, src = _peaks()
for V, PV, S in ((8, 300_000, 512), (16, 2_000_000, 1024), (16, 500_000, 1024)):
    bench_splat.V, bench_splat.PV, bench_splat.S = V, PV, S
    r = bench_splat.run(ap, dev, peaks, src, steps=3)
    print("V=%d PV=%d S=%d: pairs %d fwd %.3f ms, fwd+blend+bwd %.3f ms" % (V, PV, S, r["pixel_splats_per_call"], r["ms_fwd"], r["ms_fwd_blend_bwd"]))
    print("   ", {k: round(v["avg_ms"], 3) for k, v in r["kernels"].items()})
