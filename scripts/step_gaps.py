#!/usr/bin/env python
"""Developer tool: GPU idle gaps inside one C2 step (torch profiler / kineto): where the launch-bound glue sits.
Not part of the product or the tests."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isopoints_b200.levelset_sampling import UniformProjection  # noqa: E402
from tests.helpers import pinned_siren  # noqa: E402


def main():
    dev = "cuda"
    net = pinned_siren(0).to(dev)
    g = torch.Generator().manual_seed(1000)
    x = ((torch.rand(1, 200_000, 3, generator=g) - 0.5) * 2).to(dev)
    proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    for _ in range(5):
        proj.project_points(x, net, skip_upsampling=True)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        proj.project_points(x, net, skip_upsampling=True)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ev.sort(key=lambda e: e.time_range.start)
    t0 = ev[0].time_range.start
    prev_end = t0
    total_gap = 0.0
    print("%9s %8s %8s  kernel" % ("start_us", "dur_us", "gap_us"))
    for e in ev:
        gap = e.time_range.start - prev_end
        if gap > 0:
            total_gap += gap
        if gap > 4 or e.time_range.elapsed_us() > 50:
            print("%9.1f %8.1f %8.1f  %s" % (e.time_range.start - t0, e.time_range.elapsed_us(), gap, e.name[:90]))
        prev_end = max(prev_end, e.time_range.end)
    print("step span %.1f us, idle %.1f us over %d GPU activities" % (prev_end - t0, total_gap, len(ev)))


if __name__ == "__main__":
    main()
