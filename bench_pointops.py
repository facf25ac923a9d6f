"""Side record: farthest point sampling and sample_uniform_iso_points at the sizes the reference's
`sample_uniform_iso_points` reaches for 200 000 iso-points (levelset_sampling.py:1405-1445: wlop on ~0.6 * 4n
projected points, whose first step is farthest_sampling to <= n).  Runnable alone:  python bench_pointops.py"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def run(dev, n_points=200_000):
    from isopoints_b200 import _ext
    from isopoints_b200.levelset_sampling import sample_uniform_iso_points
    from tests.helpers import pinned_siren
    lib = _ext.lib()
    g = torch.Generator().manual_seed(0)
    P, M = 480_000, 200_000
    pts = torch.nn.functional.normalize(torch.randn(1, P, 3, generator=g), dim=-1).to(dev).contiguous()
    lens = torch.tensor([P], device=dev)
    out = {}

    def fps(m, coop):
        mm = torch.tensor([m], device=dev)
        idx = torch.empty((1, m), dtype=torch.int64, device=dev)
        nws = lib.isob200_fps_ws_floats(1, P) if coop else P
        ws = torch.empty((nws,), dtype=torch.float32, device=dev)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if coop:
            _ext.check(lib.isob200_fps_ws(_ext.ptr(pts), _ext.ptr(lens), _ext.ptr(mm), None, 1, P, m, _ext.ptr(ws),
                                          ws.numel(), _ext.ptr(idx), _ext.stream(dev)))
        else:
            _ext.check(lib.isob200_fps(_ext.ptr(pts), _ext.ptr(lens), _ext.ptr(mm), None, 1, P, m, _ext.ptr(ws),
                                       _ext.ptr(idx), _ext.stream(dev)))
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b), idx

    fps(1000, True)
    ms_c, idx_c = fps(M, True)
    ms_s, idx_s = fps(4000, False)
    out["fps"] = {"points": P, "samples": M, "all_sm_ms": ms_c, "all_sm_us_per_sample": ms_c * 1e3 / M,
                  "single_cta_us_per_sample": ms_s * 1e3 / 4000,
                  "single_cta_ms_extrapolated": ms_s / 4000 * M,
                  "same_first_4000": bool(torch.equal(idx_c[:, :4000], idx_s))}
    net = pinned_siren(0).to(dev)
    torch.manual_seed(0)
    sample_uniform_iso_points(net, 20_000)         # warm-up (packs the network, allocator)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pcl = sample_uniform_iso_points(net, n_points)
    torch.cuda.synchronize()
    out["sample_uniform_iso_points"] = {"n_points": n_points, "seconds": time.perf_counter() - t0,
                                        "iso_points_out": int(pcl.num_points_per_cloud().sum())}
    return out


if __name__ == "__main__":
    print(json.dumps(run(torch.device("cuda", 0))))
