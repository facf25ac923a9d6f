"""BASELINE.json configs[2] ("C3"): 500 000 points, FRNN radius = 0.05, K = 16 grid query, reported under
"frnn" in bench.py's JSON line (N = 1) and runnable alone:  python bench_frnn.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def _ncu(capture):
    """Record of the committed `ncu --set full` capture `capture` (profiles/ncu_traffic.json), or {}."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[capture]
    except Exception:
        return {}


def _ncu_traffic(capture):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from that capture, or None."""
    return _ncu(capture).get("dram_bytes_per_launch")


def _issue_ceiling(capture, launch_ms, sm_mhz=1965.0):
    """For a kernel that ncu shows to be instruction-issue bound, the ceiling is the issue rate: 148 SMs x 4
    schedulers x 1 warp instruction per cycle.  warp instructions per launch come from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json), the launch time is measured live."""
    n = _ncu(capture).get("warp_inst")
    if not n or not launch_ms:
        return None
    peak = 148 * 4 * sm_mhz * 1e6 / 1e9
    ach = n / (launch_ms * 1e-3) / 1e9
    return {"warp_inst_per_launch": n, "achieved_ginst_s": ach, "peak_ginst_s": peak, "frac": ach / peak,
            "peak_is": "148 SMs x 4 schedulers x %.0f MHz (max SM clock)" % sm_mhz}
P, K, R = 500_000, 16, 0.05


def run(args, dev, peaks, peak_src, steps=None):
    from isopoints_b200 import _ext, frnn
    steps = steps or max(5, args.steps)
    g = torch.Generator().manual_seed(0)
    host = torch.rand(1, P, 3, generator=g)
    pin = host.pin_memory()
    p = host.to(dev)
    lens = torch.tensor([P], device=dev)
    r = torch.tensor([R], device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def timed(n):
        out = []
        for k in range(n):
            flush.fill_(k & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = frnn.frnn_grid_points(p, p, lens, lens, K=K, r=r)
            b.record()
            torch.cuda.synchronize()
            out.append(a.elapsed_time(b))
            keep.append(res)          # the previous call's outputs stay alive while the next ones are allocated,
            del keep[:-1]             # as in a caller that rebinds `dists, idxs = ...` (two sets of blocks)
        return out

    keep = []
    timed(5)                          # warm-up with the timed loop's allocation pattern (a first cudaMalloc costs ms)
    each = timed(steps)
    ms = sum(each) / steps
    _ext.PROFILE = {}                 # per-entry CUDA events: a separate pass, so that they do not sit in `ms`
    timed(2)
    _ext.PROFILE = {}
    timed(steps)
    prof, _ext.PROFILE = _ext.PROFILE, None
    d, i = keep[-1][0], keep[-1][1]
    kern = {n.replace("isob200_", ""): sum(x.elapsed_time(y) for x, y in v) / len(v) for n, v in prof.items()}
    import time
    dh = torch.empty((1, P, K), dtype=torch.float32).pin_memory()
    ih = torch.empty((1, P, K), dtype=torch.int64).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        x = pin.to(dev, non_blocking=True)
        d, i, _, _ = frnn.frnn_grid_points(x, x, lens, lens, K=K, r=r)
        dh.copy_(d, non_blocking=True)              # results land in pinned host buffers
        ih.copy_(i, non_blocking=True)
        torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / steps * 1e3
    alg = P * (16 + 12 * K)                      # 16 + 12K bytes per query
    q = kern.get("frnn_find_nbrs")
    roof = None
    if q:
        ach = alg / (q * 1e-3) / 1e9
        roof = {"kernel": "frnn_query_collect_kernel<3,int64>", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": _ncu_traffic("prof_frnn_query"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "avg_launch_ms": q,
                "limiter": "instruction issue / L1 latency at 16 warps per SM, not HBM (candidates are served by L1/L2: "
                           "~200 candidate tests + a K-pass selection per query)",
                "issue": _issue_ceiling("prof_frnn_query", q), "ncu": _ncu("prof_frnn_query")}
    return {"metric": "FRNN queries/sec", "unit": "queries/s",
            "config": {"workload": "C3: %d uniform points in the unit box, self query, K=%d, r=%g, radius_cell_ratio=2"
                                   % (P, K, R), "l2": "flushed between steps"},
            "value": P / (ms * 1e-3), "ms_per_step": ms, "avg_neighbours_found": float((i >= 0).float().sum(-1).mean()),
            "e2e": {"value": P / (e2e_ms * 1e-3), "unit": "queries/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": dh.numel() * 4 + ih.numel() * 8},
            "roofline": roof, "kernels_avg_ms": kern, "ms_each_step": each}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))); src = "measured"
    except Exception:
        peaks, src = {"hbm_gbs": 6650.0}, "fallback"
    print(json.dumps(run(a, torch.device("cuda", 0), peaks, src)))
