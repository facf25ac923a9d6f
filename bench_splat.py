"""Second headline metric of BASELINE.json: pixel-splats / second, config C4
(8 views x 300 000 iso-point splats, 512^2, K = 8, sigma = 1.5 px, fwd and fwd+bwd).
Called by bench.py (rank 0, N = 1) and runnable alone:  python bench_splat.py"""
import json
import os
import sys
import time

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

V, PV, S, K = 8, 300_000, 512, 8


def _inputs(dev):
    from tests.helpers import make_splat_inputs
    inp = make_splat_inputs(V, PV, S, seed=0, sigma_px=1.5, aniso=False, behind_frac=0.0)
    return {k: torch.as_tensor(v) for k, v in inp.items()}


def run_ewa(dev, peaks, peak_src, steps, flush):
    """The step right before the splat: per-point EWA parameters for the C4 point set
    (V views x PV world-space points near a sphere, one camera per view) and the renderable mask.
    Algorithmic bytes per point: 12 (xyz) + 12 (normal) + 4 (h_k) in, 8 + 12 + 4 + 4 out = 56 B;
    the mask kernel: 24 B in, 1 B out."""
    from isopoints_b200 import _ext, ewa
    from tests.helpers import make_cameras, make_surface_points
    pts, nrm, first, num = make_surface_points([PV] * V, seed=0)
    w2v, proj, _ = make_cameras(V, seed=1)
    pts, nrm, first, w2v, proj = (x.to(dev) for x in (pts, nrm, first, w2v, proj))
    h = (torch.rand(V * PV, device=dev) * 1e-3 + 5e-5)
    out = {}
    for name, fn, nbytes in (
            ("ewa_point_params", lambda: ewa.get_per_point_info(pts, nrm, first, proj, h, S, 1.0, 1.0), 56),
            ("renderable_mask", lambda: ewa.renderable_mask(pts, nrm, first, w2v, 1.0, 100.0, True), 25)):
        for _ in range(3):
            fn()
        _ext.PROFILE = {}
        for k in range(steps):
            flush.fill_(k & 0xff)
            fn()
        torch.cuda.synchronize()
        prof, _ext.PROFILE = _ext.PROFILE, None
        ev = prof["isob200_" + name]
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        ach = nbytes * V * PV / (ms * 1e-3) / 1e9
        out[name] = {"points": V * PV, "avg_launch_ms": ms, "points_per_s": V * PV / (ms * 1e-3),
                     "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                  "frac": ach / peaks["hbm_gbs"], "peak_source": peak_src,
                                  "algorithmic_bytes_per_launch": nbytes * V * PV}}
    return out


def run(args, dev, peaks, peak_src, steps=None):
    from isopoints_b200 import _ext, splat
    lib = _ext.lib()
    steps = steps or max(3, args.steps)
    host = _inputs(dev)
    pin = {k: v.pin_memory() for k, v in host.items()}
    t = {k: v.to(dev) for k, v in host.items()}
    g = torch.Generator().manual_seed(0)
    occ_grad = (torch.randn(V, S, S, generator=g) * (torch.rand(V, S, S, generator=g) < 0.1)).to(dev)
    zbuf_grad = torch.randn(V, S, S, K, generator=g).to(dev)
    rgb = torch.rand(V * PV, 3, generator=g).to(dev)
    scaler = (torch.rand(V * PV, generator=g) + 0.5).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_pairs = splat.count_pixel_splats(t["points"], t["ellipse"], t["cutoff"], t["radii"], S)

    def fwd(tt, pts):
        return splat.EllipticalRasterizer.apply(pts, tt["ellipse"], tt["cutoff"], tt["radii"], tt["first_idx"],
                                                tt["num_points"], 0.05, S, K, 32 if S <= 512 else 64, 0, 10.0)

    def fwd_bwd(tt, with_loss=False):
        # the renderer's path (ewa.SurfaceSplattingRenderer): raster + RGBA blend in one pass (splat.SplatRender)
        pts = tt["points"].detach().requires_grad_(True)
        idx, zbuf, qv, occ, img = splat.SplatRender.apply(
            pts, tt["ellipse"], tt["cutoff"], tt["radii"], tt["first_idx"], tt["num_points"], 0.05, S, K,
            32 if S <= 512 else 64, 10.0, scaler, rgb, splat.NORM_WEIGHT_EPS)
        if with_loss:   # round-1 form: a synthetic loss whose own elementwise kernels (~0.4 GB of HBM traffic) are timed too
            ((occ * occ_grad).sum() + (zbuf * zbuf_grad).sum()).backward()
        else:           # the rasteriser's backward alone: the same two gradients handed to autograd directly
            torch.autograd.backward([occ, zbuf], [occ_grad, zbuf_grad])
        return img, pts.grad

    for _ in range(3):
        fwd_bwd(t)
    torch.cuda.synchronize()

    def timed(fn, n):
        ms = 0.0
        for k in range(n):
            flush.fill_(k & 0xff)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        return ms / n

    with torch.no_grad():
        ms_fwd = timed(lambda: fwd(t, t["points"]), steps)
    l0 = lib.isob200_launch_count()
    ms_fb = timed(lambda: fwd_bwd(t), steps)
    launches = (lib.isob200_launch_count() - l0) / steps
    for _ in range(3):
        fwd_bwd(t, with_loss=True)
    ms_fb_loss = timed(lambda: fwd_bwd(t, with_loss=True), steps)
    _ext.PROFILE = {}            # per-entry CUDA events: a separate pass, so that they do not sit in ms_fb
    timed(lambda: fwd_bwd(t), steps)
    prof, _ext.PROFILE = _ext.PROFILE, None
    kern = {n.replace("isob200_", ""): {"calls_per_step": len(p) / steps, "avg_ms": sum(a.elapsed_time(b) for a, b in p) / len(p)}
            for n, p in prof.items()}

    img_pin = torch.empty((V, S, S, 4), dtype=torch.float32).pin_memory()
    grad_pin = torch.empty((V * PV, 3), dtype=torch.float32).pin_memory()

    def e2e():
        tt = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
        img, grad = fwd_bwd(tt)
        img_pin.copy_(img, non_blocking=True)       # results land in pinned host buffers
        grad_pin.copy_(grad, non_blocking=True)
        torch.cuda.synchronize()
        return img_pin, grad_pin
    e2e()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        img, grad = e2e()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) / steps * 1e3
    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = img.numel() * 4 + grad.numel() * 4

    # dominant kernel: the raster kernel inside splat_forward; algorithmic bytes per launch
    # (SURVEY 8d): 36 B per point in + (12K + 4) B per pixel out
    alg = 36 * V * PV + (12 * K + 4) * V * S * S
    f = (kern.get("splat_forward_fused") or kern.get("splat_forward"))
    roof = None
    if f:
        ach = alg / (f["avg_ms"] * 1e-3) / 1e9
        roof = {"kernel": "splat_tile_fill + splat_raster_kernel<8>", "bound": "hbm", "achieved": ach,
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": _ncu_traffic("prof_splat_raster"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "avg_launch_ms": f["avg_ms"],
                "limiter": "instruction issue / shared-memory atomics, DRAM traffic = algorithmic bytes",
                # the raster kernel alone (the timed entry point also runs the tile fill): its ncu duration
                "issue": _issue_ceiling("prof_splat_raster", _ncu("prof_splat_raster").get("duration_us", 0) / 1e3),
                "ncu": _ncu("prof_splat_raster")}
    ewa_rec = run_ewa(dev, peaks, peak_src, steps, flush)
    return {"ewa_point_params": ewa_rec, "metric": "pixel-splats/sec", "unit": "pixel-splats/s",
            "config": {"workload": "C4: %d views x %d splats, %dx%d, K=%d, sigma=1.5px, occ_grad on 10%% of pixels, "
                                   "radii_backward_scaler=10" % (V, PV, S, S, K), "l2": "flushed between steps"},
            "pixel_splats_per_call": n_pairs,
            "value_fwd": n_pairs / (ms_fwd * 1e-3), "ms_fwd": ms_fwd,
            "value": n_pairs / (ms_fb * 1e-3), "ms_fwd_blend_bwd": ms_fb,
            "ms_fwd_blend_bwd_with_synthetic_loss": ms_fb_loss,
            "e2e": {"value": n_pairs / (ms_e2e * 1e-3), "unit": "pixel-splats/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches_per_step": launches, "roofline": roof, "kernels": kern}


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
