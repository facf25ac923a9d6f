"""The reference itself on the GPU: its own Python (staged under baseline/_ref by oracle/stage_ref.py, or
/root/reference) on its own CUDA extensions recompiled unmodified for sm_100a (oracle/_ref) -- the kernels
SURVEY 8d names as "the ones to beat", timed on the BASELINE configs on the same box as bench.py's own arm.

TEST / BENCH INFRASTRUCTURE ONLY (bench.py runs this module in a SUBPROCESS, outside every timed region, and
copies its JSON into the `ref_cuda` key; nothing under isopoints_b200/ imports it).

    python -m oracle.ref_gpu [--only c2,c3,c4] [--reps 3]

  c2  reference UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
      .project_points(x, Siren, skip_upsampling=True) (DSS/models/levelset_sampling.py:353-439) with the
      reference's own Siren class (DSS/models/common.py:90-165, torch.manual_seed(0)) through autograd, fp32 with
      TF32 off, and the reference's frnn (external/FRNN/frnn/frnn.py:19-162 on frnn._C + prefix_sum);
  c3  reference frnn.frnn_grid_points(p, p, K=16, r=0.05) at 500 000 points (box and sphere);
  c4  reference rasterize_elliptical_points forward and EllipticalRasterizer.backward
      (DSS/core/rasterizer.py:678-973 on DSS._C + frnn._C + prefix_sum), 8 views x 300 000 splats at 512^2, K = 8.
pytorch3d is absent: its packed/padded helpers are the pure-torch stand-ins of oracle/ref_python.py (a per-view
slice loop, 8 views), everything else is the reference's code.
"""
import argparse
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _time(fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def c2(ref, reps, n_points=200_000):
    import importlib
    common = importlib.import_module("DSS.models.common")
    torch.manual_seed(0)
    net = common.Siren(dim=3, c_dim=0, hidden_size=256, n_layers=7, first_omega_0=30, hidden_omega_0=30,
                       outermost_linear=True).cuda()
    g = torch.Generator().manual_seed(1000)
    x = ((torch.rand(1, 200_000, 3, generator=g) - 0.5) * 2)[:, :n_points].contiguous().cuda()
    proj = ref.levelset_sampling.UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    out = {}

    def run():
        out["r"] = proj.project_points(x.clone(), net, skip_upsampling=True)
    ms = _time(run, reps)
    r = out["r"]
    return {"c2_ms": ms, "c2_points_per_s": n_points / (ms * 1e-3), "c2_points": n_points,
            "c2_rows_after_filter": int(r["mask"].shape[1]), "c2_valid_after_resample": int(r["mask"].sum()),
            "c2_what": "reference UniformProjection.project_points + reference Siren (autograd, fp32, TF32 off) + "
                       "reference frnn CUDA, on this GPU"}


def c3(frnn, reps):
    out = {}
    g = torch.Generator().manual_seed(0)
    for name, p in (("box", torch.rand(1, 500_000, 3, generator=g)),
                    ("sphere", torch.nn.functional.normalize(torch.randn(1, 500_000, 3, generator=g), dim=-1))):
        p = p.cuda()
        ms = _time(lambda: frnn.frnn_grid_points(p, p, None, None, K=16, r=0.05, return_nn=False), reps)
        out["c3_%s_ms" % name] = ms
    out["c3_ms"] = out["c3_box_ms"]
    return out


def c4(rast, reps):
    from tests.helpers import make_splat_inputs
    V, PV, S, K = 8, 300_000, 512, 8
    inp = make_splat_inputs(V, PV, S, seed=0, sigma_px=1.5, aniso=False, behind_frac=0.0)
    t = {k: torch.as_tensor(v).cuda() for k, v in inp.items()}
    gg = torch.Generator().manual_seed(0)
    occ_grad = (torch.randn(V, S, S, generator=gg) * (torch.rand(V, S, S, generator=gg) < 0.1)).cuda()
    zbuf_grad = torch.randn(V, S, S, K, generator=gg).cuda()

    def screen(pts):
        return types.SimpleNamespace(points_packed=lambda: pts, cloud_to_packed_first_idx=lambda: t["first_idx"],
                                     num_points_per_cloud=lambda: t["num_points"])

    def fwd():
        return rast.rasterize_elliptical_points(screen(t["points"]), t["ellipse"], t["cutoff"], t["radii"],
                                                depth_merging_threshold=0.05, image_size=S, points_per_pixel=K,
                                                bin_size=32, radii_backward_scaler=10.0)

    def fwd_bwd():
        pts = t["points"].detach().requires_grad_(True)
        o = rast.rasterize_elliptical_points(screen(pts), t["ellipse"], t["cutoff"], t["radii"],
                                             depth_merging_threshold=0.05, image_size=S, points_per_pixel=K,
                                             bin_size=32, radii_backward_scaler=10.0)
        ((o[3] * occ_grad).sum() + (o[1] * zbuf_grad).sum()).backward()

    f = _time(fwd, reps)
    fb = _time(fwd_bwd, reps)
    return {"c4_fwd_ms": f, "c4_fwd_bwd_ms": fb, "c4_bwd_ms": fb - f}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="c2,c3,c4")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--c2-points", type=int, default=200_000)
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from oracle import ref_python
    out = {"natives": "oracle/_ref (reference CUDA sources compiled unmodified for sm_100a)",
           "python": ref_python.REF}
    try:
        mods = ref_python.use_natives("reference")
        ref = ref_python.load()
    except Exception as e:
        print(json.dumps({"unavailable": "%s: %s" % (type(e).__name__, e)}))
        return
    for name in args.only.split(","):
        try:
            if name == "c2":
                out.update(c2(ref, args.reps, args.c2_points))
            elif name == "c3":
                out.update(c3(mods["frnn"], args.reps))
            elif name == "c4":
                out.update(c4(ref_python.load_rasterizer(), args.reps))
        except Exception as e:      # one failing config must not hide the others
            out[name + "_error"] = "%s: %s" % (type(e).__name__, str(e)[:300])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
