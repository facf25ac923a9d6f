"""GPU bring-up diagnostics for csrc/siren.cu: per-GEMM raw accumulators vs fp64, end-to-end sdf/grad
error vs fp64 autograd next to the fp32 autograd error, and a first timing."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from tests.helpers import Siren  # noqa: E402
from isopoints_b200 import siren  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def ref64(model, x):
    m = Siren(256, len(model.net) - 2, float(model.net[0].omega_0)).double()
    m.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    m = m.to(dev)
    xx = x.double().clone().requires_grad_(True)
    s = m(xx).sdf
    g, = torch.autograd.grad(s, xx, torch.ones_like(s))
    return s.detach().reshape(-1), g.detach(), m


def ref32(model, x):
    xx = x.clone().requires_grad_(True)
    s = model(xx).sdf
    g, = torch.autograd.grad(s, xx, torch.ones_like(s))
    return s.detach().reshape(-1), g.detach()


def layer_acts64(m64, x):
    hs = []
    h = x.double()
    for lyr in list(m64.net)[:-1]:
        h = torch.sin(lyr.omega_0 * lyr.linear(h))
        hs.append(h)
    return hs


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    model = Siren(256, L, 30.0, seed=0).to(dev)
    torch.manual_seed(0)
    x = ((torch.rand(1000, 3, device=dev) - 0.5) * 2).contiguous()
    s64, g64, m64 = ref64(model, x)
    s32, g32 = ref32(model, x)
    print("fp32 autograd vs fp64: sdf %.3e grad %.3e (|grad| max %.3e)" % (
        (s32.double() - s64).abs().max().item(), (g32.double() - g64).abs().max().item(), g64.abs().max().item()))
    # GEMM 0 raw accumulator
    spec = siren.match(model)
    assert spec is not None
    blob, scratch, LL = siren.packed(model, spec)
    hdr = blob[:1024].view(torch.float32)
    print("hdr wscale_inv", hdr[:LL].tolist(), "gl", hdr[64].item(), hdr[65].item(), "b_last", hdr[66].item(),
          hdr[67].item(), hdr[68].item())
    hs = layer_acts64(m64, x[:128])
    for g in [0, 1, L - 1]:
        if g < 0 or g >= L:
            continue
        out = siren.sdf_and_grad(model, x, dbg_gemm=g)
        torch.cuda.synchronize()
        acc = out[2].double()
        wsi = hdr[g].double().item()
        W = m64.net[g + 1].linear.weight
        exp = (hs[g] * 4096.0) @ (W / wsi).t()
        err = (acc - exp).abs().max().item()
        print("GEMM %d raw acc: max|acc| %.4e  max err %.4e  rel %.3e" % (g, exp.abs().max().item(), err,
                                                                           err / exp.abs().max().item()))
        if err / exp.abs().max().item() > 1e-4:
            bad = ((acc - exp).abs() > 1e-4 * exp.abs().max()).nonzero()
            print("  bad entries:", bad.shape[0], "first", bad[:8].tolist())
            print("  acc[0,:8]", acc[0, :8].tolist())
            print("  exp[0,:8]", exp[0, :8].tolist())
            print("  acc[:8,0]", acc[:8, 0].tolist())
            print("  exp[:8,0]", exp[:8, 0].tolist())
    for n in [1000, 1, 128, 129, 148 * 128 * 2 + 77]:
        xx = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
        s64, g64, _ = ref64(model, xx)
        s32, g32 = ref32(model, xx)
        sf, gf = siren.sdf_and_grad(model, xx)
        torch.cuda.synchronize()
        print("n=%d  fused vs fp64: sdf %.3e grad %.3e | fp32 autograd vs fp64: sdf %.3e grad %.3e" % (
            n, (sf.double() - s64).abs().max().item(), (gf.double() - g64).abs().max().item(),
            (s32.double() - s64).abs().max().item(), (g32.double() - g64).abs().max().item()))
    # timing
    n = 200000
    xx = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
    for _ in range(3):
        siren.sdf_and_grad(model, xx)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        siren.sdf_and_grad(model, xx)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    fl = n * 2 * 2 * L * 256 * 256
    print("fused: n=%d  %.3f ms  -> %.1f Mpt/s, %.1f TFLOP/s fp32-equivalent" % (n, ms, n / ms / 1e3, fl / ms / 1e9))
    for _ in range(2):
        ref32(model, xx)
    torch.cuda.synchronize()
    a.record()
    for _ in range(3):
        ref32(model, xx)
    b.record()
    torch.cuda.synchronize()
    print("torch autograd fp32: %.3f ms" % (a.elapsed_time(b) / 3))


if __name__ == "__main__":
    main()
