"""Bring-up of the CTA-pair SIREN kernel: compare with the single-CTA kernel / fp64, then time it."""
import sys
import torch
sys.path.insert(0, ".")
from tests.helpers import Siren
from isopoints_b200 import siren, _ext
lib = _ext.lib()
dev = "cuda"
L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
model = Siren(256, L, 30.0, seed=0).to(dev)


def ref64(x):
    m = Siren(256, L, 30.0).double()
    m.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    m = m.to(dev)
    xx = x.double().clone().requires_grad_(True)
    s = m(xx).sdf
    g, = torch.autograd.grad(s, xx, torch.ones_like(s))
    return s.detach().reshape(-1), g.detach()


sizes = [int(a) for a in sys.argv[2:]] or [100, 128, 129, 300, 5000, 40000]
for n in sizes:
    torch.manual_seed(n)
    x = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
    lib.isob200_siren_set_pair_mode(0)
    s1, g1 = siren.sdf_and_grad(model, x)
    torch.cuda.synchronize()
    lib.isob200_siren_set_pair_mode(1)
    s2, g2 = siren.sdf_and_grad(model, x)
    torch.cuda.synchronize()
    s64, g64 = ref64(x)
    print("n=%6d pair vs single: sdf %.3e grad %.3e | pair vs fp64: sdf %.3e grad %.3e | single vs fp64: %.3e %.3e" % (
        n, (s1 - s2).abs().max().item(), (g1 - g2).abs().max().item(),
        (s2.double() - s64).abs().max().item(), (g2.double() - g64).abs().max().item(),
        (s1.double() - s64).abs().max().item(), (g1.double() - g64).abs().max().item()), flush=True)
n = 200000
x = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
for mode in (0, 1, 0, 1):
    lib.isob200_siren_set_pair_mode(mode)
    for _ in range(3):
        siren.sdf_and_grad(model, x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        siren.sdf_and_grad(model, x)
    b.record()
    torch.cuda.synchronize()
    print("mode %d: n=%d %.3f ms" % (mode, n, a.elapsed_time(b) / 10), flush=True)
for nn in (100, 2000, 19000):
    nd = torch.tensor([nn], dtype=torch.int32, device=dev)
    for mode in (0, 1):
        lib.isob200_siren_set_pair_mode(mode)
        for _ in range(3):
            siren.sdf_and_grad(model, x, n_dev=nd)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            siren.sdf_and_grad(model, x, n_dev=nd)
        b.record()
        torch.cuda.synchronize()
        print("mode %d: live rows %d  %.1f us" % (mode, nn, a.elapsed_time(b) * 100), flush=True)
lib.isob200_siren_set_pair_mode(0)
import ctypes
import numpy as np
lib.isob200_siren_set_pair_mode(1)
siren.sdf_and_grad(model, x)
torch.cuda.synchronize()
buf = np.zeros(512, dtype=np.int64)
_ext._RAW.isob200_siren_pair_stamps(buf.ctypes.data_as(ctypes.c_void_p), 512)
t = buf.reshape(64, 8)
base = t[0, 0]
print(" gi s  mma_first mma_done  wait_a wait_w wait_peer  span | epi sees acc_full")
for gi in range(min(4 * L + 4, 64)):
    r = t[gi]
    print("%3d %d %9d %9d %7d %6d %6d %7d | %9d" % (gi, r[5], r[0] - base, r[1] - base, r[2], r[3], r[4], r[1] - r[0], r[6] - base))
lib.isob200_siren_set_pair_mode(0)
