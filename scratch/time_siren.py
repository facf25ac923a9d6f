"""Per-stage cycle stamps of CTA 0's first two tiles (dbg_gemm = -2)."""
import sys
import torch
sys.path.insert(0, ".")
from tests.helpers import Siren
from isopoints_b200 import siren
dev = "cuda"
L = 7
model = Siren(256, L, 30.0, seed=0).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
x = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
for _ in range(2):
    siren.sdf_and_grad(model, x)
import time
for code in (None,):
    for _ in range(2):
        siren.sdf_and_grad(model, x, dbg_gemm=code)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        siren.sdf_and_grad(model, x, dbg_gemm=code)
    b.record()
    torch.cuda.synchronize()
    print("variant", code, "ms", a.elapsed_time(b) / 5)
out = siren.sdf_and_grad(model, x, dbg_gemm=-2)
torch.cuda.synchronize()
t = out[2].view(torch.int64).reshape(-1, 8)[:4 * L].cpu()
base = int(t[0, 7])
print("  G  epi_wait_begin  acc_full   stage_end | mma_first  mma_issued  wait_w  wait_a | epi_dur  mma_span")
for g in range(4 * L):
    r = [int(v) for v in t[g]]
    print("%3d %10d %10d %10d | %10d %10d %7d %7d | %7d %7d" % (
        g, r[7] - base, r[0] - base, r[2] - base, r[3] - base, r[4] - base, r[5], r[6], r[2] - r[0], r[4] - r[3]))
