"""Per-tile time of the fused SIREN kernel against the number of persistent CTAs: does the tape (917 KB live
per CTA at 7 hidden layers, 136 MB at 148 CTAs against a 126 MB L2) slow full launches down?"""
import sys
import torch
sys.path.insert(0, ".")
from tests.helpers import Siren
from isopoints_b200 import _ext, siren
dev = "cuda"
lib = _ext.lib()
for L in (7, 3):
    model = Siren(256, L, 30.0, seed=0).to(dev)
    for ctas in (148, 132, 111, 74, 37):
        lib.isob200_siren_set_max_ctas(ctas)
        n = ctas * 128 * 10            # exactly 10 tiles per CTA
        x = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
        for _ in range(3):
            siren.sdf_and_grad(model, x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            siren.sdf_and_grad(model, x)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print("L=%d  %3d CTAs x 10 tiles: %.3f ms -> %.1f us per tile, tape live %.0f MB" % (
            L, ctas, ms, ms * 100, ctas * (L - 1 if L > 1 else 1) * 256 * 128 * 4 / 1e6))
lib.isob200_siren_set_max_ctas(148)

print("tape layers 1..n stored evict-first (L = 7, 148 CTAs, 200 000 rows):")
model = Siren(256, 7, 30.0, seed=0).to(dev)
x = ((torch.rand(200000, 3, device=dev) - 0.5) * 2).contiguous()
ref = None
for spill in (0, 1, 2, 3, 4, 6, 0):
    lib.isob200_siren_set_spill_layers(spill)
    for _ in range(3):
        out = siren.sdf_and_grad(model, x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        out = siren.sdf_and_grad(model, x)
    b.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = [t.clone() for t in out]
    same = all(torch.equal(r, t) for r, t in zip(ref, out))
    print("  spill %d: %.4f ms  (bit-identical: %s)" % (spill, a.elapsed_time(b) / 10, same))
lib.isob200_siren_set_spill_layers(0)
