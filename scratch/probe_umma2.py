"""2-CTA tcgen05 probe: check the accumulator layout lane = row + 64*(col>=128), column = col % 128."""
import sys
import torch
sys.path.insert(0, ".")
from isopoints_b200 import _ext
lib = _ext.lib()
dev = torch.device("cuda")
for K in (16, 32, 64):
    torch.manual_seed(K)
    A = torch.randn(128, K, device=dev)
    B = torch.randn(256, K, device=dev)
    dump = torch.full((2, 128, 128), float("nan"), device=dev)
    _ext.check(lib.isob200_umma2_probe(_ext.ptr(A), _ext.ptr(B), K, _ext.ptr(dump), _ext.stream(dev)))
    torch.cuda.synchronize()
    ref = A.half().float() @ B.half().float().t()          # (128, 256)
    got = torch.empty_like(ref)
    for r in range(2):
        got[64 * r:64 * r + 64, :128] = dump[r, :64]
        got[64 * r:64 * r + 64, 128:] = dump[r, 64:]
    err = (got - ref).abs().max().item()
    print("K=%d  max err (assumed layout) %.3e   ref max %.2f" % (K, err, ref.abs().max().item()))
    if not err < 1e-2:
        # try to identify the layout: for a few dump entries find the matching ref entry
        for r in range(2):
            for lane in (0, 1, 16, 32, 63, 64, 65, 96, 127):
                for col in (0, 1, 64, 127):
                    v = dump[r, lane, col].item()
                    m = ((ref - v).abs() < 1e-3).nonzero()
                    print("  cta %d lane %3d col %3d = %9.4f  -> ref idx %s" % (r, lane, col, v, m[:3].tolist()))
