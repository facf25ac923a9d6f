#!/usr/bin/env python
"""Developer tool: does the reverse-mode tape fit in L2?  Runs the fused SIREN kernel with fewer persistent CTAs
(isob200_siren_set_max_ctas) on 8 full waves of tiles each and prints the time per wave; under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` the same launches show the tape
traffic that left L2.  Not part of the product or the tests.

    python scripts/siren_cta_sweep.py [--stamps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isopoints_b200 import _ext, siren  # noqa: E402
from tests.helpers import pinned_siren  # noqa: E402


def main():
    dev = "cuda"
    lib = _ext.lib()
    model = pinned_siren(0).to(dev)
    reps = 1 if "--once" in sys.argv else 5
    for ctas in (148, 128, 112, 96, 80, 74, 64):
        lib.isob200_siren_set_max_ctas(ctas)
        rows = ctas * 128 * 8
        x = ((torch.rand(rows, 3, device=dev) - 0.5) * 2).contiguous()
        if reps > 1:
            for _ in range(2):
                siren.sdf_and_grad(model, x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            siren.sdf_and_grad(model, x)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print("ctas %4d rows %7d  %.4f ms  per wave %.2f us  %.1f M evals/s  tape live %.1f MB" % (
            ctas, rows, ms, ms * 1e3 / 8, rows / ms / 1e3, ctas * 6 * 128 * 1024 / 1e6))
        if "--stamps" in sys.argv:
            out = siren.sdf_and_grad(model, x, dbg_gemm=-2)
            torch.cuda.synchronize()
            t = out[2].view(torch.int64).reshape(-1, 8).cpu()[:14]
            heads = [int(t[g + 1, 3]) - int(t[g, 0]) for g in range(13)]
            print("   acc_full(G) -> first MMA of G+1, G=0..12:", heads)
    lib.isob200_siren_set_max_ctas(148)


if __name__ == "__main__":
    main()
