#!/usr/bin/env python
"""Developer tool: per-stage cycle stamps of CTA 0's first two tiles of the fused SIREN kernel (dbg_gemm = -2),
plus the launch time at a few row counts.  Not part of the product or the tests.

    python scripts/siren_timeline.py [rows]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isopoints_b200 import siren  # noqa: E402
from tests.helpers import pinned_siren  # noqa: E402


def main():
    dev = "cuda"
    L = 7
    model = pinned_siren(0).to(dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    for rows in (n, 148 * 128, 148 * 128 * 4, 500, 128 * 148 * 10 + 7):
        x = ((torch.rand(rows, 3, device=dev) - 0.5) * 2).contiguous()
        for _ in range(3):
            siren.sdf_and_grad(model, x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            siren.sdf_and_grad(model, x)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print("rows %8d  value+grad %.4f ms  %.1f M evals/s" % (rows, ms, rows / ms / 1e3))
        a.record()
        for _ in range(10):
            siren.sdf(model, x)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print("rows %8d  value only %.4f ms  %.1f M evals/s" % (rows, ms, rows / ms / 1e3))
    from isopoints_b200 import _ext
    lib = _ext.lib()
    x = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
    for cyc in (0, 83500, 0, 83500):     # phase stagger of the odd-SM CTAs (0 = off, the default)
        lib.isob200_siren_set_stagger(cyc, 0)
        for _ in range(3):
            siren.sdf_and_grad(model, x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            siren.sdf_and_grad(model, x)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print("stagger %6d cycles: rows %8d  value+grad %.4f ms  %.1f M evals/s" % (cyc, n, ms, n / ms / 1e3))
        a.record()
        for _ in range(10):
            siren.sdf(model, x)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        print("                                         rows %8d  value only %.4f ms  %.1f M evals/s" % (n, ms, n / ms / 1e3))
    lib.isob200_siren_set_stagger(0, 0)
    out = siren.sdf_and_grad(model, x, dbg_gemm=-2)
    torch.cuda.synchronize()
    raw = out[2].view(torch.int64).reshape(-1, 8).cpu()
    t = raw[:4 * L]
    base = int(t[0, 7])
    print("  G  epi_wait_begin  acc_full   stage_end | mma_first  mma_issued  wait_w  wait_a | epi_dur  mma_span")
    for g in range(4 * L):
        r = [int(v) for v in t[g]]
        print("%3d %10d %10d %10d | %10d %10d %7d %7d | %7d %7d" % (
            g, r[7] - base, r[0] - base, r[2] - base, r[3] - base, r[4] - base, r[5], r[6], r[2] - r[0], r[4] - r[3]))
    for name, off in (("forward stage G=1", 64), ("reverse stage G=8", 72)):
        k = raw[off:off + 8]
        print(name, "(thread 0): per k-block  ld_wait  math+st.shared  fences  syncwarp+arrive  -  global st/ld tail | total")
        for kb in range(8):
            r = [int(v) for v in k[kb]]
            print("   kb %d: start +%6d | %5d %5d %5d %5d %5d %5d | %6d" % (
                kb, r[0] - int(t[1 if off == 64 else 8, 0]), r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3], 0,
                r[5] - r[4], r[5] - r[0]))
    warps(raw, t)


def warps(raw, t):
    """per-warp stamps (lane 0 of the 16 epilogue warps) of stages G = 1 / G = 8, relative to thread 0's acc_full"""
    for name, off, g in (("forward stage G=1", 1024, 1), ("reverse stage G=8", 1088, 8)):
        a0 = int(t[g, 0])
        w = raw[off // 8:(off + 64) // 8].reshape(-1)[:64].reshape(16, 4)
        print(name, "per warp: acc_full seen / k-block 0 published / k-block 7 published (cycles after thread 0 saw acc_full)")
        print("   " + "  ".join("w%d:%d/%d/%d" % (i, int(w[i, 0]) - a0, int(w[i, 1]) - a0, int(w[i, 2]) - a0) for i in range(16)))


if __name__ == "__main__":
    main()

