#!/usr/bin/env python
"""Developer tool: what does the SIREN kernel's weight stream (L2 -> shared memory, 3.6 MB per 128-row tile) cost
under the power cap?  Runs the bring-up build of the kernel back to back for a few seconds with one ingredient
removed at a time (dbg_gemm = -3: no TMA weight copies after the first pass over the ring; -5: no tape stores / loads;
-6: only the hi*hi product of the three -- results are wrong in these three, everything else is identical; -7: the
hidden layers' sines by MUFU.SIN on the reduced argument instead of the degree-11 polynomial -- results 4e-7 abs off) and prints evaluations/s, SM clock and board power.  Not part of the product or the tests."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isopoints_b200 import siren  # noqa: E402
from tests.helpers import pinned_siren  # noqa: E402


def main():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = "cuda"
    model = pinned_siren(0).to(dev)
    n = 148 * 128 * 8
    x = ((torch.rand(n, 3, device=dev) - 0.5) * 2).contiguous()
    for tag, flag in (("full kernel                  ", -4), ("without the weight stream    ", -3),
                      ("without the tape traffic     ", -5), ("one fp16 product out of three", -6),
                      ("hidden-layer sines on the SFU", -7),
                      ("full kernel                  ", -4), ("without the weight stream    ", -3),
                      ("without the tape traffic     ", -5), ("one fp16 product out of three", -6),
                      ("hidden-layer sines on the SFU", -7)):
        t_end = time.time() + 4.0
        calls = 0
        clk, pw = [], []
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        while time.time() < t_end:
            for _ in range(20):
                siren.sdf_and_grad(model, x, dbg_gemm=flag)
            calls += 20
            torch.cuda.synchronize()
            if time.time() > t_end - 2.0:
                clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / calls
        print("%s: %.4f ms per %d rows, %.1f M evals/s, SM %d MHz, %.0f W" % (
            tag, ms, n, n / ms / 1e3, sorted(clk)[len(clk) // 2], sorted(pw)[len(pw) // 2]))


if __name__ == "__main__":
    main()
