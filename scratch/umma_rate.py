"""MMA issue-rate microbenchmark (csrc/umma2_probe.cu): cycles per N=256, K=16 fp16 MMA."""
import sys
import torch
sys.path.insert(0, ".")
from isopoints_b200 import _ext
lib = _ext.lib()
dev = torch.device("cuda")
names = {3: "cta_group::1 M=64", 0: "cta_group::1 M=128", 1: "cta_group::2 M=128 (64 rows/CTA)", 2: "cta_group::2 M=256 (128 rows/CTA)"}
for mode in (0, 1000, 2000, 1100, 2100, 1003, 2003, 1, 2, 3, 10, 100, 200, 300, 103, 203, 102, 202):
    for reps in (64, 512):
        c = torch.zeros(2, dtype=torch.int64, device=dev)
        for _ in range(2):
            _ext.check(lib.isob200_umma_rate(mode, reps, _ext.ptr(c), _ext.stream(dev)))
        torch.cuda.synchronize()
        v = c.tolist()
        print("N=%3d %-52s %s reps %4d: %7d cycles -> %.1f cycles / MMA" % (256 >> ((mode // 100) % 10), names[mode % 10] + ("", " 2 accumulators", " A in TMEM")[mode // 1000], "SW128     " if (mode // 10) % 10 else "no-swizzle", reps, max(v), max(v) / reps))
