"""Timeline of one C2 step (torch.profiler / CUPTI): kernels in launch order with start offsets,
so the GPU idle gaps (host syncs, Python glue) become visible."""
import sys
import torch

sys.path.insert(0, ".")
from bench import _make_c2  # noqa: E402
from isopoints_b200.levelset_sampling import UniformProjection  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
x, net = _make_c2(0, dev)
net = net.to(dev)
x = x.to(dev)
proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
for _ in range(3):
    proj.project_points(x, net, skip_upsampling=True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    proj.project_points(x, net, skip_upsampling=True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
last_end = t0
busy = 0.0
print("%9s %8s %8s  %s" % ("start_us", "dur_us", "gap_us", "kernel"))
for e in evs:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = e.time_range.start - last_end
    busy += d
    print("%9.1f %8.1f %8.1f  %s" % (s, d, gap, e.name[:90]))
    last_end = max(last_end, e.time_range.end)
print("span %.1f us, busy %.1f us, idle %.1f us, %d kernels" % (last_end - t0, busy, last_end - t0 - busy, len(evs)))
