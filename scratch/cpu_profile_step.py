"""Host-side cost of one C2 step (cProfile, top cumulative entries)."""
import cProfile, pstats, sys, io
import torch
sys.path.insert(0, ".")
from bench import _make_c2
from isopoints_b200.levelset_sampling import UniformProjection
dev = torch.device("cuda", 0)
x, net = _make_c2(0, dev)
net, x = net.to(dev), x.to(dev)
proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
for _ in range(5):
    proj.project_points(x, net, skip_upsampling=True)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    proj.project_points(x, net, skip_upsampling=True)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:5000])
