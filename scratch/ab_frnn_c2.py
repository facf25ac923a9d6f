"""A/B of the FRNN traversal modes on the C2 resample query (K = 9 over ~197k projected points)."""
import sys
import torch
sys.path.insert(0, ".")
from bench import _make_c2
from isopoints_b200 import frnn, _ext
from isopoints_b200.levelset_sampling import UniformProjection
dev = torch.device("cuda", 0)
x, net = _make_c2(0, dev)
net, x = net.to(dev), x.to(dev)
proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
for mode in (0, 1, 2, 0, 1, 2):
    frnn.QUERY_MODE = mode
    for _ in range(2):
        proj.project_points(x, net, skip_upsampling=True)
    _ext.PROFILE = {}
    for _ in range(5):
        out = proj.project_points(x, net, skip_upsampling=True)
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in _ext.PROFILE["isob200_frnn_find_nbrs"]]
    _ext.PROFILE = None
    print("mode %d: find_nbrs %.3f ms  (checksum %.6f)" % (mode, sum(ts) / len(ts), float(out["levelset_points"].double().sum())))
