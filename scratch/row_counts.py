"""Live rows entering each Newton iteration of the C2 step (first projection, then the re-projection)."""
import sys
import torch
sys.path.insert(0, ".")
from bench import _make_c2
from isopoints_b200 import siren
from isopoints_b200.levelset_sampling import UniformProjection
dev = torch.device("cuda", 0)
x, net = _make_c2(0, dev)
net, x = net.to(dev), x.to(dev)
proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
siren.RECORD = []
proj.project_points(x, net, skip_upsampling=True)
torch.cuda.synchronize()
for kind, cnt in siren.RECORD:
    print(kind, cnt.tolist())
