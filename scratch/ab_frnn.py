import sys, torch, json
sys.path.insert(0, '.')
from isopoints_b200 import frnn, _ext
def t(p, K, r, mode, n=10):
    frnn.QUERY_MODE = mode
    lens = torch.tensor([p.shape[1]], device='cuda')
    rr = torch.tensor([r], device='cuda')
    for _ in range(3): frnn.frnn_grid_points(p, p, lens, lens, K=K, r=rr)
    torch.cuda.synchronize(); _ext.PROFILE = {}
    for _ in range(n): frnn.frnn_grid_points(p, p, lens, lens, K=K, r=rr)
    torch.cuda.synchronize(); pr, _ext.PROFILE = _ext.PROFILE, None
    v = pr['isob200_frnn_find_nbrs']; return sum(a.elapsed_time(b) for a, b in v) / len(v)
g = torch.Generator().manual_seed(0)
box = torch.rand(1, 500_000, 3, generator=g).cuda()
sph = torch.nn.functional.normalize(torch.randn(1, 500_000, 3, generator=g), dim=-1).cuda()
sph200 = torch.nn.functional.normalize(torch.randn(1, 200_000, 3, generator=g), dim=-1).cuda()
import math
diag = (sph200[0].max(0).values - sph200[0].min(0).values).norm().item()
r_c2 = math.sqrt(diag / 200_000) * 8
for name, p, K, r in (('box500k_K16', box, 16, 0.05), ('sphere500k_K16', sph, 16, 0.05), ('sphere200k_K9_c2radius', sph200, 9, r_c2),
                      ('box100k_K8', box[:, :100_000].contiguous(), 8, 0.05), ('sphere500k_K9', sph, 9, 0.02)):
    print(name, 'r=%.4f' % r, {m: round(t(p, K, r, m), 4) for m in (1, 2)}, flush=True)
