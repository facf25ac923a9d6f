import sys, types, torch, numpy as np
sys.path.insert(0, '.')
from oracle import ref_python
import tests.golden.make_golden as mg
ref = ref_python.load(frnn_module=mg._CpuFrnn)
LS = ref.levelset_sampling
torch.manual_seed(0)
P, K = 500, 15
pts = torch.nn.functional.normalize(torch.randn(1, P, 3), dim=-1)
nrm = torch.nn.functional.normalize(pts + 0.05 * torch.randn(1, P, 3), dim=-1)
d = torch.cdist(pts[0], pts[0]) ** 2
v, i = torch.topk(d, K + 1, largest=False)
idx = i[None, :, 1:]
knn_pts = pts[0][idx[0]][None]
knn_normals = nrm[0][idx[0]][None]
# reference formula (levelset_sampling.py:611-628)
mid_points = (knn_pts + 2 * pts[..., None, :]) / 3
mid_nn_diff = mid_points.unsqueeze(-2) - knn_pts.unsqueeze(-3)
dot_product = (2 - torch.sum(nrm.unsqueeze(-2) * knn_normals, dim=-1)) ** 1
min_dist2 = torch.norm(mid_nn_diff, dim=-1)
min_dist2 = min_dist2 - torch.sum((mid_nn_diff * knn_normals.unsqueeze(-2)) ** 2, dim=-1)
print('shapes', mid_nn_diff.shape, knn_normals.unsqueeze(-2).shape)
min_dist2 = min_dist2.min(dim=-1)[0]
min_dist2 = torch.clamp(min_dist2.abs(), 1e-17).sqrt()
fs, fnb = (dot_product * min_dist2).max(dim=-1)
# my kernel's formula
best = []
for k in range(K):
    mk = mid_points[0, :, k]            # (P,3)
    diff = mk[:, None, :] - knn_pts[0]   # (P,K,3) over j
    val = diff.norm(dim=-1) - ((diff * knn_normals[0][:, k:k+1]) ** 2).sum(-1)
    best.append(val.min(-1)[0])
best = torch.stack(best, -1)
mine = torch.clamp(best.abs(), 1e-17).sqrt() * (2 - (nrm[0][:, None] * knn_normals[0]).sum(-1))
ms, mnb = mine.max(-1)
print('score close', torch.allclose(ms, fs[0], rtol=1e-5), 'father same', (mnb == fnb[0]).float().mean().item())
print('max abs diff', (ms - fs[0]).abs().max().item(), 'rel', ((ms - fs[0]).abs() / fs[0].abs()).max().item())
full_ref = (dot_product * min_dist2)[0]
print('full close', torch.allclose(mine, full_ref, rtol=1e-4, atol=1e-7), (mine - full_ref).abs().max().item())
bad = (mnb != fnb[0]).nonzero().squeeze(1)[:3]
for b in bad.tolist():
    print(b, mine[b], full_ref[b])
