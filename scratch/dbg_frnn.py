import torch, sys
sys.path.insert(0, '.')
from isopoints_b200 import frnn
from oracle import ref_native
g = torch.Generator().manual_seed(0)
p = torch.rand(1, 500_000, 3, generator=g).cuda()
lens = torch.tensor([500_000], device='cuda')
r = torch.tensor([0.05], device='cuda')
d, i, _, grid = frnn.frnn_grid_points(p, p, lens, lens, K=16, r=0.05)
ri, rd, *_ = ref_native.frnn_grid_points_cuda(p, p, lens, lens, 16, r)
print(d[0, :3]); print(i[0, :3]); print(rd[0, :3]); print(ri[0, :3])
print('idx equal frac', (i == ri).float().mean().item(), 'd equal', (d == rd).float().mean().item())
ties = (d[0, :, 1:] == d[0, :, :-1]) & (d[0, :, 1:] >= 0)
rt = (rd[0, :, 1:] == rd[0, :, :-1]) & (rd[0, :, 1:] >= 0)
print('ties ours', ties.sum().item(), 'ref', rt.sum().item())
print('unique coords', torch.unique(p[0], dim=0).shape)
