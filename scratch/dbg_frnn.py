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
bad = (d != rd)
print('d mismatches', bad.sum().item(), 'rows', bad.any(-1).sum().item(), 'idx mismatches', (i != ri).sum().item())
rows = bad.any(-1)[0].nonzero().squeeze(1)[:5]
torch.set_printoptions(precision=10, linewidth=200)
for rr in rows.tolist():
    print('row', rr); print(d[0, rr]); print(rd[0, rr]); print(i[0, rr]); print(ri[0, rr])
    # brute force for this row
    q = p[0, rr]
    diff = p[0] - q
    dd = diff[:, 0] * diff[:, 0]
    dd = torch.addcmul(dd, diff[:, 1], diff[:, 1]); dd = torch.addcmul(dd, diff[:, 2], diff[:, 2])
    v, ix = torch.topk(dd, 18, largest=False)
    print('bf', v); print(ix)
# run ours twice: deterministic?
d2, i2, _, _ = frnn.frnn_grid_points(p, p, lens, lens, K=16, r=0.05)
print('self-consistent', torch.equal(d, d2), torch.equal(i, i2))
ri2, rd2, *_ = ref_native.frnn_grid_points_cuda(p, p, lens, lens, 16, r)
print('ref self-consistent', torch.equal(rd, rd2), torch.equal(ri, ri2))
