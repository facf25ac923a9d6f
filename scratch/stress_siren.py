"""Stress: big / odd sizes and other network scales through the fused projection path vs fp64 autograd."""
import sys
import torch
sys.path.insert(0, ".")
from tests.helpers import Siren
from isopoints_b200 import siren
from isopoints_b200.levelset_sampling import UniformProjection
dev = "cuda"


def ref64(model, L, om, x):
    m = Siren(256, L, om).double()
    m.load_state_dict({k: v.double().cpu() for k, v in model.state_dict().items()})
    m = m.to(dev)
    out_s, out_g = [], []
    for xs in torch.split(x, 200000):
        xx = xs.double().clone().requires_grad_(True)
        s = m(xx).sdf
        g, = torch.autograd.grad(s, xx, torch.ones_like(s))
        out_s.append(s.detach().reshape(-1)); out_g.append(g.detach())
    return torch.cat(out_s), torch.cat(out_g)


for L, om, wmul, n in ((7, 30.0, 1.0, 1_000_003), (5, 45.0, 1.5, 300_001), (3, 10.0, 0.3, 77)):
    model = Siren(256, L, om, seed=L).to(dev)
    with torch.no_grad():
        for lyr in list(model.net)[1:-1]:
            lyr.linear.weight.mul_(wmul)
    x = ((torch.rand(n, 3, device=dev) - 0.5) * 3).contiguous()
    s, g = siren.sdf_and_grad(model, x)
    s64, g64 = ref64(model, L, om, x)
    print("L=%d omega=%g wmul=%g n=%d: sdf err %.3e (max |sdf| %.3f)  grad err %.3e (max |grad| %.3f)  nan %d" % (
        L, om, wmul, n, (s.double() - s64).abs().max().item(), s64.abs().max().item(),
        (g.double() - g64).abs().max().item(), g64.abs().max().item(), int(torch.isnan(s).sum() + torch.isnan(g).sum())))
    out = UniformProjection(proj_max_iters=10).project_points(x[None], model, skip_upsampling=True)
    pts = out["levelset_points"][0][out["mask"][0]]
    s2, _ = ref64(model, L, om, pts[:200000])
    print("   projected+resampled %d -> %d valid; |sdf| at results (fp64): median %.2e max %.2e" % (
        n, pts.shape[0], s2.abs().median().item(), s2.abs().max().item()))
