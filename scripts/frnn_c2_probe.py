#!/usr/bin/env python
"""Developer tool: the FRNN query of BASELINE config 2's resample (iso-points of the pinned SIREN, K = 9,
r = sqrt(diag / n) * 8) in the three traversal modes, plus the statistics that decide between them: points per cell,
candidates in a query's (2c+1)^3 block, neighbours within r.  Not part of the product or the tests."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isopoints_b200 import frnn  # noqa: E402
from isopoints_b200.levelset_sampling import UniformProjection  # noqa: E402
from tests.helpers import pinned_siren  # noqa: E402


def main():
    dev = "cuda"
    net = pinned_siren(0).to(dev)
    g = torch.Generator().manual_seed(1000)
    x = ((torch.rand(1, 200_000, 3, generator=g) - 0.5) * 2).to(dev)
    proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    out = proj.project_points(x, net, skip_resampling=True, skip_upsampling=True)
    p = out["levelset_points"][out["mask"]][None].contiguous()
    n = p.shape[1]
    mn, mx = torch.aminmax(p, dim=1)
    diag = (mx - mn).norm(dim=-1)
    r = float(torch.sqrt(diag / n) * 8)
    print("points", n, "r", r)
    lens = torch.tensor([n], device=dev)
    res = {}
    for mode in (1, 2, 3):
        frnn.QUERY_MODE = mode
        for _ in range(3):
            d, i, _, grid = frnn.frnn_grid_points(p, p, lens, lens, K=9, r=r)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        grid_t = tuple(grid)
        a.record()
        for _ in range(10):
            frnn.find_nbrs(grid.sorted_points2, lens, lens, grid, 9, torch.tensor([r], device=dev),
                           q_points=grid.sorted_points2, q_order=grid.sorted_points2_idxs)
        b.record()
        torch.cuda.synchronize()
        print("mode %d: query %.3f ms" % (mode, a.elapsed_time(b) / 10))
        res[mode] = (d, i)
    frnn.QUERY_MODE = 0
    print("modes equal:", all(torch.equal(res[1][k], res[m][k]) for m in (2, 3) for k in (0, 1)))
    prm = grid.grid_params[0].cpu().numpy()
    off = grid.pc2_grid_off[0].cpu().numpy().astype(np.int64)
    delta = prm[3]
    rx, ry, rz = int(prm[4]), int(prm[5]), int(prm[6])
    G = int(prm[7])
    cnt = np.diff(np.concatenate([off[:G], [n]]))
    print("cell size %.4f (r / cell = %.2f), grid %d x %d x %d = %d cells, %.3f points per cell, %.1f %% occupied"
          % (1 / delta, r * delta, rx, ry, rz, G, n / G, 100 * (cnt > 0).mean()))
    vol = cnt.reshape(rx, ry, rz)
    # block counts by a box filter of half-width c = ceil(r * delta)
    c = int(np.ceil(r * delta))
    cs = np.pad(vol, c).cumsum(0).cumsum(1).cumsum(2)
    cs = np.pad(cs, ((1, 0), (1, 0), (1, 0)))
    w = 2 * c + 1
    blk = (cs[w:, w:, w:] - cs[:-w, w:, w:] - cs[w:, :-w, w:] - cs[w:, w:, :-w] + cs[:-w, :-w, w:] + cs[:-w, w:, :-w]
           + cs[w:, :-w, :-w] - cs[:-w, :-w, :-w])
    per_query = np.repeat(blk.reshape(-1), cnt)
    print("candidates in a query's %d^3 block: mean %.1f, median %.0f, 90 %% %.0f, 99 %% %.0f, > 48: %.1f %% of queries"
          % (w, per_query.mean(), np.median(per_query), np.percentile(per_query, 90), np.percentile(per_query, 99),
             100 * (per_query > 48).mean()))
    frnn.QUERY_MODE = 1
    d32 = frnn.frnn_grid_points(p, p, lens, lens, K=32, r=r)[0]
    frnn.QUERY_MODE = 0
    within = (d32[0] >= 0).sum(-1).float()
    print("neighbours within r (capped at 32): mean %.1f, median %.0f, == 32: %.1f %%"
          % (float(within.mean()), float(within.median()), 100 * float((within == 32).float().mean())))


    # emulate the collect kernel's trial-radius logic on a sample of queries
    K, CAP = 9, 48
    lam = min(K + 4 * K ** 0.5, CAP - 8)
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(0))[:4000].to(dev)
    q = p[0, sel]
    d2 = torch.cdist(q.double(), p[0].double()).pow(2).float()          # (4000, n)
    rel = (q - torch.as_tensor(prm[:3], device=dev)) * float(delta)
    cell = rel.floor().long().clamp_min(0)
    cell[:, 0].clamp_(max=rx - 1); cell[:, 1].clamp_(max=ry - 1); cell[:, 2].clamp_(max=rz - 1)
    nb = torch.as_tensor(blk, device=dev)[cell[:, 0], cell[:, 1], cell[:, 2]].float()
    lo = (rel - r * float(delta)).floor().clamp_min(0)
    hi = torch.minimum((rel + r * float(delta)).floor(), torch.tensor([rx - 1, ry - 1, rz - 1], device=dev).float())
    cells = (hi - lo + 1).prod(-1)
    vol = lam * cells / nb
    rc = (vol * 0.238732415).pow(1 / 3)
    res_t = torch.tensor([rx, ry, rz], device=dev).float()
    t = torch.minimum(rel, res_t - rel)
    inside = (0.5 * (t / rc[:, None] + 1)).clamp(0.5, 1.0).prod(-1)
    rc = rc * (1 / inside).pow(1 / 3)
    tau = torch.minimum(torch.full_like(rc, r * r), (rc / float(delta)) ** 2)
    tau = torch.where(nb > CAP, tau, torch.full_like(tau, r * r))
    exact = torch.zeros_like(tau, dtype=torch.bool)
    hist = []
    for attempt in range(3):
        cnt = (d2 <= tau[:, None]).sum(-1).float()
        over = (cnt > CAP) & ~exact
        under = (cnt < K) & (tau < r * r) & ~exact
        ok = ~over & ~under & ~exact
        hist.append((int(ok.sum()), int(over.sum()), int(under.sum()), float(cnt[~exact].median()) if (~exact).any() else 0))
        exact |= ok
        tau = torch.where(over, tau * (lam / cnt).pow(0.8), tau)
        tau = torch.where(under, torch.minimum(torch.full_like(tau, r * r), tau * (1.5 * lam / cnt.clamp_min(1)).pow(0.8)), tau)
    print("trial-radius emulation on 4000 queries: per attempt (accepted, overflow, short, median count of the open ones):", hist,
          "-> fallback %.1f %%" % (100 * float((~exact).float().mean())))
    kth = d2.kthvalue(K, dim=-1).values
    print("K-th neighbour distance / r: median %.3f; first trial radius / r: median %.3f"
          % (float((kth.sqrt() / r).median()), float(((rc / float(delta)) / r).median())))


if __name__ == "__main__":
    main()
