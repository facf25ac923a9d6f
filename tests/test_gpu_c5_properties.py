"""BASELINE configs[4] ("C5": 2 000 000 points, project + resample + splat at 1024^2) at FULL size on one GPU.

No oracle finishes two million points in seconds, so the checks are properties that do not depend on the size
(the small-size parity against the oracle / the reference's own kernels lives in the other test files):

  * projection: every row flagged converged satisfies the loop's own criterion when its SDF is evaluated again,
    its normal is the SDF gradient there, and projecting the result again changes nothing (idempotence);
  * FRNN on the 1.2 M survivors with the resample tree's radius: per row the distances ascend, lie inside r^2, the
    indices are distinct, the row itself comes first at distance 0, and the traversal kernels agree bit for bit;
  * splat of the survivors, 2 views x 1024^2: depth-sorted, distinct ids per pixel, -1 padding only at the tail,
    occupancy = first slot taken; the backward is positively homogeneous in the incoming image gradients and its
    depth part linear; rendering the views together equals rendering them one at a time (the view sharding of the
    multi-GPU path).
"""
import math

import pytest
import torch

from isopoints_b200 import frnn, siren, splat
from isopoints_b200.levelset_sampling import UniformProjection
from tests.helpers import pinned_siren

pytestmark = pytest.mark.gpu
DEV = "cuda"
N_POINTS = 2_000_000
TOL = 5e-5
S, K = 1024, 8


@pytest.fixture(scope="module")
def c5():
    model = pinned_siren(0).to(DEV)
    g = torch.Generator().manual_seed(7)                      # bench_c5.py's cloud
    x = ((torch.rand(N_POINTS, 3, generator=g) - 0.5) * 2).to(DEV)[None]
    proj = UniformProjection(proj_max_iters=10, proj_tolerance=TOL, knn_k=8, sample_iters=1)
    out = proj.project_points(x, model, skip_upsampling=True)
    mask = out["mask"][0]
    pts = out["levelset_points"][0][mask].contiguous()
    nrm = out["levelset_normals"][0][mask].contiguous()
    assert 1_000_000 < pts.shape[0] < N_POINTS
    return model, proj, pts, nrm


def test_converged_rows_meet_the_criterion_and_reprojection_is_idempotent(c5):
    model, proj, pts, nrm = c5
    sdf, grad = siren.sdf_and_grad(model, pts)
    assert float(sdf.abs().max()) < TOL                       # the flag was set from this very evaluation
    assert torch.allclose(nrm, grad, rtol=1e-6, atol=1e-7)
    n = torch.tensor([pts.shape[0]], device=DEV)
    again = proj._project_points(model, pts[None], n, proj_max_iters=10, num_points_list=[pts.shape[0]])
    assert bool(again.mask.all())
    assert torch.equal(again.points[0], pts)                  # converged at iteration 0: not moved


def test_frnn_on_the_survivors_ordered_distinct_self_first_and_modes_agree(c5):
    _, _, pts, _ = c5
    n = pts.shape[0]
    diag = (pts.max(0).values - pts.min(0).values).norm()
    r = (torch.sqrt(diag / n) * 8).reshape(1)                 # levelset_sampling.py:128-131
    lens = torch.tensor([n], device=DEV)
    res = {}
    old = frnn.QUERY_MODE
    try:
        for mode in (1, 2):                                   # exhaustive / pruned group kernels
            frnn.QUERY_MODE = mode
            d, i, _, _ = frnn.frnn_grid_points(pts[None], pts[None], lens, lens, K=9, r=r)
            res[mode] = (d[0], i[0])
    finally:
        frnn.QUERY_MODE = old
    d, i = res[1]
    assert torch.equal(d, res[2][0]) and torch.equal(i, res[2][1])
    found = i >= 0
    assert bool(found[:, 0].all())
    big = torch.full_like(d, float("inf"))
    dd = torch.where(found, d, big)
    assert bool((dd[:, 1:] >= dd[:, :-1]).all())              # ascending, the -1 padding at the tail
    assert bool((d[found] <= r * r).all()) and bool((d[found] >= 0).all())
    assert bool((d[:, 0] == 0).all())
    rows = torch.arange(n, device=DEV)
    dup = d[:, 1] == 0                                        # coincident points: the lower index comes first
    assert bool((i[:, 0] == rows)[~dup].all())
    srt = torch.where(found, i, torch.arange(-9, 0, device=DEV)[None].expand_as(i)).sort(dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())            # distinct per row


def _views(pts, views, n_views=16):
    scr = []
    for v in views:
        a = 2 * math.pi * v / n_views
        b = 0.35 * math.sin(3 * a)
        ca, sa, cb, sb = math.cos(a), math.sin(a), math.cos(b), math.sin(b)
        ry = torch.tensor([[ca, 0, sa], [0, 1, 0], [-sa, 0, ca]], device=DEV)
        rx = torch.tensor([[1, 0, 0], [0, cb, -sb], [0, sb, cb]], device=DEV)
        q = pts @ (rx @ ry).T
        scr.append(torch.stack([q[:, 0] * 0.45, q[:, 1] * 0.45, q[:, 2] + 3.0], dim=1))
    return torch.cat(scr, 0).contiguous()


def _render(pts, views, occ_grad=None, z_grad=None):
    P, nv = pts.shape[0], len(views)
    scr = _views(pts, views).requires_grad_(occ_grad is not None)
    sig = 1.5 * 2.0 / S
    ell = torch.tensor([1 / sig ** 2, 0.0, 1 / sig ** 2], device=DEV).expand(nv * P, 3).contiguous()
    radii = torch.full((nv * P, 2), sig, device=DEV)
    first = torch.arange(nv, device=DEV, dtype=torch.int64) * P
    num = torch.full((nv,), P, device=DEV, dtype=torch.int64)
    idx, zbuf, qv, occ = splat.EllipticalRasterizer.apply(scr, ell, torch.ones(nv * P, device=DEV), radii, first, num,
                                                          0.05, S, K, 64, 0, 10.0)
    if occ_grad is None:
        return idx, zbuf, qv, occ
    torch.autograd.backward([occ, zbuf], [occ_grad, z_grad])
    return scr.grad


def test_splat_of_the_survivors_pixel_lists_linearity_and_view_sharding(c5):
    _, _, pts, _ = c5
    P = pts.shape[0]
    with torch.no_grad():
        idx, zbuf, qv, occ = _render(pts, [0, 5])
        taken = idx >= 0
        assert bool((taken[..., 1:] <= taken[..., :-1]).all())            # -1 only at the tail
        assert bool(((zbuf[..., 1:] >= zbuf[..., :-1]) | ~taken[..., 1:]).all())
        assert bool((occ == taken[..., 0].float()).all())
        assert bool((idx[taken] < 2 * P).all()) and bool((qv[taken] >= 0).all())
        srt = torch.where(taken, idx, torch.arange(-K, 0, device=DEV).expand_as(idx)).sort(dim=-1).values
        assert bool((srt[..., 1:] != srt[..., :-1]).all())                # distinct ids per pixel
        assert bool((idx[0][taken[0]] < P).all()) and bool((idx[1][taken[1]] >= P).all())
        # views one at a time: the same pixels (ids are offsets into the packed views)
        for k, v in enumerate((0, 5)):
            i1, z1, q1, o1 = _render(pts, [v])
            assert torch.equal(torch.where(i1 >= 0, i1 + k * P, i1), idx[k][None])
            assert torch.equal(z1, zbuf[k][None]) and torch.equal(q1, qv[k][None]) and torch.equal(o1, occ[k][None])
    g = torch.Generator().manual_seed(11)
    og = [(torch.randn(2, S, S, generator=g) * (torch.rand(2, S, S, generator=g) < 0.1)).to(DEV) for _ in range(2)]
    zg = [torch.randn(2, S, S, K, generator=g).to(DEV) for _ in range(2)]
    # the occupancy gradient depends on the SIGN of the incoming gradient (rasterize_points_backward.cu:163-172), so
    # it is positively homogeneous, not linear; the depth gradient (rasterize_points.cu:835-843) is linear
    ga = _render(pts, [0, 5], og[0], zg[0])
    g2 = _render(pts, [0, 5], 2.0 * og[0], 2.0 * zg[0])
    scale = ga.abs().amax(0).clamp_min(1e-20)
    assert torch.allclose(g2 / scale, 2.0 * ga / scale, rtol=1e-4, atol=2e-5)
    zero = torch.zeros_like(og[0])
    za = _render(pts, [0, 5], zero, zg[0])
    zb = _render(pts, [0, 5], zero, zg[1])
    zc = _render(pts, [0, 5], zero, 2.0 * zg[0] - 0.5 * zg[1])
    assert float(za[:, :2].abs().max()) == 0.0                            # depth gradients move z only
    ref = 2.0 * za - 0.5 * zb
    zs = ref.abs().amax(0).clamp_min(1e-20)
    assert torch.allclose(zc / zs, ref / zs, rtol=1e-4, atol=2e-5)
    assert float(ga[:, :2].abs().sum()) > 0 and float(za[:, 2].abs().sum()) > 0
