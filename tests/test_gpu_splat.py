"""GPU parity: elliptical splat forward / backward / blend through the C ABI vs the oracle
(oracle/port.py) and the reference's own CUDA kernels (oracle/_ref/ref_dss_C.so)."""
import numpy as np
import pytest
import torch

from isopoints_b200 import splat
from isopoints_b200.structures import Pointclouds
from oracle import port, ref_native
from tests.helpers import make_splat_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"
REF = pytest.mark.skipif(not ref_native.available(), reason="oracle/_ref not built")


def _t(inp):
    return {k: torch.as_tensor(v, device=DEV) for k, v in inp.items()}


def _fwd(t, S, K, thres=0.05, bin_size=16):
    return splat._C.splat_points(t["points"], t["ellipse"], t["cutoff"], t["radii"], t["first_idx"],
                                 t["num_points"], thres, S, K, bin_size, 10000)


@pytest.fixture(params=[1, 2, 0], ids=["raster_v1", "raster_v2_two_ctas", "raster_v2_default"])
def raster_variant(request):
    old = splat.RASTER_VARIANT
    splat.RASTER_VARIANT = request.param
    yield request.param
    splat.RASTER_VARIANT = old


def test_forward_dense_overdraw_takes_the_overflow_path(raster_variant):
    """~60 covering splats per pixel: far beyond the per-pixel column capacity of raster v2."""
    S, K = 32, 8
    inp = make_splat_inputs(1, 9000, S, seed=23, sigma_px=1.5, behind_frac=0.0)
    idx, zbuf, qv, occ = _fwd(_t(inp), S, K)
    wi, wz, wq, wo = port.splat_forward(inp["points"], inp["ellipse"], inp["cutoff"], inp["radii"],
                                        inp["first_idx"], inp["num_points"], 0.05, S, K)
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(zbuf.cpu().numpy(), wz)
    assert np.array_equal(qv.cpu().numpy(), wq) and np.array_equal(occ.cpu().numpy(), wo)


@pytest.mark.parametrize("S,K", [(64, 1), (64, 3), (64, 4), (50, 5), (64, 8), (48, 16), (40, 20)])
def test_forward_bit_exact_vs_oracle(S, K, raster_variant):
    inp = make_splat_inputs(2, [1500, 900], S, seed=S + K, sigma_px=1.7)
    idx, zbuf, qv, occ = _fwd(_t(inp), S, K)
    wi, wz, wq, wo = port.splat_forward(inp["points"], inp["ellipse"], inp["cutoff"], inp["radii"],
                                        inp["first_idx"], inp["num_points"], 0.05, S, K)
    assert idx.dtype == torch.int32 and idx.shape == (2, S, S, K) and occ.shape == (2, S, S)
    assert np.array_equal(idx.cpu().numpy(), wi)
    assert np.array_equal(zbuf.cpu().numpy(), wz)
    assert np.array_equal(qv.cpu().numpy(), wq)          # same fp32 expression
    assert np.array_equal(occ.cpu().numpy(), wo)
    assert (wi[..., 0] >= 0).mean() > 0.3
    t = _t(inp)
    n_pairs = len(port.splat_pairs(inp["points"], inp["ellipse"], inp["cutoff"], inp["radii"], inp["first_idx"],
                                   inp["num_points"], S)[0])
    assert splat.count_pixel_splats(t["points"], t["ellipse"], t["cutoff"], t["radii"], S) == n_pairs


def test_forward_max_points_per_pixel_150():
    """kMaxPointsPerPixel = 150 (rasterization_utils.cuh:18): the large-K path, dense overlap."""
    S, K = 24, 150
    inp = make_splat_inputs(1, 4000, S, seed=2, sigma_px=2.5, behind_frac=0.0)
    idx, zbuf, qv, occ = _fwd(_t(inp), S, K, thres=10.0)
    wi, wz, wq, wo = port.splat_forward(inp["points"], inp["ellipse"], inp["cutoff"], inp["radii"],
                                        inp["first_idx"], inp["num_points"], 10.0, S, K)
    assert (wi >= 0).sum(-1).max() > 100                      # well past the register-list kernels
    assert np.array_equal(idx.cpu().numpy(), wi) and np.array_equal(zbuf.cpu().numpy(), wz)
    assert np.array_equal(qv.cpu().numpy(), wq) and np.array_equal(occ.cpu().numpy(), wo)


def test_forward_naive_occupancy_rule_and_z0():
    """bin_size == 0 is the naive kernel (occupied when z >= 0), otherwise z > 0 (cu:196 vs :581)."""
    S, K = 32, 4
    inp = make_splat_inputs(1, 300, S, seed=3, behind_frac=0.0)
    inp["points"][:, 2] = 0.0
    t = _t(inp)
    i0, _, _, o0 = _fwd(t, S, K, bin_size=0)
    i1, _, _, o1 = _fwd(t, S, K, bin_size=8)
    assert torch.equal(i0, i1) and (i0[..., 0] >= 0).any()
    assert torch.equal(o0.bool(), i0[..., 0] >= 0) and not o1.any()
    # equal z everywhere: ties resolved by ascending point id
    valid = i0 >= 0
    assert ((i0[..., 1:] > i0[..., :-1]) | ~valid[..., 1:]).all()


def test_forward_edge_cases(raster_variant):
    S, K = 32, 4
    t = _t(make_splat_inputs(2, [200, 0], S, seed=1))
    idx, zbuf, qv, occ = _fwd(t, S, K)
    assert (idx[1] == -1).all() and (zbuf[1] == -1).all() and (qv[1] == -1).all() and (occ[1] == 0).all()
    # everything behind the camera / one splat covering the whole frame / no points at all
    inp = make_splat_inputs(1, 50, S, seed=2)
    inp["points"][:, 2] = -1.0
    assert (_fwd(_t(inp), S, K)[0] == -1).all()
    inp = make_splat_inputs(1, 1, S, seed=2, aniso=False, behind_frac=0)
    inp["points"][0] = [0.1, -0.2, 2.0]
    inp["radii"][:] = 5.0
    inp["ellipse"][:] = [1e-3, 0.0, 1e-3]
    idx = _fwd(_t(inp), S, K)[0]
    assert (idx[..., 0] == 0).all() and (idx[..., 1:] == -1).all()
    e = {k: v[:0] if k not in ("first_idx", "num_points") else np.zeros(1, np.int64) for k, v in inp.items()}
    idx, _, _, occ = _fwd(_t(e), S, K)
    assert (idx == -1).all() and (occ == 0).all()
    with pytest.raises(RuntimeError):
        _fwd(_t(inp), S, 151)
    with pytest.raises(RuntimeError):
        _fwd(_t(inp), 512, K, bin_size=8)           # 64 bins >= 22 (rasterize_points.cu:462)
    with pytest.raises(TypeError):
        splat._C.splat_points(*[torch.as_tensor(inp[k]) for k in
                                ("points", "ellipse", "cutoff", "radii", "first_idx", "num_points")], 0.05, S, K, 0, 0)


def test_bin_counts_bit_exact():
    S = 96
    inp = make_splat_inputs(3, [2000, 1500, 10], S, seed=11, sigma_px=2.5)
    t = _t(inp)
    for b in (8, 16, 32):
        got = splat.bin_counts(t["points"], t["radii"], t["first_idx"], t["num_points"], S, b).cpu().numpy()
        assert np.array_equal(got, port.splat_bin_counts(inp["points"], inp["radii"], inp["first_idx"],
                                                         inp["num_points"], S, b))


@REF
@pytest.mark.parametrize("bin_size", [0, 16])
def test_forward_vs_reference_cuda_kernels(bin_size, raster_variant):
    S, K = 128, 8
    inp = make_splat_inputs(3, [6000, 4000, 5000], S, seed=21, sigma_px=1.5)
    t = _t(inp)
    C = ref_native.dss_C()
    M = int(max(10000, inp["num_points"].max()))
    ri, rz, rq, ro = C.splat_points(t["points"], t["ellipse"], t["cutoff"], t["radii"], t["first_idx"],
                                    t["num_points"], 0.05, S, K, bin_size, M)
    idx, zbuf, qv, occ = _fwd(t, S, K, bin_size=bin_size)
    assert torch.equal(idx, ri) and torch.equal(zbuf, rz) and torch.equal(occ, ro)
    assert torch.equal(qv, rq)
    if bin_size:
        bp = C._rasterize_coarse(t["points"], t["radii"], t["first_idx"], t["num_points"], S, bin_size, M)
        want = (bp >= 0).sum(-1).int()
        got = splat.bin_counts(t["points"], t["radii"], t["first_idx"], t["num_points"], S, bin_size)
        assert torch.equal(got, want)                 # per-tile point counts bit-exact


@REF
def test_forward_c4_scale_vs_reference_and_properties(raster_variant):
    """BASELINE config 4: 8 views x 300 000 splats at 512^2, K = 8, bin_size = 32."""
    S, K, V, Pv = 512, 8, 8, 300_000
    inp = make_splat_inputs(V, Pv, S, seed=0, sigma_px=1.5, aniso=False)
    t = _t(inp)
    idx, zbuf, qv, occ = _fwd(t, S, K, bin_size=32)
    C = ref_native.dss_C()
    ri, rz, rq, ro = C.splat_points(t["points"], t["ellipse"], t["cutoff"], t["radii"], t["first_idx"],
                                    t["num_points"], 0.05, S, K, 32, Pv)
    assert torch.equal(occ, ro)
    same = (idx == ri).all(-1)
    assert float(same.float().mean()) > 0.99999          # equal-z ties are order dependent in the reference
    assert torch.equal(zbuf, rz) and torch.equal(qv[same], rq[same])
    # size-independent properties
    valid = idx >= 0
    assert torch.equal(occ.bool(), valid[..., 0])
    zz = torch.where(valid, zbuf, torch.full_like(zbuf, float("inf")))
    assert (zz[..., 1:] >= zz[..., :-1]).all()
    assert ((zbuf - zbuf[..., :1] <= 0.05) | ~valid).all()
    assert (valid[..., 1:] <= valid[..., :-1]).all()      # -1 padding is a suffix
    view = torch.arange(V, device=DEV).view(V, 1, 1, 1).expand_as(idx)
    assert ((idx // Pv == view) | ~valid).all()           # a view only sees its own points
    assert ((qv <= 1.0) | ~valid).all() and (qv[valid] >= 0).all()


def _backward_inputs(S, V, pts, seed, K=6):
    inp = make_splat_inputs(V, pts, S, seed=seed, sigma_px=1.6)
    t = _t(inp)
    idx, zbuf, qv, occ = _fwd(t, S, K)
    g = torch.Generator().manual_seed(seed)
    occ_grad = torch.randn(V, S, S, generator=g) * (torch.rand(V, S, S, generator=g) < 0.3)
    zbuf_grad = torch.randn(V, S, S, K, generator=g) * (torch.rand(V, S, S, K, generator=g) < 0.7)
    return inp, t, idx, occ_grad.to(DEV), zbuf_grad.to(DEV)


@pytest.fixture(params=[True, False], ids=["hybrid", "window"])
def occ_sweep(request):
    old = splat.OCC_BACKWARD_HYBRID
    splat.OCC_BACKWARD_HYBRID = request.param
    yield request.param
    splat.OCC_BACKWARD_HYBRID = old


def test_backward_matches_oracle(occ_sweep):
    S, V = 64, 2
    inp, t, idx, occ_grad, zbuf_grad = _backward_inputs(S, V, [700, 500], seed=5)
    pts = t["points"].clone().requires_grad_(True)
    pcl = Pointclouds([pts[:700], pts[700:]])
    out = splat.rasterize_elliptical_points(pcl, t["ellipse"], t["cutoff"][:1], t["radii"], 0.05, S, 6,
                                            bin_size=None, radii_backward_scaler=4.0)
    assert torch.equal(out[0], idx)
    loss = (out[3] * occ_grad).sum() + (out[1] * zbuf_grad).sum() + out[2].sum()   # qvalue grad is ignored
    loss.backward()
    want, rs = port.splat_backward(inp["points"], inp["radii"], idx.cpu().numpy(), inp["first_idx"],
                                   inp["num_points"], occ_grad.cpu().numpy(), zbuf_grad.cpu().numpy(), 4.0)
    got = pts.grad.cpu().numpy()
    scale = np.abs(want).max(0)
    np.testing.assert_allclose(got / scale, want / scale, rtol=1e-4, atol=1e-5)
    assert np.abs(want[:, :2]).sum() > 0 and np.abs(want[:, 2]).sum() > 0
    vis = port.visibility(idx.cpu().numpy(), (idx[..., 0] >= 0).cpu().numpy(), len(inp["points"]))
    assert np.array_equal(splat.visibility_mask(idx, len(inp["points"])).cpu().numpy(), vis)
    assert (got[~vis, :2] == 0).all()


def test_backward_dense_gradient_hybrid_equals_window():
    """Every pixel carries a gradient: the hybrid sweep takes its dense-tile branch."""
    S, V = 96, 2
    inp = make_splat_inputs(V, [1500, 1100], S, seed=17, sigma_px=1.6)
    t = _t(inp)
    occ_grad = torch.randn(V, S, S, device=DEV)
    res = []
    for hybrid in (True, False):
        splat.OCC_BACKWARD_HYBRID = hybrid
        pts = t["points"].clone().requires_grad_(True)
        out = splat.EllipticalRasterizer.apply(pts, t["ellipse"], t["cutoff"], t["radii"], t["first_idx"],
                                               t["num_points"], 0.05, S, 6, 16, 0, 5.0)
        (out[3] * occ_grad).sum().backward()
        res.append(pts.grad.clone())
    splat.OCC_BACKWARD_HYBRID = True
    scale = res[1].abs().amax(0).clamp_min(1e-20)
    np.testing.assert_allclose((res[0] / scale).cpu().numpy(), (res[1] / scale).cpu().numpy(), rtol=1e-4, atol=2e-5)
    assert res[1][:, :2].abs().sum() > 0


@REF
@pytest.mark.parametrize("V", [1, 2])
def test_backward_vs_reference_cuda(V, occ_sweep):
    """Fast-path occupancy + z-buffer backward vs the reference's own kernels and host sequence.
    For views n >= 1 the reference drops the points of its last 2-D grid cell (packed-vs-local
    offset bug, rasterize_points_backward.cu:124-126): those rows are excluded, and counted."""
    S = 128
    inp, t, idx, occ_grad, zbuf_grad = _backward_inputs(S, V, [5000, 4000][:V], seed=9)
    ref_g, ref_rs, off, params, gidx, vis = ref_native.splat_backward_fast_cuda(
        t["points"], t["radii"], idx, t["first_idx"], t["num_points"], occ_grad, zbuf_grad, 10.0)
    pts = t["points"].clone().requires_grad_(True)
    out = splat.EllipticalRasterizer.apply(pts, t["ellipse"], t["cutoff"], t["radii"], t["first_idx"],
                                           t["num_points"], 0.05, S, 6, 16, 10000, 10.0)
    ((out[3] * occ_grad).sum() + (out[1] * zbuf_grad).sum()).backward()
    got = pts.grad
    assert torch.equal(splat.visibility_mask(idx, pts.shape[0]), vis)
    rs = splat.per_view_median_radius(t["radii"], vis, t["first_idx"], t["num_points"]) * 10.0
    assert torch.equal(rs, ref_rs)
    keep = torch.ones(pts.shape[0], dtype=torch.bool, device=DEV)
    if V > 1:   # rows living in the last grid cell of views >= 1
        num_v = torch.stack([x.sum() for x in torch.split(vis, inp["num_points"].tolist())])
        first_v = torch.cumsum(num_v, 0) - num_v
        vis_rows = vis.nonzero().squeeze(1)
        for n in range(1, V):
            total = int(params[n, 5].item())
            start = int(off[n, total - 1].item())
            end = int(first_v[n] + num_v[n])
            keep[vis_rows[gidx[start:end]]] = False
        assert int((~keep).sum()) < 0.02 * pts.shape[0]
    scale = ref_g.abs().amax(0).clamp_min(1e-20)
    np.testing.assert_allclose((got / scale)[keep].cpu().numpy(), (ref_g / scale)[keep].cpu().numpy(),
                               rtol=1e-4, atol=2e-5)
    assert torch.equal(got[:, 2], ref_g[:, 2]) or torch.allclose(got[:, 2], ref_g[:, 2], rtol=1e-5, atol=1e-6)


@REF
def test_slow_path_occ_backward_vs_reference(occ_sweep):
    S = 64
    inp, t, idx, occ_grad, _ = _backward_inputs(S, 2, [800, 600], seed=13)
    C = ref_native.dss_C()
    want = C._splat_points_occ_backward(t["points"], t["radii"], occ_grad, t["first_idx"], t["num_points"], 3.0, 0.05)
    got = splat._C._splat_points_occ_backward(t["points"], t["radii"], occ_grad, t["first_idx"], t["num_points"],
                                              3.0, 0.05)
    scale = want.abs().amax(0)
    np.testing.assert_allclose((got / scale).cpu().numpy(), (want / scale).cpu().numpy(), rtol=1e-4, atol=2e-5)
    gz = torch.zeros(t["points"].shape[0], 1, device=DEV)
    gz_ref = torch.zeros_like(gz)
    zg = torch.randn(idx.shape, device=DEV)
    splat._C._backward_zbuf(idx, zg, gz)
    C._backward_zbuf(idx, zg, gz_ref)
    assert torch.allclose(gz, gz_ref, rtol=1e-5, atol=1e-5)


def test_blend_and_feature_gradient():
    S, K = 48, 5
    inp = make_splat_inputs(2, [900, 700], S, seed=8)
    t = _t(inp)
    idx, zbuf, qv, occ = _fwd(t, S, K)
    P = inp["points"].shape[0]
    g = torch.Generator().manual_seed(0)
    scaler = (torch.rand(P, generator=g) + 0.5).to(DEV)
    rgb = torch.rand(P, 3, generator=g).to(DEV).requires_grad_(True)
    img = splat.blend_rgba(idx, qv, occ, scaler, rgb)
    want = port.blend(idx.cpu().numpy(), qv.cpu().numpy(), occ.cpu().numpy(), scaler.cpu().numpy(),
                      rgb.detach().cpu().numpy())
    np.testing.assert_allclose(img.detach().cpu().numpy(), want, rtol=1e-4, atol=1e-6)
    w = torch.rand_like(img)
    (img * w).sum().backward()
    # torch autograd of the same formula as the fp32 reference of the backward kernel
    rgb2 = rgb.detach().clone().requires_grad_(True)
    m = (idx >= 0)
    j = idx.clamp_min(0).long()
    wt = torch.exp(-0.5 * qv) * scaler[j] * m
    col = (wt[..., None] * rgb2[j]).sum(-2) / wt.sum(-1, keepdim=True).clamp_min(1e-4)
    (col * w[..., :3]).sum().backward()
    np.testing.assert_allclose(rgb.grad.cpu().numpy(), rgb2.grad.cpu().numpy(), rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("K,C", [(8, 3), (5, 4), (1, 1), (16, 2)])
def test_fused_blend_and_visibility_equal_the_separate_kernels(K, C):
    """SplatRender (RGBA blend + per-point visibility in the raster kernel's epilogue) against the two-step path it
    replaces -- EllipticalRasterizer + blend_rgba, visibility_mask: same bits forward, same gradients to the screen
    points (occupancy incl. the image's alpha channel, depth) and to the features."""
    S = 96
    inp = make_splat_inputs(3, [2500, 0, 1800], S, seed=31 + K, sigma_px=1.7)
    t = _t(inp)
    P = inp["points"].shape[0]
    g = torch.Generator().manual_seed(K)
    scaler = (torch.rand(P, generator=g) + 0.5).to(DEV)
    feat0 = torch.rand(P, C, generator=g).to(DEV)
    w_img = torch.randn(3, S, S, C + 1, generator=g).to(DEV)
    w_occ = (torch.randn(3, S, S, generator=g) * (torch.rand(3, S, S, generator=g) < 0.2)).to(DEV)
    w_z = torch.randn(3, S, S, K, generator=g).to(DEV)
    res = {}
    for fused in (False, True):
        pts = t["points"].clone().requires_grad_(True)
        feat = feat0.clone().requires_grad_(True)
        if fused:
            idx, zbuf, qv, occ, img = splat.SplatRender.apply(pts, t["ellipse"], t["cutoff"], t["radii"],
                                                              t["first_idx"], t["num_points"], 0.05, S, K, 16, 10.0,
                                                              scaler, feat, splat.NORM_WEIGHT_EPS)
        else:
            idx, zbuf, qv, occ = splat.EllipticalRasterizer.apply(pts, t["ellipse"], t["cutoff"], t["radii"],
                                                                  t["first_idx"], t["num_points"], 0.05, S, K, 16,
                                                                  10000, 10.0)
            img = splat.blend_rgba(idx, qv, occ, scaler, feat)
        ((img * w_img).sum() + (occ * w_occ).sum() + (zbuf * w_z).sum()).backward()
        res[fused] = (idx, zbuf, qv, occ, img.detach(), pts.grad, feat.grad)
    for a, b in zip(res[False][:5], res[True][:5]):
        assert torch.equal(a, b)
    # xy (occupancy) gradients are plain per-point sums in a fixed order: identical; z and feature gradients are
    # fp32 atomics in both paths (rasterize_points.cu:835-843 does the same): equal up to summation order
    assert torch.equal(res[True][5][:, :2], res[False][5][:, :2])
    np.testing.assert_allclose(res[True][5][:, 2].cpu().numpy(), res[False][5][:, 2].cpu().numpy(), rtol=5e-4, atol=1e-5)
    np.testing.assert_allclose(res[True][6].cpu().numpy(), res[False][6].cpu().numpy(), rtol=5e-4, atol=1e-5)
    # the epilogue's visibility == the idx sweep
    out = splat._splat(t["points"], t["ellipse"], t["cutoff"], t["radii"], t["first_idx"], t["num_points"], 0.05, S, K,
                       False, want_visible=True)
    assert torch.equal(out[6].bool(), splat.visibility_mask(out[0], P))


@pytest.mark.parametrize("scaler,density", [(24.0, 1.0), (9.0, 0.15), (60.0, 0.02)])
def test_backward_tiled_sweep_many_passes_equals_window(scaler, density):
    """Search radii of several tiles: the tile-owner kernel stages its neighbourhood in more than one pass (a pass
    holds 2 304 records or 64 tiles) and accumulates across passes; it must still equal the plain window sweep."""
    S, V = 160, 2
    inp = make_splat_inputs(V, [2500, 1800], S, seed=29, sigma_px=1.6)
    t = _t(inp)
    g = torch.Generator().manual_seed(int(scaler))
    occ_grad = (torch.randn(V, S, S, generator=g) * (torch.rand(V, S, S, generator=g) < density)).to(DEV)
    res = []
    for hybrid in (True, False):
        splat.OCC_BACKWARD_HYBRID = hybrid
        pts = t["points"].clone().requires_grad_(True)
        out = splat.EllipticalRasterizer.apply(pts, t["ellipse"], t["cutoff"], t["radii"], t["first_idx"],
                                               t["num_points"], 0.05, S, 6, 16, 0, scaler)
        (out[3] * occ_grad).sum().backward()
        res.append(pts.grad.clone())
    splat.OCC_BACKWARD_HYBRID = True
    scale = res[1].abs().amax(0).clamp_min(1e-20)
    np.testing.assert_allclose((res[0] / scale).cpu().numpy(), (res[1] / scale).cpu().numpy(), rtol=2e-4, atol=2e-5)
    assert res[1][:, :2].abs().sum() > 0


def test_blend_request_with_more_than_16_points_per_pixel_uses_the_separate_kernels():
    """The fused epilogue exists for K <= 16; above that rasterize_elliptical_points(blend=...) composes the same
    image with blend_rgba."""
    S, K = 48, 20
    inp = make_splat_inputs(1, 1200, S, seed=4, sigma_px=2.2, behind_frac=0.0)
    t = _t(inp)
    P = inp["points"].shape[0]
    rgb = torch.rand(P, 3, device=DEV)
    pcl = Pointclouds([t["points"]])
    out = splat.rasterize_elliptical_points(pcl, t["ellipse"], t["cutoff"][:1], t["radii"], 0.05, S, K, bin_size=None,
                                            blend=(None, rgb, splat.NORM_WEIGHT_EPS))
    assert len(out) == 5 and out[4].shape == (1, S, S, 4)
    assert torch.equal(out[4], splat.blend_rgba(out[0], out[2], out[3], None, rgb))
