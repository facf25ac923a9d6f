"""GPU parity: FRNN grid build / query / gather / backward through the C ABI vs the oracle
(oracle/port.py) and, when present, the reference's own CUDA kernels (oracle/_ref)."""
import numpy as np
import pytest
import torch

from isopoints_b200 import _ext, frnn
from oracle import port, ref_native

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _scan(x):
    lib = _ext.lib()
    rows, n = x.shape
    out = torch.empty_like(x)
    ws = _ext.workspace(lib.isob200_exclusive_scan_ws_bytes(n, rows), x.device)
    _ext.check(lib.isob200_exclusive_scan_i32(_ext.ptr(x), _ext.ptr(out), n, rows, n, n, _ext.ptr(ws),
                                              ws.numel(), _ext.stream(x.device)))
    return out


@pytest.mark.parametrize("n,rows", [(1, 1), (7, 3), (2048, 1), (2049, 2), (68921, 1), (1 << 20, 2), (3_000_001, 1)])
def test_exclusive_scan(n, rows):
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 50, (rows, n), generator=g, dtype=torch.int32)
    want = torch.cumsum(x.long(), 1) - x.long()
    got = _scan(x.to(DEV)).cpu().long()
    assert torch.equal(got, want)


def test_prefix_sum_cuda_api_in_place():
    x = torch.randint(0, 9, (5000,), dtype=torch.int32, device=DEV)
    off = torch.zeros_like(x)
    frnn.prefix_sum_cuda(x, 5000, off)
    assert torch.equal(off.cpu().long(), (torch.cumsum(x.long(), 0) - x.long()).cpu())


@pytest.fixture(params=[1, 2, 3], ids=["exhaustive", "pruned", "collect"])
def query_mode(request):
    old = frnn.QUERY_MODE
    frnn.QUERY_MODE = request.param
    yield request.param
    frnn.QUERY_MODE = old


@pytest.mark.parametrize("D", [3, 2])
@pytest.mark.parametrize("K", [1, 5, 8, 9, 16, 17, 32])
def test_frnn_small_bit_exact_vs_oracle(D, K, query_mode):
    rng = np.random.RandomState(10 * D + K)
    N, P = 2, 700
    pts = rng.rand(N, P, D).astype(np.float32)
    lens = np.array([700, 523])
    rs = np.array([0.12, 0.2], np.float32)
    want_i, want_d = port.frnn_bruteforce(pts, pts, lens, lens, K=K, r=rs, inclusive=True)
    t = torch.as_tensor(pts, device=DEV)
    l = torch.as_tensor(lens, device=DEV)
    d, i, nn, grid = frnn.frnn_grid_points(t, t, l, l, K=K, r=torch.as_tensor(rs), return_nn=True)
    assert i.dtype == torch.int64 and d.dtype == torch.float32
    assert np.array_equal(i.cpu().numpy(), want_i)
    assert np.array_equal(d.cpu().numpy(), want_d)          # same fp32 expression -> bit equal
    np.testing.assert_array_equal(nn.cpu().numpy(), port.frnn_gather(pts, want_i))
    # grid structure == restated reference structure (params, offsets, sorted points)
    params, G = port.frnn_grid_params(pts, lens, rs)
    sp, off, sidx = port.frnn_build_grid(pts, lens, params, G)
    assert np.array_equal(grid.grid_params.cpu().numpy(), params)
    assert np.array_equal(grid.pc2_grid_off.cpu().numpy(), off)
    assert np.array_equal(grid.sorted_points2_idxs.cpu().numpy(), sidx)
    assert np.array_equal(grid.sorted_points2.cpu().numpy(), sp)


def test_frnn_multi_tile_multi_pass_build(query_mode):
    """Several radix tiles x 3 digit passes in the deterministic grid build (N*G > 2^16)."""
    rng = np.random.RandomState(77)
    pts = rng.rand(2, 5000, 3).astype(np.float32)
    lens = np.array([5000, 4100])
    rs = np.array([0.03, 0.045], np.float32)
    t = torch.as_tensor(pts, device=DEV)
    l = torch.as_tensor(lens, device=DEV)
    d, i, _, grid = frnn.frnn_grid_points(t, t, l, l, K=8, r=torch.as_tensor(rs))
    params, G = port.frnn_grid_params(pts, lens, rs)
    assert 2 * G > (1 << 16)
    sp, off, sidx = port.frnn_build_grid(pts, lens, params, G)
    assert np.array_equal(grid.pc2_grid_off.cpu().numpy(), off)
    assert np.array_equal(grid.sorted_points2_idxs.cpu().numpy(), sidx)
    assert np.array_equal(grid.sorted_points2.cpu().numpy(), sp)
    want_i, want_d = port.frnn_bruteforce(pts, pts, lens, lens, K=8, r=rs)
    assert np.array_equal(i.cpu().numpy(), want_i) and np.array_equal(d.cpu().numpy(), want_d)


def test_frnn_two_clouds_and_grid_reuse(query_mode):
    rng = np.random.RandomState(5)
    p1 = rng.rand(1, 400, 3).astype(np.float32) * 1.2 - 0.1      # queries partly outside the grid
    p2 = rng.rand(1, 900, 3).astype(np.float32)
    want_i, want_d = port.frnn_bruteforce(p1, p2, K=6, r=0.15)
    a, b = torch.as_tensor(p1, device=DEV), torch.as_tensor(p2, device=DEV)
    d, i, _, grid = frnn.frnn_grid_points(a, b, K=6, r=0.15)
    assert np.array_equal(i.cpu().numpy(), want_i) and np.array_equal(d.cpu().numpy(), want_d)
    d2, i2, _, _ = frnn.frnn_grid_points(a, b, K=6, r=0.15, grid=grid)      # cached grid
    assert torch.equal(i, i2) and torch.equal(d, d2)


def test_frnn_r_forms_and_errors():
    p = torch.rand(2, 300, 3, device=DEV)
    base = frnn.frnn_grid_points(p, p, K=4, r=0.2)[1]
    for r in (torch.tensor([0.2]), torch.tensor([0.2], device=DEV), torch.tensor([0.2, 0.2]),
              torch.tensor([0.2, 0.2], device=DEV)):       # tests/frnn_r.py:32-46
        assert torch.equal(frnn.frnn_grid_points(p, p, K=4, r=r)[1], base)
    with pytest.raises(ValueError):
        frnn.frnn_grid_points(p[..., :1].contiguous(), p[..., :1].contiguous(), K=4, r=0.2)
    with pytest.raises(RuntimeError):
        frnn.frnn_grid_points(p, p, K=33, r=0.2)
    with pytest.raises(TypeError):
        frnn.frnn_grid_points(p.cpu(), p.cpu(), K=4, r=0.2)


def test_frnn_empty_and_degenerate():
    p = torch.rand(1, 64, 3, device=DEV)
    l0 = torch.tensor([0], device=DEV)
    d, i, _, _ = frnn.frnn_grid_points(p, p, l0, torch.tensor([64], device=DEV), K=3, r=0.5)
    assert (i == -1).all() and (d == -1).all()
    same = torch.zeros(1, 40, 3, device=DEV)                   # all points identical: zero extent
    d, i, _, _ = frnn.frnn_grid_points(same, same, K=4, r=0.1)
    assert (d == 0).all() and (i[0, :, 0] == 0).all() and (i[0, :, 3] == 3).all()   # ties -> index order


def test_frnn_backward_matches_oracle():
    rng = np.random.RandomState(9)
    p1 = rng.rand(2, 200, 3).astype(np.float32)
    p2 = rng.rand(2, 300, 3).astype(np.float32)
    a = torch.as_tensor(p1, device=DEV).requires_grad_(True)
    b = torch.as_tensor(p2, device=DEV).requires_grad_(True)
    d, i, _, _ = frnn.frnn_grid_points(a, b, K=5, r=0.25)
    g = torch.as_tensor(rng.randn(2, 200, 5).astype(np.float32), device=DEV)
    (d * g).sum().backward()
    ga, gb = port.frnn_backward(p1, p2, i.cpu().numpy(), g.cpu().numpy())
    np.testing.assert_allclose(a.grad.cpu().numpy(), ga, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(b.grad.cpu().numpy(), gb, rtol=1e-4, atol=1e-5)


def test_frnn_gather_backward():
    x = torch.rand(2, 50, 3, device=DEV, requires_grad=True)
    idx = torch.randint(-1, 50, (2, 70, 4), device=DEV)
    out = frnn.frnn_gather(x, idx)
    want = port.frnn_gather(x.detach().cpu().numpy(), idx.cpu().numpy())
    np.testing.assert_array_equal(out.detach().cpu().numpy(), want)
    w = torch.rand_like(out)
    (out * w).sum().backward()
    ref = torch.zeros(2, 50, 3, dtype=torch.float64)
    for n in range(2):
        m = idx[n].cpu() >= 0
        ref[n].index_add_(0, idx[n].cpu()[m], w[n].cpu().double()[m])
    np.testing.assert_allclose(x.grad.cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)


def test_compat_primitives_match_reference_layout():
    """frnn._C.insert_points_cuda / counting_sort_cuda (used directly by DSS's backward,
    rasterizer.py:909-929): same cells and counts, a valid in-cell ranking."""
    rng = np.random.RandomState(3)
    pts = torch.as_tensor(rng.rand(2, 500, 2).astype(np.float32), device=DEV)
    lens = torch.tensor([500, 320], device=DEV)
    rs = np.array([0.08, 0.08], np.float32)
    params, G = port.frnn_grid_params(pts.cpu().numpy(), lens.cpu().numpy(), rs)
    prm = torch.as_tensor(params, device=DEV)
    cnt = torch.zeros((2, G), dtype=torch.int32, device=DEV)
    cell = torch.full((2, 500), -1, dtype=torch.int32, device=DEV)
    gidx = torch.full((2, 500), -1, dtype=torch.int32, device=DEV)
    frnn._C.insert_points_cuda(pts, lens, prm, cnt, cell, gidx, G)
    off = torch.zeros_like(cnt)
    for n in range(2):
        frnn.prefix_sum_cuda(cnt[n], G, off[n])
    spts = torch.zeros_like(pts)
    sidx = torch.full((2, 500), -1, dtype=torch.int32, device=DEV)
    frnn._C.counting_sort_cuda(pts, lens, cell, gidx, off, spts, sidx)
    for n in range(2):
        L = int(lens[n])
        want_cell = port.frnn_cell_ids(pts[n, :L].cpu().numpy(), params[n])
        assert np.array_equal(cell[n, :L].cpu().numpy(), want_cell)
        assert (cell[n, L:] == -1).all()
        assert np.array_equal(cnt[n].cpu().numpy(), np.bincount(want_cell, minlength=G))
        s = sidx[n, :L].cpu().numpy()
        assert np.array_equal(np.sort(s), np.arange(L))
        assert np.array_equal(spts[n, :L].cpu().numpy(), pts[n].cpu().numpy()[s])
        assert (np.diff(want_cell[s]) >= 0).all()            # cell-sorted


@pytest.mark.skipif(not ref_native.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", ["box", "sphere"])
def test_frnn_500k_bit_exact_vs_reference_cuda(shape, query_mode):
    """BASELINE config 3: 500 000 points, r = 0.05, K = 16, index bit-exact vs the reference."""
    g = torch.Generator().manual_seed(0)
    if shape == "box":
        p = torch.rand(1, 500_000, 3, generator=g)
    else:
        p = torch.nn.functional.normalize(torch.randn(1, 500_000, 3, generator=g), dim=-1)
    p = p.to(DEV)
    lens = torch.tensor([500_000], device=DEV)
    r = torch.tensor([0.05], device=DEV)
    d, i, _, grid = frnn.frnn_grid_points(p, p, lens, lens, K=16, r=0.05)
    ri, rd, rsp2, roff, rsidx, rparams = ref_native.frnn_grid_points_cuda(p, p, lens, lens, 16, r)
    assert torch.equal(grid.grid_params, rparams)
    assert torch.equal(grid.pc2_grid_off, roff)
    # distances must match bit for bit everywhere; indices too, except inside a run of EXACTLY equal
    # distances (a handful of genuine fp32 ties at this size), where the reference's order is its
    # atomic insertion order (mink.cuh:64) and ours is ascending index -- there the sets must agree
    full = d[0, :, -1] >= 0
    assert torch.equal(d, rd)
    tie = torch.zeros_like(d[0], dtype=torch.bool)
    eq = (d[0, :, 1:] == d[0, :, :-1]) & (d[0, :, 1:] >= 0)
    tie[:, 1:] |= eq
    tie[:, :-1] |= eq
    assert int(eq.sum()) < 50
    assert torch.equal(i[0][~tie], ri[0][~tie])
    rows = tie.any(-1)
    assert torch.equal(i[0][rows].sort(-1).values, ri[0][rows].sort(-1).values)
    # size-independent properties
    assert (i[0, :, 0] == torch.arange(500_000, device=DEV)).all()          # self is nearest
    dd = torch.where(d < 0, torch.full_like(d, float("inf")), d)
    assert (dd[..., 1:] >= dd[..., :-1]).all() and (d <= 0.05 * 0.05).all()
    if shape == "box":
        assert 0.9 < float(full.float().mean()) <= 1.0


@pytest.mark.parametrize("D", [3, 2])
def test_frnn_collect_equals_pruned_on_clustered_cloud(D):
    """The thread-per-query kernel fixes a trial radius from the MEAN density of the candidate block; on a cloud
    whose density varies by orders of magnitude the trial ball overflows (clusters) or comes up short (voids) and
    the warp-cooperative exact search takes over -- results must not depend on which path answered."""
    g = torch.Generator().manual_seed(5)
    P = 60_000
    centers = torch.rand(40, D, generator=g)
    clustered = centers[torch.randint(0, 40, (P // 2,), generator=g)] + 0.004 * torch.randn(P // 2, D, generator=g)
    p = torch.cat([clustered, torch.rand(P - P // 2, D, generator=g)])[None].to(DEV)
    lens = torch.tensor([P - 17], device=DEV)
    old = frnn.QUERY_MODE
    try:
        for K in (1, 4, 9, 16, 20, 32):
            out = {}
            for mode in (2, 3):
                frnn.QUERY_MODE = mode
                d, i, _, _ = frnn.frnn_grid_points(p, p, lens, lens, K=K, r=0.08)
                out[mode] = (d, i)
            assert torch.equal(out[2][0], out[3][0]) and torch.equal(out[2][1], out[3][1]), K
    finally:
        frnn.QUERY_MODE = old
