"""GPU parity: SphereTracing.project_points (isopoints_b200/levelset_sampling.py, csrc/project.cu
``trace_step_kernel``) vs golden vectors of the reference's SphereTracing and the oracle.

Ray marching iterates x <- x + sdf(x) d: an SDF difference of 1e-7 between two float32 implementations
(GPU kernels vs torch-CPU) moves a ray by as much and can flip a ray sitting on a threshold, so masks are
compared as an agreement rate and positions on the rays both sides call converged (tolerance 1e-4, the
north_star bar for fp32 positions)."""
import numpy as np
import pytest
import torch

from isopoints_b200 import _ext, siren
from isopoints_b200.levelset_sampling import SphereTracing
from oracle import port
from tests.helpers import Siren, SphereSDF, TinySiren, make_rays

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _compare(res, want_pts, want_eval, want_mask, min_agree=0.995, atol=1e-4):
    pts = res["levelset_points"].reshape(-1, 3).cpu().numpy()
    ev = res["network_eval_on_levelset_points"].reshape(-1).cpu().numpy()
    mask = res["mask"].reshape(-1).cpu().numpy()
    agree = mask == want_mask
    assert agree.mean() >= min_agree, agree.mean()
    both = mask & want_mask
    assert both.mean() > 0.2
    np.testing.assert_allclose(pts[both], want_pts[both], rtol=0, atol=atol)
    assert np.abs(ev[both]).max() <= 5e-5 + 1e-7
    # rays neither side calls converged: same fate (left the sphere / ran out of steps) for nearly all
    neither = ~mask & ~want_mask
    if neither.any():
        close = np.abs(pts[neither] - want_pts[neither]).max(axis=1) < 1e-3
        assert close.mean() > 0.98, close.mean()
    return agree.mean()


@pytest.mark.parametrize("name,net", [("siren", TinySiren(seed=3)), ("sphere", SphereSDF(radius=0.5))])
def test_matches_reference_golden(golden, name, net):
    g = golden("sphere_trace")
    ray0 = torch.as_tensor(g[name + "_ray0"], device=DEV).view(2, -1, 3)
    dirs = torch.as_tensor(g[name + "_dirs"], device=DEV).view(2, -1, 3)
    tracer = SphereTracing(proj_max_iters=int(g["proj_max_iters"]), proj_tolerance=5e-5)
    res = tracer.project_points(ray0, dirs, net.to(DEV))
    assert tuple(res["levelset_points"].shape) == tuple(ray0.shape) and tuple(res["mask"].shape) == tuple(ray0.shape[:2])
    assert res["levelset_points_Dx"] is res["levelset_points"]        # the reference's return dict (:806)
    _compare(res, g[name + "_points"], g[name + "_eval"], g[name + "_mask"])
    assert torch.equal(ray0, torch.as_tensor(g[name + "_ray0"], device=DEV).view(2, -1, 3))   # inputs untouched


def _zero_mean_siren(seed=4, layers=2):
    m = Siren(256, layers, 30.0, seed=seed)
    with torch.no_grad():
        x = (torch.rand(4000, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 2
        m.net[-1].bias -= m(x).sdf.mean()          # zero crossings everywhere inside the unit sphere
    return m


@pytest.mark.parametrize("n", [1, 1000, 40000])
def test_fused_siren_loop_vs_oracle_and_vs_opaque_path(n):
    """The reference's decoder takes the device-count loop (fused tcgen05 SDF kernel, no read-back)."""
    model = _zero_mean_siren()
    ray0, dirs = make_rays(n, seed=n, target_radius=0.7)
    tracer = SphereTracing(proj_max_iters=40, proj_tolerance=5e-5)
    calls0 = siren.STATS["calls"]
    res = tracer.project_points(ray0.to(DEV), dirs.to(DEV), model.to(DEV))
    assert siren.STATS["calls"] - calls0 == 41          # every iteration through the fused kernel
    assert tracer.last_gradient is None                 # forward half of the network only
    old = siren.ENABLED
    siren.ENABLED = False
    try:
        opaque = tracer.project_points(ray0.to(DEV), dirs.to(DEV), model)
    finally:
        siren.ENABLED = old
    assert siren.STATS["calls"] - calls0 == 41
    if n > 1:
        _compare(res, opaque["levelset_points"].cpu().numpy(), opaque["network_eval_on_levelset_points"].cpu().numpy(),
                 opaque["mask"].cpu().numpy(), min_agree=0.99)
        assert tuple(tracer.last_gradient.shape) == (n, 3)
    if n <= 1000:
        pts, sdf, grad, mask = port.sphere_trace(model.cpu(), ray0, dirs, proj_max_iters=40, proj_tolerance=5e-5)
        if n > 1:
            _compare(res, pts.numpy(), sdf.numpy(), mask.numpy(), min_agree=0.99)
        else:
            assert bool(res["mask"].cpu()[0]) == bool(mask[0])


@pytest.mark.parametrize("layers", [1, 3, 7])
def test_fused_forward_only_step_equals_sdf_kernel_plus_trace_step(layers):
    """isob200_siren_trace_step == isob200_siren_sdf_grad (value) + isob200_trace_step, bit for bit, for odd and
    even GEMM counts per tile and a ragged last tile."""
    model = _zero_mean_siren(seed=layers, layers=layers).to(DEV)
    M = 128 * 150 + 37
    ray0, dirs = make_rays(M, seed=layers, target_radius=0.7)
    lib = _ext.lib()
    dev = torch.device(DEV)
    spec = siren.match(model, {})
    blob, scratch, L = siren.packed(model, spec)
    dirs_d = dirs.to(DEV)
    args = (0.1 * 5e-5, 1.0, 0.1, 1.1)
    out = {}
    for fused in (False, True):
        pts = ray0.to(DEV).clone()
        ev = torch.zeros(M, device=DEV)
        act = [torch.empty(M, dtype=torch.int32, device=DEV) for _ in range(2)]
        nxt = [torch.empty(M, 3, device=DEV) for _ in range(2)]
        cnt = torch.zeros(4, dtype=torch.int32, device=DEV)
        ws = _ext.workspace(lib.isob200_project_step_ws_bytes(M), dev)
        for it in range(3):
            cur = pts if it == 0 else nxt[it & 1]
            n_dev = None if it == 0 else cnt[it:]
            a_in = None if it == 0 else act[it & 1]
            if fused:
                _ext.check(lib.isob200_siren_trace_step(
                    _ext.ptr(cur), M, _ext.ptr(n_dev), _ext.ptr(blob), L, _ext.ptr(scratch), scratch.numel(),
                    _ext.ptr(pts), _ext.ptr(dirs_d), _ext.ptr(ev), _ext.ptr(a_in), *args, 1, _ext.ptr(act[(it + 1) & 1]),
                    _ext.ptr(nxt[(it + 1) & 1]), _ext.ptr(cnt[it + 1:]), _ext.stream(DEV)))
            else:
                sdf, grad = siren.sdf_and_grad(model, cur, n_dev=n_dev, spec=spec)
                _ext.check(lib.isob200_trace_step(
                    _ext.ptr(pts), _ext.ptr(dirs_d), _ext.ptr(ev), None, _ext.ptr(a_in), M, _ext.ptr(n_dev), _ext.ptr(sdf),
                    None, *args, 1, _ext.ptr(act[(it + 1) & 1]), _ext.ptr(nxt[(it + 1) & 1]), _ext.ptr(cnt[it + 1:]),
                    _ext.ptr(ws), ws.numel(), _ext.stream(DEV)))
        k = int(cnt[3])
        out[fused] = (pts.cpu(), ev.cpu(), cnt.cpu(), set(act[1][:k].cpu().tolist()))
    assert torch.equal(out[True][2], out[False][2]) and 0 < int(out[True][2][3]) < M
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    assert out[True][3] == out[False][3]              # same active set (tiles append in completion order)


def test_trace_step_kernel_bit_exact_vs_oracle_step():
    """One call of isob200_trace_step on given sdf values = one iteration of the oracle's update rule."""
    M = 5000
    g = torch.Generator().manual_seed(3)
    pts, dirs = make_rays(M, seed=5, start_radius=1.05)
    sdf = (torch.rand(M, generator=g) - 0.3) * 0.3
    sdf[::7] = 1e-6                                      # below 0.1 tol: retire without moving
    grad = torch.randn(M, 3, generator=g)
    alpha, tol, bound = 0.9, 5e-5, 1.1
    move = alpha * sdf[:, None] * dirs
    nrm = move.norm(dim=-1, keepdim=True)
    new = pts + move / nrm.clamp_min(1e-15) * nrm.clamp_max(0.1)
    still = sdf.abs() > 0.1 * tol
    ok = new.norm(dim=-1) < bound
    want = torch.where((still & ok)[:, None], new, pts)
    lib = _ext.lib()
    d = lambda x: x.to(DEV).contiguous()  # noqa: E731
    p_d, dirs_d, sdf_d, grad_d = d(pts.clone()), d(dirs), d(sdf), d(grad)
    ev = torch.zeros(M, device=DEV)
    go = torch.zeros(M, 3, device=DEV)
    act = torch.empty(M, dtype=torch.int32, device=DEV)
    nxt = torch.empty(M, 3, device=DEV)
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    ws = _ext.workspace(lib.isob200_project_step_ws_bytes(M), torch.device(DEV))
    _ext.check(lib.isob200_trace_step(_ext.ptr(p_d), _ext.ptr(dirs_d), _ext.ptr(ev), _ext.ptr(go), None, M, None,
                                      _ext.ptr(sdf_d), _ext.ptr(grad_d), 0.1 * tol, alpha, 0.1, bound, 1, _ext.ptr(act),
                                      _ext.ptr(nxt), _ext.ptr(cnt), _ext.ptr(ws), ws.numel(), _ext.stream(DEV)))
    keep = still & ok
    # positions: identical arithmetic up to the rounding of the normalise / clamp (1 ulp)
    np.testing.assert_allclose(p_d.cpu().numpy(), want.numpy(), rtol=0, atol=2e-7)
    assert torch.equal(p_d.cpu()[~keep], pts[~keep])                      # retired rays do not move
    k = int(cnt.item())
    got_keep = torch.zeros(M, dtype=torch.bool)
    got_keep[act[:k].cpu().long()] = True
    border = (new.norm(dim=-1) - bound).abs() < 1e-6
    assert torch.equal(got_keep[~border], keep[~border])
    assert torch.equal(act[:k].cpu().long(), torch.nonzero(got_keep).reshape(-1))   # ascending, order-preserving
    assert torch.equal(nxt[:k].cpu(), p_d.cpu()[got_keep])
    assert torch.equal(ev.cpu(), sdf) and torch.equal(go.cpu(), grad)


def test_edge_cases():
    net = SphereSDF(radius=0.5).to(DEV)
    tracer = SphereTracing(proj_max_iters=5)
    res = tracer.project_points(torch.zeros(0, 3, device=DEV), torch.zeros(0, 3, device=DEV), net)
    assert res["levelset_points"].shape == (0, 3) and res["mask"].shape == (0,)
    with pytest.raises(TypeError):
        tracer.project_points(torch.zeros(4, 3), torch.zeros(4, 3), net)
    # a ray whose first step ends outside radius + padding retires where it started (:774-777)
    ray0 = torch.tensor([[0.0, 0.0, 1.05], [0.0, 0.0, -3.0]], device=DEV)
    dirs = torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.0, 1.0]], device=DEV)
    res = SphereTracing(proj_max_iters=30).project_points(ray0, dirs, net)
    assert res["mask"].tolist() == [True, False]
    np.testing.assert_allclose(res["levelset_points"][0].cpu().numpy(), [0, 0, 0.5], atol=1e-5)
    np.testing.assert_allclose(res["levelset_points"][1].cpu().numpy(), [0, 0, -3.0], atol=0)   # left at once
    # a per-cloud latent code is handed to the model per ray (:44-52)
    seen = {}

    class WithLatent(torch.nn.Module):
        def forward(self, x, c=None, **kw):
            seen.setdefault("shapes", []).append((tuple(x.shape), None if c is None else tuple(c.shape)))
            import types
            return types.SimpleNamespace(sdf=x.norm(dim=-1, keepdim=True) - 0.5 + 0 * c[:, :1])
    o, dd = make_rays(64, seed=2)
    res = SphereTracing(proj_max_iters=12).project_points(o.view(2, 32, 3).to(DEV), dd.view(2, 32, 3).to(DEV),
                                                          WithLatent().to(DEV), latent=torch.rand(2, 5, device=DEV))
    assert seen["shapes"][0] == ((64, 3), (64, 5)) and all(a[0] == b[0] for a, b in seen["shapes"])
    assert res["mask"].shape == (2, 32)
