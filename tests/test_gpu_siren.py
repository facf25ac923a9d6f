"""GPU parity of the fused SIREN SDF + input-gradient kernel (csrc/siren.cu) against what it
replaces: model.forward(x).sdf + autograd.grad (DSS/models/levelset_sampling.py:142-170) on the
reference's Siren decoder structure (DSS/models/common.py:56-165).

Ground truth is the same network evaluated through autograd in float64.  The bar for the fused
kernel is "the accuracy class of the fp32 autograd path it replaces": its error against float64
must be within a small factor of the error of torch's own fp32 (TF32 off) evaluation or below
1e-5 relative, i.e. far inside the 1e-4 relative tolerance north_star states."""
import numpy as np
import pytest
import torch

from isopoints_b200 import siren
from isopoints_b200.levelset_sampling import UniformProjection
from tests.helpers import Siren, SirenSDF

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref(model, x, dtype):
    m = Siren(256, len(model.net) - 2, float(model.net[0].omega_0)).to(dtype)
    m.load_state_dict({k: v.to(dtype).cpu() for k, v in model.state_dict().items()})
    m = m.to(DEV)
    xx = x.to(dtype).clone().requires_grad_(True)
    s = m(xx).sdf
    g, = torch.autograd.grad(s, xx, torch.ones_like(s))
    return s.detach().reshape(-1).double(), g.detach().double()


def _check(model, x, factor=4.0):
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        s64, g64 = _ref(model, x, torch.float64)
        s32, g32 = _ref(model, x, torch.float32)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out = siren.sdf_and_grad(model, x)
    assert out is not None, "model not recognised as a fusable SIREN"
    sf, gf = out
    assert sf.shape == (x.shape[0],) and gf.shape == (x.shape[0], 3)
    e_s, e_g = (sf.double() - s64).abs().max().item(), (gf.double() - g64).abs().max().item()
    r_s, r_g = (s32 - s64).abs().max().item(), (g32 - g64).abs().max().item()
    gmax = g64.abs().max().item()
    smax = max(s64.abs().max().item(), 1e-3)
    # same accuracy class as fp32 autograd (measured: ~3x its error, tensor-core accumulation is not
    # round-to-nearest), and two orders of magnitude inside north_star's 1e-4 relative bar
    assert e_s <= max(factor * r_s, 1e-6 + 1e-5 * smax), (e_s, r_s, smax)
    assert e_g <= max(factor * r_g, 3e-5 * gmax), (e_g, r_g, gmax)
    return e_s, e_g


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 148 * 128 + 5, 2 * 148 * 128 + 300])
def test_sdf_and_grad_match_fp64_autograd(n):
    model = Siren(256, 7, 30.0, seed=0).to(DEV)
    g = torch.Generator().manual_seed(n)
    x = ((torch.rand(n, 3, generator=g) - 0.5) * 2).to(DEV)
    _check(model, x)


@pytest.mark.parametrize("n_layers", [1, 2, 3, 5])
def test_layer_counts(n_layers):
    model = Siren(256, n_layers, 30.0, seed=n_layers).to(DEV)
    x = ((torch.rand(700, 3) - 0.5) * 2).to(DEV)
    _check(model, x)


def test_rescaled_weights_and_other_omega():
    """Per-layer weight magnitudes far from the init (what training produces): the power-of-two
    operand scaling must adapt per layer / per row."""
    model = Siren(256, 4, 12.0, seed=5).to(DEV)
    with torch.no_grad():
        for i, f in zip((1, 2, 3, 4), (3.0, 0.05, 1.7, 0.4)):
            model.net[i].linear.weight.mul_(f)
        model.net[-1].weight.mul_(250.0)
        model.net[-1].bias.fill_(-0.3)
    x = ((torch.rand(900, 3) - 0.5) * 2).to(DEV)
    _check(model, x, factor=6.0)


def test_repack_on_parameter_update():
    model = Siren(256, 2, 30.0, seed=1).to(DEV)
    x = ((torch.rand(300, 3) - 0.5) * 2).to(DEV)
    a = siren.sdf_and_grad(model, x)[0].clone()
    with torch.no_grad():
        model.net[1].linear.weight.add_(0.001)          # in-place update bumps ._version
    b = siren.sdf_and_grad(model, x)[0]
    assert (a - b).abs().max().item() > 1e-4
    _check(model, x)


def test_device_side_row_count():
    model = Siren(256, 3, 30.0, seed=2).to(DEV)
    x = ((torch.rand(1000, 3) - 0.5) * 2).to(DEV)
    full_s, full_g = siren.sdf_and_grad(model, x)
    n_dev = torch.tensor([517], dtype=torch.int32, device=DEV)
    spec = siren.match(model)
    assert spec is not None
    s, g = siren.sdf_and_grad(model, x, n_dev=n_dev)
    assert torch.equal(s[:517], full_s[:517]) and torch.equal(g[:517], full_g[:517])


def test_not_fusable_models_keep_autograd():
    assert siren.match(SirenSDF(seed=0).to(DEV)) is None            # different structure (opaque)
    assert siren.match(Siren(128, 2, 30.0).to(DEV)) is None          # other width
    m = Siren(256, 2, 30.0).to(DEV)
    assert siren.match(m, {"c": torch.ones(1, 4, device=DEV)}) is None   # latent code present
    assert siren.match(m, {"c": None}) is not None


def test_projection_with_fused_sdf_matches_autograd_path():
    """project + resample through the operator surface: fused SIREN vs the same weights as an
    opaque module (autograd path, which the golden-vector tests pin to the reference)."""
    torch.manual_seed(0)
    x = ((torch.rand(1, 6000, 3) - 0.5) * 2).to(DEV)
    fused, opaque = Siren(256, 3, 30.0, seed=3).to(DEV), SirenSDF(256, 3, 30.0, seed=3).to(DEV)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        outs = []
        for m in (fused, opaque):
            proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
            outs.append(proj.project_points(x.clone(), m, skip_resampling=True, skip_upsampling=True))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    a, b = outs
    ma, mb = a["mask"].cpu().numpy(), b["mask"].cpu().numpy()
    agree = ma == mb
    assert agree.mean() > 0.995          # points within float noise of the tolerance may flip
    both = (ma & mb)[0]
    pa, pb = a["levelset_points"][0].cpu().numpy()[both], b["levelset_points"][0].cpu().numpy()[both]
    # a random-init SIREN is chaotic (omega = 30 per layer): compare the bulk, bound the tail
    d = np.abs(pa - pb).max(axis=1)
    assert np.quantile(d, 0.99) < 1e-4 and np.median(d) < 2e-6


def test_project_resample_and_ragged_batch_with_fused_sdf():
    """The sync-free loop (device-side live counts) through filter + resample + re-projection, and
    on a ragged two-cloud batch, against the opaque-module path on the same weights."""
    torch.manual_seed(1)
    fused, opaque = Siren(256, 2, 30.0, seed=7).to(DEV), SirenSDF(256, 2, 30.0, seed=7).to(DEV)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        x = ((torch.rand(1, 5000, 3) - 0.5) * 2).to(DEV)
        outs = []
        for m in (fused, opaque):
            proj = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
            outs.append(proj.project_points(x.clone(), m, skip_upsampling=True))
        a, b = outs
        # survivors of the first projection may differ by a few borderline points, which shifts rows:
        # compare as point sets through nearest-neighbour distance
        pa, pb = a["levelset_points"][0][a["mask"][0]], b["levelset_points"][0][b["mask"][0]]
        assert abs(pa.shape[0] - pb.shape[0]) <= 0.01 * pb.shape[0]
        d = torch.cdist(pa, pb).min(dim=1).values
        assert d.median().item() < 1e-5 and d.quantile(0.98).item() < 1e-3
        # ragged batch straight into _project_points
        xr = ((torch.rand(2, 1500, 3) - 0.5) * 2).to(DEV)
        num = torch.tensor([1500, 900], device=DEV)
        ra = UniformProjection(proj_max_iters=8)._project_points(fused, xr.clone(), num)
        rb = UniformProjection(proj_max_iters=8)._project_points(opaque, xr.clone(), num)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    agree = (ra.mask == rb.mask)
    assert agree.float().mean().item() > 0.995
    both = ra.mask & rb.mask
    dd = (ra.points - rb.points).abs().max(dim=-1).values[both]
    assert dd.median().item() < 2e-6 and dd.quantile(0.99).item() < 1e-4
    assert not ra.mask[1, 900:].any() and (ra.points[1, 900:] == 0).all()


def test_fused_newton_iteration_equals_sdf_grad_plus_project_step():
    """isob200_siren_project_step == isob200_siren_sdf_grad followed by isob200_project_step: same
    positions / normals / flags bit for bit, the same SET of still-active rows (tiles append in
    completion order) with matching compacted positions."""
    from isopoints_b200 import _ext
    lib = _ext.lib()
    model = Siren(256, 3, 30.0, seed=9).to(DEV)
    spec = siren.match(model)
    blob, scratch, L = siren.packed(model, spec)
    torch.manual_seed(3)
    M = 20000
    pts0 = ((torch.rand(M, 3) - 0.5) * 2).to(DEV)
    st = _ext.stream(torch.device(DEV))
    tol = 5e-5

    def fused(points):
        normals = torch.zeros_like(points)
        nc = torch.ones(M, dtype=torch.uint8, device=DEV)
        act = torch.full((M,), -1, dtype=torch.int32, device=DEV)
        nxt = torch.zeros((M, 3), device=DEV)
        cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        _ext.check(lib.isob200_siren_project_step(
            _ext.ptr(points), M, None, _ext.ptr(blob), L, _ext.ptr(scratch), scratch.numel(), _ext.ptr(points),
            _ext.ptr(normals), _ext.ptr(nc), None, tol, 0.1, 1, _ext.ptr(act), _ext.ptr(nxt), _ext.ptr(cnt), st))
        return normals, nc, act, nxt, int(cnt.item())

    def split(points):
        sdf, grad = siren.sdf_and_grad(model, points.clone())
        normals = torch.zeros_like(points)
        nc = torch.ones(M, dtype=torch.uint8, device=DEV)
        act = torch.full((M,), -1, dtype=torch.int32, device=DEV)
        nxt = torch.zeros((M, 3), device=DEV)
        cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
        ws = _ext.workspace(lib.isob200_project_step_ws_bytes(M), torch.device(DEV))
        _ext.check(lib.isob200_project_step(
            _ext.ptr(points), _ext.ptr(normals), _ext.ptr(nc), None, M, None, _ext.ptr(sdf), _ext.ptr(grad), tol, 0.1, 1,
            _ext.ptr(act), _ext.ptr(nxt), _ext.ptr(cnt), _ext.ptr(ws), ws.numel(), st))
        return normals, nc, act, nxt, int(cnt.item())

    pa, pb = pts0.clone(), pts0.clone()
    na, fa, aa, xa, ca = fused(pa)
    nb, fb, ab, xb, cb = split(pb)
    assert ca == cb and 0 < ca < M
    assert torch.equal(pa, pb) and torch.equal(na, nb) and torch.equal(fa, fb)
    oa, ob = torch.argsort(aa[:ca]), torch.argsort(ab[:cb])
    assert torch.equal(aa[:ca][oa], ab[:cb][ob])                  # same active rows
    assert torch.equal(xa[:ca][oa], xb[:cb][ob])                  # with the same updated positions
    assert torch.equal(xa[:ca], pa[aa[:ca].long()])               # next_points == points[act_out]


@pytest.mark.parametrize("layers,n", [(1, 300), (2, 128 * 149 + 5), (7, 50000)])
def test_value_only_kernel_is_the_forward_half(layers, n):
    """isob200_siren_sdf (forward GEMMs only) returns the bits of isob200_siren_sdf_grad's value."""
    model = Siren(256, layers, 30.0, seed=layers + 10).to(DEV)
    x = ((torch.rand(n, 3, device=DEV) - 0.5) * 2).contiguous()
    both = siren.sdf_and_grad(model, x)
    only = siren.sdf(model, x)
    assert torch.equal(both[0], only)
    f = siren.sdf_fn(model)
    assert torch.equal(f(x.view(1, n, 3)), only.view(1, n))
    cnt = torch.tensor([n // 2], dtype=torch.int32, device=DEV)
    out = torch.full((n,), 7.0, device=DEV)
    siren.sdf(model, x, n_dev=cnt, out=out)
    assert torch.equal(out[: n // 2], only[: n // 2]) and bool((out[n // 2:] == 7.0).all())
