"""True drop-in (north_star: "drops into train_mvr.py unchanged"): the reference's OWN Python -- loaded unmodified
from the reference tree (/root/reference, or the staged copy under baseline/_ref on the GPU box) -- running on top
of `isopoints_b200.install()`, i.e. with `frnn`, `prefix_sum` and `DSS._C` resolved to libisob200.so, reproduces
the golden vectors its own natives produced and the results of this package's operators."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_python
from tests.helpers import SphereSDF, make_splat_inputs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_python.available(), reason="reference tree not staged")]
DEV = "cuda"


@pytest.fixture(scope="module")
def dropin():
    from isopoints_b200 import install
    mods = install.install()
    ref = ref_python.load()
    ref.levelset_sampling.frnn = mods["frnn"]
    ref.point_processing.frnn = mods["frnn"]
    yield types.SimpleNamespace(mods=mods, ref=ref)
    install.uninstall()


def test_reference_uniform_projection_on_installed_frnn_matches_golden(dropin, golden):
    """DSS.models.levelset_sampling.UniformProjection (the reference's Python loop) with `import frnn` -> ours."""
    g = golden("resample_sphere")
    LS = dropin.ref.levelset_sampling
    proj = LS.UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    out = proj.project_points(torch.as_tensor(g["x"], device=DEV), SphereSDF().to(DEV), skip_upsampling=True)
    assert np.array_equal(out["mask"].cpu().numpy(), g["mask"])
    np.testing.assert_allclose(out["levelset_points"].cpu().numpy(), g["points"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["levelset_normals"].cpu().numpy(), g["normals"], rtol=1e-4, atol=1e-5)
    # ... and equals this package's own operator on the same input
    from isopoints_b200.levelset_sampling import UniformProjection
    mine = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1).project_points(
        torch.as_tensor(g["x"], device=DEV), SphereSDF().to(DEV), skip_upsampling=True)
    np.testing.assert_allclose(mine["levelset_points"].cpu().numpy(), out["levelset_points"].cpu().numpy(),
                               rtol=1e-4, atol=1e-5)


def test_reference_frnn_python_on_installed_C_is_bit_exact(dropin):
    """external/FRNN/frnn/frnn.py (the reference's host sequence: grid params loop, insert, prefix sum, counting
    sort, find_nbrs) driving OUR `frnn._C` + `prefix_sum`: same idx / dists as our frnn_grid_points."""
    path = os.path.join(ref_python.REF, "external", "FRNN", "frnn", "frnn.py")
    spec = importlib.util.spec_from_file_location("ref_frnn_py_on_isob200", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                      # `from frnn import _C`, `from prefix_sum import ...` -> ours
    assert mod._C is dropin.mods["frnn"]._C
    torch.manual_seed(0)
    p = torch.rand(2, 20_000, 3, device=DEV)
    lens = torch.tensor([20_000, 15_000], device=DEV)
    d_ref, i_ref, _, grid = mod.frnn_grid_points(p, p, lens, lens, K=9, r=0.04, return_nn=False)
    from isopoints_b200 import frnn as ours
    d, i, _, _ = ours.frnn_grid_points(p, p, lens, lens, K=9, r=0.04, return_nn=False)
    assert torch.equal(i, i_ref) and torch.equal(d, d_ref)
    q = torch.rand(2, 3_000, 3, device=DEV)           # grid reuse with other queries (frnn.py:84-98)
    d2_ref, i2_ref, _, _ = mod.frnn_grid_points(q, p, None, lens, K=5, r=0.04, grid=grid)
    d2, i2, _, _ = ours.frnn_grid_points(q, p, None, lens, K=5, r=0.04)
    assert torch.equal(i2, i2_ref) and torch.equal(d2, d2_ref)


def test_reference_rasterizer_python_on_installed_natives(dropin, golden):
    """DSS.core.rasterizer.rasterize_elliptical_points + EllipticalRasterizer.backward (the reference's Python:
    visible filter, per-view median radius, 2-D grid through frnn._C / prefix_sum, occupancy kernel, scatter,
    rasterizer.py:743-973) on OUR DSS._C / frnn._C / prefix_sum: forward equals the golden of the reference's own
    CPU twin, forward + backward equal this package's operator."""
    rast = ref_python.load_rasterizer()
    rast._C = dropin.mods["DSS._C"]
    rast.frnn = dropin.mods["frnn"]
    g = golden("splat_naive_cpu")
    S, K = int(g["S"]), int(g["K"])
    t = {k: torch.as_tensor(g[k], device=DEV) for k in ("points", "ellipse", "cutoff", "radii", "first_idx", "num_points")}

    def screen(pts, tt):
        return types.SimpleNamespace(points_packed=lambda: pts, cloud_to_packed_first_idx=lambda: tt["first_idx"],
                                     num_points_per_cloud=lambda: tt["num_points"])
    idx, zbuf, qv, occ = rast.rasterize_elliptical_points(screen(t["points"], t), t["ellipse"], t["cutoff"], t["radii"],
                                                          depth_merging_threshold=0.05, image_size=S,
                                                          points_per_pixel=K, bin_size=0)
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    np.testing.assert_allclose(zbuf.cpu().numpy(), g["zbuf"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(occ.cpu().numpy(), g["occ"], rtol=1e-6, atol=1e-6)
    # forward + backward at a larger size, against this package's own autograd Function
    from isopoints_b200 import splat
    V, S, K = 3, 128, 6
    inp = make_splat_inputs(V, [9000, 7000, 8000], S, seed=9, sigma_px=1.5)
    tt = {k: torch.as_tensor(v, device=DEV) for k, v in inp.items()}
    gg = torch.Generator().manual_seed(1)
    occ_grad = (torch.randn(V, S, S, generator=gg) * (torch.rand(V, S, S, generator=gg) < 0.2)).to(DEV)
    zbuf_grad = torch.randn(V, S, S, K, generator=gg).to(DEV)
    grads, outs = [], []
    for fn in (rast.rasterize_elliptical_points, splat.rasterize_elliptical_points):
        pts = tt["points"].clone().requires_grad_(True)
        o = fn(screen(pts, tt), tt["ellipse"], tt["cutoff"], tt["radii"], depth_merging_threshold=0.05,
               image_size=S, points_per_pixel=K, radii_backward_scaler=10.0)
        ((o[3] * occ_grad).sum() + (o[1] * zbuf_grad).sum()).backward()
        grads.append(pts.grad.clone())
        outs.append(o)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    np.testing.assert_allclose(grads[0].cpu().numpy(), grads[1].cpu().numpy(), rtol=1e-4,
                               atol=1e-4 * float(grads[1].abs().max()))
    assert float(grads[0].abs().max()) > 0


def test_reference_wlop_and_upsample_python_on_installed_frnn(dropin, golden):
    """DSS.utils.point_processing.wlop / upsample (reference Python) on our frnn; pytorch3d's knn_points (absent
    here) is answered by this package's exact grid K-NN."""
    from isopoints_b200 import point_processing as pp
    from isopoints_b200.structures import Pointclouds
    g = golden("wlop_upsample")
    PP = dropin.ref.point_processing
    PP.knn_points = pp.knn_points
    PP.Pointclouds = Pointclouds
    noise = torch.as_tensor(g["noise"], device=DEV)
    real = torch.randn_like
    PP.torch.randn_like = lambda x: noise.clone()
    try:
        wl = PP.wlop(Pointclouds(torch.as_tensor(g["P"], device=DEV)), ratio=1.0, neighborhood_size=16, iters=3,
                     repulsion_mu=0.5)
    finally:
        PP.torch.randn_like = real
    np.testing.assert_allclose(wl.points_padded().cpu().numpy(), g["wlop"], rtol=1e-4, atol=5e-6)
    up, num = PP.upsample(torch.as_tensor(g["up_in"], device=DEV), 1300, num_points=torch.tensor([1000], device=DEV),
                          neighborhood_size=16)
    assert int(num[0]) == 1300
    # rows 300.. = the 1000 input points; rows ..300 = inserted mid-points: the sparsity ranking is computed by the
    # reference's torch ops on the GPU here and on the CPU in the golden, so a near-tie may pick another mid-point
    want = torch.as_tensor(g["up_pts"], device=DEV)
    np.testing.assert_allclose(up[0, 300:].cpu().numpy(), g["up_pts"][0, 300:], rtol=1e-4, atol=2e-6)
    same = torch.isclose(up[0, :300], want[0, :300], rtol=1e-4, atol=2e-6).all(-1)
    assert float(same.float().mean()) > 0.97
    d = pp.knn_points(up[:, :300], want[:, :300], K=1).dists[0, :, 0].sqrt()
    assert float(d.max()) < 0.1
