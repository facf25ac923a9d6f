"""The view-invariant V_k^r setting (``Vrk_invariant=True``; reference ``_compute_global_Vrk``,
DSS/core/rasterizer.py:292-343) against golden vectors made by the reference's own ``_get_per_point_info``
(tests/golden/make_golden.py --only ewa_global).  CPU part: the per-cloud h (a pure function of the FRNN distances)
and the claim the product relies on -- the invariant variant is the isotropic parameter computation fed with that
per-cloud h -- checked through the oracle of the parameter kernel.  GPU part: tests/test_gpu_ewa.py."""
import numpy as np
import pytest
import torch

from isopoints_b200 import ewa
from oracle import port
from tests.helpers import make_cameras, make_surface_points


def _inputs(g, tag):
    num = g[tag + "_num"].tolist()
    seed = int(g[tag + "_seed"])
    pts, nrm, first, numt = make_surface_points(num, seed=seed)
    w2v, proj, nmat = make_cameras(len(num), seed=seed + 1, znear=float(g["znear"]), zfar=float(g["zfar"]))
    return pts, nrm, first, numt, proj, num


def _padded(pts, first, num):
    out = torch.zeros(len(num), max(num), 3)
    for b, (f, n) in enumerate(zip(first.tolist(), num)):
        out[b, :n] = pts[f:f + n]
    return out


@pytest.mark.parametrize("tag", ["ragged", "equal"])
def test_per_cloud_h_and_parameters_match_reference(golden, tag):
    g = golden("ewa_point_info_global")
    pts, nrm, first, numt, proj, num = _inputs(g, tag)
    padded = _padded(pts, first, num).numpy()
    _, d = port.frnn_bruteforce(padded, padded, numt.numpy(), numt.numpy(), K=7, r=float(g["frnn_radius"]))
    h_cloud = ewa.global_vrk_h_from_sq_dist(torch.as_tensor(d), numt)
    np.testing.assert_allclose(h_cloud.numpy(), g[tag + "_h_cloud"], rtol=1e-6, atol=0)
    if tag == "ragged":       # clamp high / the `< 7 points` rule / dragged under the lower clamp by -1 padding rows
        assert h_cloud.tolist() == pytest.approx([1e-3, 5e-4, 5e-5], rel=1e-6)
    else:                     # inside the clamp range: the mean itself is checked
        assert all(5e-5 < v < 1e-3 for v in h_cloud.tolist())
    h = torch.repeat_interleave(h_cloud, numt)
    radii, ellipse, cutoff, scaler, _ = port.ewa_point_params(
        pts, nrm, first.tolist(), num, proj, h, int(g["image_size"]), float(g["antialiasing_sigma"]), float(g["cutoff"]))
    step = int(g[tag + "_step"])
    for got, want, tol in ((radii, g[tag + "_radii"], 5e-4), (ellipse, g[tag + "_ellipse"], 5e-4)):
        got = got.numpy()[::step]
        sc = np.abs(want).max(axis=-1, keepdims=True)
        assert float((np.abs(got - want) / sc).max()) < tol
    w = np.abs(g[tag + "_scaler"].astype(np.float64))
    assert (np.abs(scaler.numpy()[::step] - g[tag + "_scaler"]) <= 5e-4 * w + 1e-5 * np.median(w)).all()


def test_setting_precedence():
    """Vrk_invariant is looked at before Vrk_isotropic (rasterizer.py:418-424); anisotropic stays unbuilt."""
    rs = ewa.PointsRasterizationSettings(Vrk_invariant=False, Vrk_isotropic=False)
    with pytest.raises(NotImplementedError):
        ewa.SurfaceSplatting(raster_settings=rs)._get_per_point_info(None)
    with pytest.raises(NotImplementedError):
        ewa.compute_global_vrk_h(torch.zeros(1, 8, 3), torch.tensor([8]), -1.0)
    with pytest.raises(TypeError):      # cpu tensors
        ewa.compute_global_vrk_h(torch.zeros(1, 8, 3), torch.tensor([8]), 0.2)
