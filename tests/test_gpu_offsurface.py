"""GPU parity of the in-surface / off-surface sampler (isopoints_b200/offsurface.py, csrc/rays.cu) and of
``get_visible_points`` (isopoints_b200/ewa.py) -- reference: DSS/models/combined_modeling.py:237-388,
DSS/utils/__init__.py:699-711, DSS/core/cloud.py:289-367."""
import types

import numpy as np
import pytest
import torch

from isopoints_b200 import _ext, ewa, offsurface
from isopoints_b200.structures import Pointclouds
from oracle import port
from tests.helpers import PinholeCameras, Siren, TinySiren, make_surface_points, offsurface_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(decoder, points=None, thr=0.05):
    return types.SimpleNamespace(
        _points=points, decoder=decoder, max_points_per_pass=10000, object_bounding_sphere=1.0,
        renderer=types.SimpleNamespace(rasterizer=types.SimpleNamespace(
            raster_settings=types.SimpleNamespace(depth_merging_threshold=thr))))


@pytest.mark.parametrize("R,M,per_ray_origin", [(1, 1, False), (7, 33, True), (1000, 5000, False), (257, 2048 * 3 + 5, True),
                                                (4000, 20000, False)])
def test_point_to_ray_kernel_vs_dense_oracle(R, M, per_ray_origin):
    """Against the reference's dense (R,M) formulation.  ``|p - o|^2 - t^2`` cancels ~5 digits, so two float32
    evaluations of it differ by a few 1e-7 and can pick different points among near-ties: the kernel's choice
    must be a minimiser of the oracle's matrix up to that rounding, and its t^2 the oracle's at that point."""
    g = torch.Generator().manual_seed(R + M)
    pts = (torch.rand(M, 3, generator=g) - 0.5) * 1.2
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1)
    o = torch.nn.functional.normalize(torch.randn(R if per_ray_origin else 1, 3, generator=g), dim=-1) * 2.5
    l0 = _ext.lib().isob200_launch_count()
    t_sq, idx, dist = offsurface.closest_point_to_rays(o.to(DEV), d.to(DEV), pts.to(DEV), return_dist=True)
    assert _ext.lib().isob200_launch_count() == l0 + 1 and idx.dtype == torch.int64
    t_sq, idx, dist = t_sq.cpu(), idx.cpu(), dist.cpu()
    assert int(idx.min()) >= 0 and int(idx.max()) < M
    tol = 1e-5
    n_same = 0
    for r0 in range(0, R, 500):                       # the dense matrices of the reference, 500 rays at a time
        r1 = min(R, r0 + 500)
        if per_ray_origin:
            res = [port.ray_nearest_point(o[r], d[r:r + 1], pts) for r in range(r0, r1)]
            wi, dm, rs = (torch.cat([x[k] for x in res]) for k in (1, 2, 3))
        else:
            _, wi, dm, rs = port.ray_nearest_point(o[0], d[r0:r1], pts)
        mine = idx[r0:r1].view(-1, 1)
        d_mine = dm.gather(1, mine).view(-1)
        assert float((d_mine - dm.min(dim=1).values).max()) <= tol
        np.testing.assert_allclose(dist[r0:r1].numpy(), d_mine.numpy(), rtol=0, atol=tol)
        np.testing.assert_allclose(t_sq[r0:r1].numpy(), rs.gather(1, mine).view(-1).numpy(), rtol=1e-5, atol=1e-6)
        n_same += int((mine.view(-1) == wi).sum())
    assert n_same >= 0.9 * R


def test_point_to_ray_ties_and_empty():
    pts = torch.tensor([[0.0, 0.5, 1.0], [0.0, -0.5, 1.0], [0.0, 0.5, 1.0], [0.0, 0.0, -9.0]], device=DEV)
    d = torch.tensor([[0.0, 0.0, 1.0]], device=DEV)
    o = torch.zeros(3, device=DEV)
    t_sq, idx, dist = offsurface.closest_point_to_rays(o, d, pts, return_dist=True)
    assert idx.tolist() == [3] and t_sq.tolist() == [81.0] and dist.tolist() == [0.0]      # behind the camera counts
    t_sq, idx, dist = offsurface.closest_point_to_rays(o, d, pts[:3], return_dist=True)
    assert idx.tolist() == [0] and t_sq.tolist() == [1.0] and dist.tolist() == [0.25]      # ties: lowest index
    t_sq, idx = offsurface.closest_point_to_rays(o, d, pts[:0])
    assert idx.tolist() == [-1] and t_sq.tolist() == [0.0]
    t_sq, idx = offsurface.closest_point_to_rays(o, d[:0], pts)
    assert t_sq.shape == (0,) and idx.shape == (0,)
    with pytest.raises(ValueError):
        offsurface.closest_point_to_rays(torch.zeros(2, 3, device=DEV), torch.zeros(3, 3, device=DEV), pts)


def test_sampler_matches_reference_golden(golden):
    """The reference's Model.sample_offsurface_using_isopoints, run from the reference tree on the same inputs
    (its get_visible_points answered by the same preset point sets, its rand_like draw replayed)."""
    g = golden("offsurface")
    cams, pixels, mask_img, frontal, occluded, iso_pcl = offsurface_inputs()
    cams = PinholeCameras(cams.R.to(DEV), cams.T.to(DEV), cams.focal)
    answers = [Pointclouds([p.to(DEV) for p in frontal]), Pointclouds([p.to(DEV) for p in occluded])]
    calls = []

    def visible(points, cameras, depth_merge_threshold=0.05):
        calls.append((cameras.R.clone(), cameras.T.clone(), depth_merge_threshold))
        return answers[len(calls) - 1]
    R0 = cams.R.clone()
    p_off, p_ins, n_off, n_ins = offsurface.sample_offsurface_using_isopoints(
        _model(TinySiren(seed=3).to(DEV)), pixels.to(DEV), mask_img.to(DEV), cams,
        n_points_per_ray=int(g["n_points_per_ray"]), max_insurface_per_batch=g["max_insurface"].tolist(),
        iso_pcl=Pointclouds([p.to(DEV) for p in iso_pcl]), rand=torch.as_tensor(g["rand"], device=DEV),
        visible_points_fn=visible)
    assert np.array_equal(n_off.cpu().numpy(), g["n_off"]) and np.array_equal(n_ins.cpu().numpy(), g["n_ins"])
    np.testing.assert_allclose(p_off.cpu().numpy(), g["p_off"], rtol=0, atol=1e-5)
    close = np.abs(p_ins.cpu().numpy() - g["p_ins"]).max(axis=1) < 1e-4      # arg-min over candidates: near-ties
    assert close.mean() > 0.98, close.mean()
    assert len(calls) == 2 and calls[0][2] == 0.05 and torch.equal(cams.R, R0)            # the input cameras are untouched
    np.testing.assert_allclose(calls[1][0].cpu().numpy(), g["back_R"], atol=1e-6)
    np.testing.assert_allclose(calls[1][1].cpu().numpy(), g["back_T"], atol=1e-5)


def _sphere_cloud(n, seed, radius=0.6):
    pts, nrm, _, _ = make_surface_points([n], seed=seed, noise=0.002)
    pts = pts * (radius / 0.8)
    return Pointclouds([pts.to(DEV)], normals=[torch.nn.functional.normalize(pts, dim=-1).to(DEV)])


def test_get_visible_points_is_the_visibility_of_the_256_splat():
    pc = _sphere_cloud(20000, seed=3)
    cams = PinholeCameras.look_at_origin(1, seed=5, device=DEV)
    vis_pc, mask = ewa.get_visible_points(pc, cams, depth_merge_threshold=0.05, return_mask=True)
    assert tuple(mask.shape) == (1, 20000) and mask.dtype == torch.bool
    assert int(mask.sum()) == int(vis_pc.num_points_per_cloud()[0])
    assert torch.equal(vis_pc.points_list()[0], pc.points_list()[0][mask[0]])
    assert torch.equal(vis_pc.normals_list()[0], pc.normals_list()[0][mask[0]])
    # the reference's definition, restated on the fragments of the same settings (DSS/utils/__init__.py:378-399)
    rs = ewa.PointsRasterizationSettings(depth_merging_threshold=0.05, image_size=256, cutoff_threshold=1.0,
                                         backface_culling=True)
    ras = ewa.SurfaceSplatting(cameras=cams, raster_settings=rs)
    frag, filtered = ras(pc)
    in_filter = ras.filter_renderable(pc)[1]
    ids = frag.idx[frag.occupancy.bool()].unique()
    ids = ids[ids >= 0]
    want = torch.zeros(20000, dtype=torch.bool, device=DEV)
    want[in_filter.nonzero().view(-1)[ids.long()]] = True
    assert torch.equal(mask[0], want)
    # only front-facing points of the near hemisphere can be visible
    centre = cams.get_camera_center()[0]
    assert bool(((vis_pc.points_list()[0] * centre).sum(-1) > -0.05).all())
    assert 0.2 < mask.float().mean().item() < 0.6


def test_sampler_end_to_end_on_the_fused_siren_decoder():
    """Real visibility passes + point-to-ray kernel + forward-only SIREN kernel through one call."""
    from isopoints_b200 import siren
    pc = _sphere_cloud(30000, seed=8)
    cams = PinholeCameras.look_at_origin(2, seed=9, device=DEV)
    g = torch.Generator().manual_seed(0)
    pixels = (torch.rand(2, 2000, 2, generator=g) * 1.6 - 0.8).to(DEV)
    S = 64
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, S), torch.linspace(-1, 1, S), indexing="ij")
    mask_img = ((xx ** 2 + yy ** 2) < 0.3 ** 2).float().expand(2, 1, S, S).contiguous().to(DEV)
    decoder = Siren(256, 2, 30.0, seed=2).to(DEV)
    calls0 = siren.STATS["calls"]
    p_off, p_ins, n_off, n_ins = offsurface.sample_offsurface_using_isopoints(
        _model(decoder, pc.extend(2)), pixels, mask_img, cams, n_points_per_ray=16, max_insurface_per_batch=[300, 300])
    assert siren.STATS["calls"] > calls0
    assert p_off.shape == (int(n_off.sum()), 3) and p_ins.shape == (int(n_ins.sum()), 3)
    assert int(n_ins.min()) > 100 and int(n_ins.max()) <= 300
    # in-surface samples lie between the front and the back of the sphere of iso-points
    assert float(p_ins.norm(dim=-1).max()) < 0.62
    assert float(p_off.abs().max()) <= 1.05 + 1e-5                 # off-surface samples stay in the padded cube
    none = offsurface.sample_offsurface_using_isopoints(_model(decoder, pc.extend(2)), pixels, mask_img, cams)
    assert none[1].shape == (0, 3) and none[3].tolist() == [0, 0]


@pytest.mark.parametrize("cap,n_iso", [(2000, 20000), (6000, 9000)])
def test_get_visible_iso_points(cap, n_iso):
    """Model.get_visible_iso_points (combined_modeling.py:390-455): thinning (visible > cap) and topping up
    (visible < 0.75 cap) branches; the result lies on the level set, faces the camera and carries normals."""
    from isopoints_b200.levelset_sampling import UniformProjection
    from tests.helpers import SphereSDF
    pc = _sphere_cloud(n_iso, seed=4)
    cams = PinholeCameras.look_at_origin(2, seed=6, device=DEV)
    model = _model(SphereSDF(radius=0.6).to(DEV), pc)
    model.max_iso_per_batch = cap
    model.device = torch.device(DEV)
    model.projection = UniformProjection(proj_max_iters=10, proj_tolerance=5e-5)
    g = torch.Generator().manual_seed(0)
    iso = offsurface.get_visible_iso_points(model, cams, generator=g)
    assert len(iso) == 2 and iso.normals_packed() is not None
    centre = cams.get_camera_center()
    n_vis0 = ewa.get_visible_points(pc.extend(2), cams).num_points_per_cloud().tolist()
    for b in range(2):
        p, n = iso.points_list()[b], iso.normals_list()[b]
        assert 0.5 * min(cap, n_vis0[b]) < p.shape[0] <= max(cap, n_vis0[b])
        assert float((p.norm(dim=-1) - 0.6).abs().max()) < 1e-4                       # on the level set
        assert bool(((p * centre[b]).sum(-1) > -0.05).all())                          # on the camera's side (back-face cull)
        np.testing.assert_allclose(n.cpu().numpy(), (p / 0.6).cpu().numpy(), atol=1e-3)   # normals = SDF gradient
    if cap == 2000:
        assert max(n_vis0) > cap                 # the thinning branch ran
    else:
        assert max(n_vis0) < 0.75 * cap          # the topping-up branch ran
    model.max_iso_per_batch = 0
    assert tuple(offsurface.get_visible_iso_points(model, cams).shape) == (1, 0, 3)


def test_subsample_randomly():
    pc = Pointclouds([torch.rand(100, 3, device=DEV), torch.rand(40, 3, device=DEV)],
                     normals=[torch.rand(100, 3, device=DEV), torch.rand(40, 3, device=DEV)])
    out = offsurface.subsample_randomly(pc, torch.tensor([0.25, 2.0]), torch.Generator().manual_seed(1))
    assert out.num_points_per_cloud().tolist() == [25, 40] and out.normals_packed().shape == (65, 3)
    src = {tuple(r) for r in pc.points_list()[0].cpu().numpy().round(6).tolist()}
    assert all(tuple(r) in src for r in out.points_list()[0].cpu().numpy().round(6).tolist())
    assert offsurface.subsample_randomly(pc, 1.0).num_points_per_cloud().tolist() == [100, 40]


def test_single_cloud_is_extended_to_the_camera_count_like_the_reference():
    """SurfaceSplatting.forward extends a single cloud to len(cameras) at the top (rasterizer.py:597-598): the
    visibility filter has one row per VIEW, get_visible_points returns one cloud per view, and the sampler accepts
    the model's single-cloud `_points` with B = 2 cameras."""
    from isopoints_b200 import siren
    pc = _sphere_cloud(20000, seed=8)
    assert len(pc) == 1
    cams = PinholeCameras.look_at_origin(2, seed=9, device=DEV)
    vis1, m1 = ewa.get_visible_points(pc, cams, return_mask=True)
    vis2, m2 = ewa.get_visible_points(pc.extend(2), cams, return_mask=True)
    assert m1.shape == m2.shape == (2, 20000) and torch.equal(m1, m2)
    assert len(vis1) == 2 and vis1.num_points_per_cloud().tolist() == vis2.num_points_per_cloud().tolist()
    assert not torch.equal(m1[0], m1[1])                    # two different views, not view 0 twice
    # union over views (what get_visible_iso_points computes for the reference cloud, combined_modeling.py:405-409)
    assert int(m1.any(dim=0).sum()) > int(m1[0].sum())
    g = torch.Generator().manual_seed(0)
    pixels = (torch.rand(2, 1500, 2, generator=g) * 1.6 - 0.8).to(DEV)
    S = 64
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, S), torch.linspace(-1, 1, S), indexing="ij")
    mask_img = ((xx ** 2 + yy ** 2) < 0.3 ** 2).float().expand(2, 1, S, S).contiguous().to(DEV)
    decoder = Siren(256, 2, 30.0, seed=2).to(DEV)
    a = offsurface.sample_offsurface_using_isopoints(_model(decoder, pc), pixels, mask_img, cams,
                                                     n_points_per_ray=16, max_insurface_per_batch=[200, 200])
    b = offsurface.sample_offsurface_using_isopoints(_model(decoder, pc.extend(2)), pixels, mask_img, cams,
                                                     n_points_per_ray=16, max_insurface_per_batch=[200, 200])
    assert a[3].tolist() == b[3].tolist() and int(a[3].min()) > 50
