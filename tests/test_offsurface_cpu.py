"""Host-side pieces of the off-surface / in-surface sampler (isopoints_b200/offsurface.py, cloud.py) against the
reference's own functions (golden: tests/golden/make_golden.py --only offsurface).  The full
``sample_offsurface_using_isopoints`` needs the point-to-ray kernel and runs in tests/test_gpu_offsurface.py."""
import numpy as np
import pytest
import torch

from isopoints_b200 import offsurface
from isopoints_b200.cloud import PointCloudsFilters
from isopoints_b200.structures import Pointclouds
from tests.helpers import PinholeCameras, offsurface_inputs


def test_cube_intersection_matches_reference(golden):
    g = golden("offsurface")
    cams = PinholeCameras.look_at_origin(2, seed=41, focal=2.0)
    rays = torch.as_tensor(g["cube_rays"])
    c0, c1, m = offsurface.intersection_with_unit_cube(cams.get_camera_center().view(-1, 1, 3), rays, side_length=2.0)
    assert np.array_equal(m.numpy(), g["cube_mask"])
    np.testing.assert_allclose(c0.numpy(), g["cube0"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(c1.numpy(), g["cube1"], rtol=0, atol=1e-6)
    # a ray that misses the cube, and one along an axis (division by zero -> inf / nan never "on a face")
    o = torch.tensor([[3.0, 3.0, 3.0], [0.0, 0.0, 3.0]])
    d = torch.nn.functional.normalize(torch.tensor([[1.0, 0.2, 0.1], [0.0, 0.0, -1.0]]), dim=-1)
    c0, c1, m = offsurface.intersection_with_unit_cube(o, d, side_length=2.0)
    assert m.tolist() == [False, True] and bool((c0[0] == 0).all())
    np.testing.assert_allclose(c0[1].numpy(), [0, 0, 1.05], atol=1e-6)
    np.testing.assert_allclose(c1[1].numpy(), [0, 0, -1.05], atol=1e-6)


def test_tensor_values_match_reference(golden):
    g = golden("offsurface")
    _, pixels, mask_img, *_ = offsurface_inputs()
    vals = offsurface.get_tensor_values(mask_img, pixels.clamp(-1, 1), squeeze_channel_dim=True)
    np.testing.assert_allclose(vals.numpy(), g["mask_values"], rtol=0, atol=1e-6)


def test_point_cloud_filters():
    pts = [torch.rand(5, 3), torch.rand(3, 3)]
    nrm = [torch.rand(5, 3), torch.rand(3, 3)]
    pc = Pointclouds(pts, normals=nrm)
    f = PointCloudsFilters()
    assert f.filter(pc).num_points_per_cloud().tolist() == [5, 3]            # the (1,1) True default keeps all
    vis = torch.tensor([[1, 0, 1, 0, 1], [1, 1, 1, 1, 1]]).bool()           # padded slots of cloud 1 set on purpose
    f.set_filter(visibility=vis)
    q = f.filter_with(pc, ("visibility",))
    assert q.num_points_per_cloud().tolist() == [3, 3]
    assert torch.equal(q.points_list()[0], pts[0][[0, 2, 4]]) and torch.equal(q.normals_list()[0], nrm[0][[0, 2, 4]])
    assert torch.equal(q.points_list()[1], pts[1])
    f.set_filter(activation=torch.tensor([[0, 1, 1, 1, 1]]).bool())           # one row broadcasts over the clouds
    assert f.filter(pc).num_points_per_cloud().tolist() == [2, 2]
    assert f.filter_with(pc, ("activation",)).num_points_per_cloud().tolist() == [4, 2]
    one = Pointclouds([pts[0]])                                               # one cloud, per-view filters -> N views
    f2 = PointCloudsFilters()
    f2.set_filter(visibility=torch.tensor([[1, 1, 0, 0, 0], [0, 0, 0, 1, 1], [0, 0, 0, 0, 0]]).bool())
    out = f2.filter_with(one, ("visibility",))
    assert len(out) == 3 and out.num_points_per_cloud().tolist() == [2, 2, 0]
    with pytest.raises(ValueError):
        f.set_filter(visibility=torch.ones(5, dtype=torch.bool))
    with pytest.raises(AttributeError):
        f.set_filter(colour=torch.ones(1, 1, dtype=torch.bool))
    empty = Pointclouds([torch.zeros(0, 3)])
    assert f.filter_with(empty, ("visibility",)) is empty


def test_kernel_entry_points_refuse_cpu_tensors():
    with pytest.raises(TypeError):
        offsurface.closest_point_to_rays(torch.zeros(3), torch.zeros(4, 3), torch.zeros(5, 3))


def test_sampler_host_sequence_matches_reference_golden(golden, monkeypatch):
    """The host sequence of ``sample_offsurface_using_isopoints`` on CPU tensors, with the point-to-ray kernel
    stood in for by the oracle's dense formulation (test-only: the product entry point has no CPU path, see
    test_kernel_entry_points_refuse_cpu_tensors) -- must reproduce the reference's own method exactly."""
    import types
    from oracle import port
    from tests.helpers import TinySiren
    g = golden("offsurface")
    cams, pixels, mask_img, frontal, occluded, iso_pcl = offsurface_inputs()

    def dense(origins, rays, points, return_dist=False):
        t_sq, idx, _, _ = port.ray_nearest_point(origins.view(3), rays, points)
        return t_sq, idx
    monkeypatch.setattr(offsurface, "closest_point_to_rays", dense)
    answers = [Pointclouds(frontal), Pointclouds(occluded)]
    calls = []

    def visible(points, cameras, depth_merge_threshold=0.05):
        calls.append((cameras.R.clone(), cameras.T.clone()))
        return answers[len(calls) - 1]
    model = types.SimpleNamespace(
        _points=None, decoder=TinySiren(seed=3), max_points_per_pass=10000, object_bounding_sphere=1.0,
        renderer=types.SimpleNamespace(rasterizer=types.SimpleNamespace(
            raster_settings=types.SimpleNamespace(depth_merging_threshold=0.05))))
    R0, T0 = cams.R.clone(), cams.T.clone()
    p_off, p_ins, n_off, n_ins = offsurface.sample_offsurface_using_isopoints(
        model, pixels, mask_img, cams, n_points_per_ray=int(g["n_points_per_ray"]),
        max_insurface_per_batch=g["max_insurface"].tolist(), iso_pcl=Pointclouds(iso_pcl),
        rand=torch.as_tensor(g["rand"]), visible_points_fn=visible)
    assert np.array_equal(n_off.numpy(), g["n_off"]) and np.array_equal(n_ins.numpy(), g["n_ins"])
    np.testing.assert_allclose(p_off.numpy(), g["p_off"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(p_ins.numpy(), g["p_ins"], rtol=0, atol=1e-6)
    # the mirrored camera handed to the second visibility pass; the caller's cameras are left alone
    np.testing.assert_allclose(calls[1][0].numpy(), g["back_R"], atol=1e-7)
    np.testing.assert_allclose(calls[1][1].numpy(), g["back_T"], atol=1e-6)
    assert torch.equal(cams.R, R0) and torch.equal(cams.T, T0)
    # the mirrored camera sits at the antipode of the original one
    back = PinholeCameras(torch.as_tensor(g["back_R"]), torch.as_tensor(g["back_T"]))
    np.testing.assert_allclose(back.get_camera_center().numpy(), -cams.get_camera_center().numpy(), atol=1e-5)


def test_visible_iso_points_host_sequence(monkeypatch):
    """Branch logic of ``get_visible_iso_points`` (combined_modeling.py:390-455) on CPU tensors with the three
    kernel-backed stages stood in for: which clouds are thinned / topped up / kept (incl. the 0.8 factors with a
    reference cloud), where the jitter goes, and what reaches the projection and the final visibility pass."""
    import types
    from isopoints_b200 import ewa, point_processing
    g = torch.Generator().manual_seed(0)
    base = torch.rand(1000, 3, generator=g)
    seen_counts = [900, 500, 200]                       # per view: > cap, within [0.75 cap, cap], < 0.75 cap
    log = {"visible": [], "upsample": [], "project": None}

    def fake_visible(pcl, cameras, depth_merge_threshold=0.05, return_mask=False):
        log["visible"].append((len(pcl), pcl.num_points_per_cloud().tolist(), depth_merge_threshold))
        if len(log["visible"]) == 1:                    # the first pass: per-view subsets of the model's points
            keep = [torch.arange(1000) < n for n in seen_counts]
            out = Pointclouds([pcl.points_list()[b][keep[b]] for b in range(3)])
            return (out, torch.stack(keep)) if return_mask else out
        return pcl                                      # the last pass: everything stays visible

    def fake_upsample(pcl, n_points):
        log["upsample"].append((pcl.num_points_per_cloud().tolist(), n_points))
        p = pcl.points_packed()
        reps = -(-n_points // p.shape[0])
        return Pointclouds([p.repeat(reps, 1)[:n_points]])

    class FakeProjection:
        def project_points(self, pcl, decoder, skip_resampling=False, skip_upsampling=False, **kw):
            log["project"] = (pcl.num_points_per_cloud().tolist(), skip_resampling, skip_upsampling, sorted(kw))
            pts = pcl.points_padded()
            n = pcl.num_points_per_cloud()
            mask = torch.arange(pts.shape[1])[None, :] < n[:, None]
            mask[:, 0] = False                          # one point per view fails to converge
            return {"levelset_points": pts + 1.0, "levelset_normals": torch.ones_like(pts), "mask": mask}

    monkeypatch.setattr(ewa, "get_visible_points", fake_visible)
    monkeypatch.setattr(point_processing, "upsample", fake_upsample)
    cams = PinholeCameras.look_at_origin(3, seed=2)
    model = types.SimpleNamespace(
        _points=Pointclouds([base], normals=[torch.ones(1000, 3)]), decoder=None, max_iso_per_batch=600,
        projection=FakeProjection(), device=torch.device("cpu"),
        renderer=types.SimpleNamespace(rasterizer=types.SimpleNamespace(
            raster_settings=types.SimpleNamespace(depth_merging_threshold=0.07))))
    jitter = torch.full((600 + 500 + 600, 3), 0.5)      # zero offset after the (u - 0.5) shift
    out = offsurface.get_visible_iso_points(model, cams, jitter=jitter, generator=torch.Generator().manual_seed(3))
    assert log["visible"][0] == (3, [1000, 1000, 1000], 0.07)                # the model's cloud, one copy per view
    assert log["upsample"] == [([200], 600)]                                 # only the sparse view is topped up
    assert log["project"] == ([600, 500, 600], True, True, [])               # thinned / kept / topped up; no resampling
    assert log["visible"][1][1] == [599, 499, 599]                           # converged points only
    assert out.num_points_per_cloud().tolist() == [599, 499, 599] and out.normals_packed() is not None
    kept_mid = out.points_list()[1] - 1.0                                    # the untouched view: points 1..499 of base
    assert torch.allclose(kept_mid, base[1:500], atol=1e-6)
    thinned = out.points_list()[0] - 1.0                                     # a random subset of the 900 seen points
    gap = (thinned[:, None, :] - base[None, :900, :]).abs().amax(-1)              # (599, 900)
    assert float(gap.min(dim=1).values.max()) < 1e-6
    assert gap.argmin(dim=1).unique().numel() == thinned.shape[0]                   # no point taken twice
    # with a reference cloud the bounds shrink by 0.8 (int(0.8 * 600) = 480, int(0.8 * 450) = 360) and the
    # projection is allowed to upsample
    log["visible"].clear(); log["upsample"].clear()
    ref = Pointclouds([torch.rand(50, 3, generator=g)], normals=[torch.ones(50, 3)])
    ref_seen = torch.arange(50)[None, :] < torch.tensor([[10], [20], [0]])     # per view; the union keeps 20 points

    def fake_visible_ref(pcl, cameras, depth_merge_threshold=0.05, return_mask=False):
        if pcl.num_points_per_cloud().tolist() == [50]:                          # the reference cloud's own pass
            return pcl, ref_seen
        return fake_visible(pcl, cameras, depth_merge_threshold, return_mask)
    monkeypatch.setattr(ewa, "get_visible_points", fake_visible_ref)
    seen_ref = {}

    class RefProjection(FakeProjection):
        def project_points(self, pcl, decoder, **kw):
            seen_ref["n"] = kw["ref_pcl"].num_points_per_cloud().tolist()
            return super().project_points(pcl, decoder, **kw)
    model.projection = RefProjection()
    jitter = torch.full((480 + 480 + 480, 3), 0.5)
    offsurface.get_visible_iso_points(model, cams, jitter=jitter, ref_pcl=ref)
    assert log["upsample"] == [([200], 480)]
    counts, skip_rs, skip_up, extra = log["project"]
    assert counts == [480, 480, 480] and skip_rs and not skip_up and extra == ["ref_pcl"]
    assert seen_ref["n"] == [20]                      # the reference cloud reduced to what any camera sees
    model.max_iso_per_batch = 0
    assert tuple(offsurface.get_visible_iso_points(model, cams).shape) == (1, 0, 3)
