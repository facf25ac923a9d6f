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
