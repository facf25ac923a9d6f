"""Host logic of the RayTracing mirror (isopoints_b200/ray_tracing.py) against golden vectors made by running
the reference's own ``RayTracing.forward`` (levelset_sampling.py:810-1167; tests/golden/make_golden.py --only
rays).  The module has no kernel of its own -- the SDF is a callable -- so its bookkeeping (bounds, two-sided
march with step-back, sampler + secant, minimal-SDF search) is checked here on CPU tensors through ``_trace``;
``forward`` itself refuses CPU tensors like the reference's hard-coded ``.cuda()`` does.  The GPU run of the same
cases, and the fused SIREN evaluator, are in tests/test_gpu_rays.py."""
import numpy as np
import pytest
import torch

from isopoints_b200.ray_tracing import RayTracing, intersection_with_unit_sphere
from tests.helpers import SphereSDF, TinySiren, make_camera_rays

CASES = [("siren", lambda: TinySiren(seed=3), 10), ("sphere", lambda: SphereSDF(radius=0.5), 10),
         ("siren3", lambda: TinySiren(seed=5), 3)]


def _sdf_of(net):
    def sdf(x):
        with torch.no_grad():
            return net(x).sdf.squeeze(-1)
    return sdf


@pytest.mark.parametrize("name,make,iters", CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_matches_reference_golden(golden, name, make, iters, mode):
    g = golden("ray_tracing")
    cam, dirs, om = (torch.as_tensor(g[name + k]) for k in ("_cam", "_dirs", "_object_mask"))
    tracer = RayTracing(1.0, sphere_tracing_iters=iters, n_steps=int(g["n_steps"]), n_secant_steps=8)
    tracer.train(mode == "train")
    key = "%s_%s_" % (name, mode)
    steps = torch.as_tensor(g[key + "steps"]) if mode == "train" else None
    pts, mask, dists = tracer._trace(_sdf_of(make()), cam, om, dirs, steps)
    assert pts.shape == (3000, 3) and mask.shape == (3000,) and mask.dtype == torch.bool and dists.shape == (3000,)
    # same float32 op sequence as the reference up to the batch shape the SDF sees: rounding-level agreement
    assert np.array_equal(mask.numpy(), g[key + "mask"])
    np.testing.assert_allclose(pts.numpy(), g[key + "points"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(dists.numpy(), g[key + "dists"], rtol=0, atol=2e-6)


def test_golden_exercises_every_branch(golden):
    """The fixture is only worth something if the sampler, the secant and the training-only branches ran."""
    g = golden("ray_tracing")
    cam, dirs = torch.as_tensor(g["siren_cam"]), torch.as_tensor(g["siren_dirs"])
    _, _, hit = intersection_with_unit_sphere(cam, dirs)
    assert 0.05 < 1 - hit.float().mean().item() < 0.6            # rays that miss the bounding sphere
    tracer = RayTracing(1.0, sphere_tracing_iters=10, n_steps=64)
    calls = []
    sdf = _sdf_of(TinySiren(seed=3))

    def counting(x):
        calls.append(x.shape[0])
        return sdf(x)
    tracer.train(True)
    tracer._trace(counting, cam, torch.as_tensor(g["siren_object_mask"]), dirs)
    assert any(c % 64 == 0 and c >= 64 * 50 for c in calls)      # sampler / minimal-SDF batches (n_steps per ray)
    assert (g["siren_train_points"] != g["siren_eval_points"]).any()   # the training-only branches changed rays
    assert np.array_equal(g["siren_train_mask"], g["siren_eval_mask"])


def test_forward_refuses_cpu_tensors_and_keeps_the_reference_defaults():
    t = RayTracing()
    assert (t.object_bounding_sphere, t.sdf_threshold, t.line_search_step, t.line_step_iters,
            t.sphere_tracing_iters, t.n_steps, t.n_secant_steps) == (1.0, 5.0e-5, 0.5, 1, 10, 100, 8)
    cam, dirs = make_camera_rays(1, 8, seed=0)
    with pytest.raises(TypeError):
        t(lambda x: x.norm(dim=-1) - 0.5, cam, torch.ones(8, dtype=torch.bool), dirs)


def test_sphere_has_the_closed_form_answer():
    """Unit-slope SDF of a sphere of radius 0.5: the traced distance is the analytic ray/sphere intersection."""
    cam, dirs = make_camera_rays(2, 400, seed=5, target_radius=0.45)      # every ray hits the inner sphere
    t = RayTracing(1.0, sphere_tracing_iters=30).eval()
    pts, mask, dists = t._trace(lambda x: x.norm(dim=-1) - 0.5, cam, torch.ones(800, dtype=torch.bool), dirs)
    o = cam[:, None, :].expand(2, 400, 3).reshape(-1, 3).double()
    d = dirs.reshape(-1, 3).double()
    b = (o * d).sum(-1)
    want = -b - torch.sqrt(b * b - (o * o).sum(-1) + 0.25)
    assert bool(mask.all())
    np.testing.assert_allclose(dists.numpy(), want.numpy(), rtol=0, atol=2e-4)
    np.testing.assert_allclose(pts.norm(dim=-1).numpy(), 0.5, atol=2e-4)
