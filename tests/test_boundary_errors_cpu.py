"""Every public operator of the package refuses CPU tensors (TypeError, the reference's own message for its
CUDA-only ops, external/FRNN/frnn/frnn.py:255-256) instead of computing anything on the host: there is no CPU or
PyTorch fallback behind the C ABI.  Argument validation that the reference performs before touching the device
(dimension check, frnn.py:253-254) keeps its exception type."""
import pytest
import torch

from isopoints_b200 import ewa, frnn, levelset_sampling as ls, offsurface, point_processing as pp, splat
from isopoints_b200.ray_tracing import RayTracing
from isopoints_b200.structures import Pointclouds
from tests.helpers import SphereSDF

X = torch.rand(1, 64, 3)
I32 = torch.zeros(1, 2, 2, 1, dtype=torch.int32)

CPU_CALLS = {
    "frnn_grid_points": lambda: frnn.frnn_grid_points(X, X, K=2, r=0.1),
    "frnn_gather": lambda: frnn.frnn_gather(X, torch.zeros(1, 64, 2, dtype=torch.long)),
    "prefix_sum_cuda": lambda: frnn.prefix_sum_cuda(torch.zeros(4, dtype=torch.int32), 4, torch.zeros(4, dtype=torch.int32)),
    "splat_points": lambda: splat._C.splat_points(torch.rand(4, 3), torch.rand(4, 3), torch.ones(4), torch.rand(4, 2),
                                                  torch.zeros(1, dtype=torch.long), torch.tensor([4]), 0.05, 32, 2, 0, 0),
    "rasterize_elliptical_points": lambda: splat.rasterize_elliptical_points(
        Pointclouds([torch.rand(4, 3)]), torch.rand(4, 3), torch.ones(4), torch.rand(4, 2)),
    "blend_rgba": lambda: splat.blend_rgba(I32, torch.zeros(1, 2, 2, 1), torch.zeros(1, 2, 2), None, torch.rand(4, 3)),
    "visibility_mask": lambda: splat.visibility_mask(I32, 4),
    "wlop": lambda: pp.wlop(Pointclouds([torch.rand(64, 3)])),
    "upsample": lambda: pp.upsample(X, 80),
    "resample_uniformly": lambda: pp.resample_uniformly(Pointclouds([torch.rand(64, 3)])),
    "UniformProjection.project_points": lambda: ls.UniformProjection().project_points(X, SphereSDF()),
    "SphereTracing.project_points": lambda: ls.SphereTracing().project_points(X[:, :8], X[:, 8:16], SphereSDF()),
    "RayTracing.forward": lambda: RayTracing()(lambda p: p.norm(dim=-1) - 0.5, torch.rand(1, 3) + 2,
                                               torch.ones(8, dtype=torch.bool), X[:, :8]),
    "compute_isotropic_vrk_h": lambda: ewa.compute_isotropic_vrk_h(X, torch.tensor([64]), 0.2),
    "closest_point_to_rays": lambda: offsurface.closest_point_to_rays(torch.zeros(3), X[0, :4], X[0]),
}


@pytest.mark.parametrize("name", sorted(CPU_CALLS))
def test_cpu_tensors_are_refused(name):
    with pytest.raises(TypeError):
        CPU_CALLS[name]()


def test_dimension_check_precedes_the_device_check():
    with pytest.raises(ValueError):
        frnn.frnn_grid_points(torch.rand(1, 8, 4), torch.rand(1, 8, 4), K=2, r=0.1)
