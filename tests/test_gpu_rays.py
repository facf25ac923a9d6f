"""GPU run of the RayTracing mirror (isopoints_b200/ray_tracing.py; reference levelset_sampling.py:810-1167):
the reference-generated golden cases on cuda tensors, and the reference's Siren decoder evaluated by the
forward-only fused kernel (``siren.sdf_fn`` -> ``isob200_siren_sdf``) against the same weights through PyTorch.

Two float32 SDF implementations differ by ~1e-7, which moves a marching ray by as much and can flip a ray that
sits on a threshold or change which of the ``n_steps`` samples is the first negative one; masks are therefore
compared as an agreement rate, positions (tolerance 1e-4, the north_star bar) on all but a small fraction of the
rays both sides call hits."""
import numpy as np
import pytest
import torch

from isopoints_b200 import siren
from isopoints_b200.ray_tracing import RayTracing
from tests.helpers import Siren, SirenSDF, SphereSDF, TinySiren, make_camera_rays

pytestmark = pytest.mark.gpu
DEV = "cuda"
CASES = [("siren", lambda: TinySiren(seed=3), 10), ("sphere", lambda: SphereSDF(radius=0.5), 10),
         ("siren3", lambda: TinySiren(seed=5), 3)]


def _agree(got, want, min_agree=0.995, frac_close=0.995, atol=1e-4):
    pts, mask, dist = (np.asarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t) for t in got)
    wpts, wmask, wdist = (np.asarray(t.detach().cpu().numpy() if torch.is_tensor(t) else t) for t in want)
    same = mask == wmask
    assert same.mean() >= min_agree, same.mean()
    both = mask & wmask
    assert both.mean() > 0.1
    close = (np.abs(pts[both] - wpts[both]).max(axis=1) < atol) & (np.abs(dist[both] - wdist[both]) < atol)
    assert close.mean() >= frac_close, close.mean()
    return same.mean(), close.mean()


@pytest.mark.parametrize("name,make,iters", CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_matches_reference_golden(golden, name, make, iters, mode):
    g = golden("ray_tracing")
    net = make().to(DEV)
    cam, dirs, om = (torch.as_tensor(g[name + k], device=DEV) for k in ("_cam", "_dirs", "_object_mask"))

    def sdf(x):
        with torch.no_grad():
            return net(x).sdf.squeeze(-1)
    tracer = RayTracing(1.0, sphere_tracing_iters=iters, n_steps=int(g["n_steps"]), n_secant_steps=8)
    tracer.train(mode == "train")
    key = "%s_%s_" % (name, mode)
    steps = torch.as_tensor(g[key + "steps"], device=DEV) if mode == "train" else None
    got = tracer(sdf, cam, om, dirs, minimal_sdf_steps=steps)
    assert got[0].shape == (3000, 3) and got[1].dtype == torch.bool and got[2].shape == (3000,)
    # training mode moves the network hits outside the ground-truth mask to the arg-min of 64 samples along the
    # ray (:904-913): near-ties between two samples on either side of the minimum pick differently at 1e-7
    _agree(got, (g[key + "points"], g[key + "mask"], g[key + "dists"]), frac_close=0.995 if mode == "eval" else 0.97)
    if mode == "train":
        # rays outside the network's object keep the minimal-SDF sample: compare those too (sample choice can
        # differ on near-ties, so a fraction)
        miss = ~g[key + "mask"] & ~got[1].cpu().numpy()
        d = np.abs(got[0].cpu().numpy()[miss] - g[key + "points"][miss]).max(axis=1)
        assert (d < 1e-4).mean() > 0.97, (d < 1e-4).mean()


def _zero_mean(net_cls, layers=3, seed=6):
    ref = Siren(256, layers, 30.0, seed=seed)
    net = net_cls(256, layers, 30.0, seed=seed)
    with torch.no_grad():
        x = (torch.rand(4000, 3, generator=torch.Generator().manual_seed(0)) - 0.5) * 2
        shift = ref(x).sdf.mean()
        (net.lin[-1] if net_cls is SirenSDF else net.net[-1]).bias -= shift
    return net


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_fused_siren_evaluator_vs_pytorch_evaluation(mode):
    """Same RayTracing run, SDF of the reference's Siren decoder through isob200_siren_sdf vs through PyTorch."""
    fused_net, torch_net = _zero_mean(Siren).to(DEV), _zero_mean(SirenSDF).to(DEV)
    cam, dirs = make_camera_rays(2, 6000, seed=9)
    cam, dirs = cam.to(DEV), dirs.to(DEV)
    om = torch.rand(12000, device=DEV) < 0.8
    steps = torch.rand(100, generator=torch.Generator().manual_seed(1)).to(DEV)
    tracer = RayTracing()
    tracer.train(mode == "train")
    calls0 = siren.STATS["calls"]
    f = siren.sdf_fn(fused_net)
    got = tracer(f, cam, om, dirs, minimal_sdf_steps=steps)
    assert siren.STATS["calls"] > calls0                    # the fused kernel, not the module, did the evaluations

    def sdf(x):
        with torch.no_grad():
            return torch_net(x).sdf.squeeze(-1)
    want = tracer(sdf, cam, om, dirs, minimal_sdf_steps=steps)
    # a random high-frequency SIREN is far from a distance function: many near-tie sample choices
    _agree(got, want, min_agree=0.99, frac_close=0.97)
    hit = got[1]
    assert 0.2 < hit.float().mean().item() < 1.0


def test_no_ray_hits_the_bounding_sphere_and_single_ray():
    net = TinySiren(seed=3).to(DEV)

    def sdf(x):
        with torch.no_grad():
            return net(x).sdf.squeeze(-1)
    cam = torch.tensor([[0.0, 0.0, 3.0]], device=DEV)
    away = torch.tensor([[[0.0, 1.0, 0.0], [1.0, 0.0, 0.0]]], device=DEV)       # parallel to the image plane
    t = RayTracing().eval()
    pts, mask, dist = t(sdf, cam, torch.ones(2, dtype=torch.bool, device=DEV), away)
    assert not bool(mask.any()) and pts.shape == (2, 3) and bool((dist == 0).all())
    t.train(True)
    pts, mask, dist = t(sdf, cam, torch.ones(2, dtype=torch.bool, device=DEV), away)
    assert not bool(mask.any())
    np.testing.assert_allclose(pts.cpu().numpy(), [[0, 0, 3.0], [0, 0, 3.0]], atol=1e-6)   # closest point to the origin
    one = torch.tensor([[[0.0, 0.0, -1.0]]], device=DEV)
    pts, mask, dist = t.eval()(sdf, cam, torch.ones(1, dtype=torch.bool, device=DEV), one)
    assert bool(mask.all()) and abs(sdf(pts).item()) < 1e-4
