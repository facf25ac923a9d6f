"""Generate tests/golden/*.npz by running the REFERENCE's own code (authoring container only).

    python tests/golden/make_golden.py

* projection / resample vectors: the reference's Python (DSS/models/levelset_sampling.py loaded
  from /root/reference through oracle/ref_python.py's asserted 4-entry torch-2.x patch list) on
  CPU tensors; its `frnn` dependency (CUDA-only) is replaced by a stand-in built on the
  reference's own CPU brute force `frnn._C.frnn_bf_cpu` (oracle/_ref/ref_frnn_C.so).
* FRNN vectors: `frnn._C.frnn_bf_cpu` (bruteforce_cpu.cpp:4-58).
* splat vectors: the reference's `DSS._C._splat_points_naive` CPU twin
  (rasterize_points_cpu.cpp:27-144) -- NB its bbox reject uses && where the CUDA kernel uses ||
  (SURVEY 7.3), which is invisible when radii bound the cutoff ellipse, as they do here.
* EWA per-point parameters (`--only ewa` regenerates just this one): the reference's
  `SurfaceSplatting._get_per_point_info` / `_filter_points_with_invalid_depth` /
  `_filter_backface_points` (DSS/core/rasterizer.py) on CPU tensors, with duck-typed point clouds and
  cameras (pytorch3d is not installed: the camera object hands out the 4x4 matrices directly and its
  world-to-view transform restates Transform3d.transform_points / transform_normals), the K = 7
  neighbour query through the reference's own `frnn_bf_cpu`.
* sphere tracing (`--only trace`): the reference's `SphereTracing.project_points`
  (levelset_sampling.py:679-808) on CPU tensors.
* IDR ray tracing (`--only rays`): the reference's `RayTracing.forward` (levelset_sampling.py:810-1167) on CPU
  tensors (its hard-coded `.cuda()` calls made no-ops for the duration of the run; the `uniform_` draw of
  `minimal_sdf_points` recorded so the parity test can inject the same positions), eval and training mode.
* in-surface / off-surface sampler (`--only offsurface`): the reference's
  `Model.sample_offsurface_using_isopoints` (DSS/models/combined_modeling.py:237-388) called unbound on a
  duck-typed model (decoder = TinySiren) and cameras (tests/helpers.PinholeCameras), with its two
  `get_visible_points` calls answered by preset front / back point sets (the visibility pass has its own tests)
  and the `torch.rand_like` draw recorded; plus `intersection_with_unit_cube` and `get_tensor_values` alone.
The fixtures are small (< 1 MB total) and committed; the GPU box has no /root/reference.
"""
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_native, ref_python  # noqa: E402
from tests.helpers import (SphereSDF, TinySiren, make_splat_inputs, make_cameras, make_surface_points,  # noqa: E402
                           make_rays, make_camera_rays, offsurface_inputs)


class _CpuFrnn(types.SimpleNamespace):
    """frnn.frnn_grid_points / frnn_gather on CPU tensors via the reference's frnn_bf_cpu."""

    @staticmethod
    def frnn_grid_points(points1, points2, lengths1=None, lengths2=None, K=-1, r=-1, grid=None,
                         return_nn=False, return_sorted=True, radius_cell_ratio=2.0):
        N = points1.shape[0]
        if lengths1 is None:
            lengths1 = torch.full((N,), points1.shape[1], dtype=torch.long)
        if lengths2 is None:
            lengths2 = torch.full((N,), points2.shape[1], dtype=torch.long)
        rr = float(r.reshape(-1)[0]) if torch.is_tensor(r) else float(r)
        idxs, dists = ref_native.frnn_bf_cpu(points1.contiguous(), points2.contiguous(), lengths1, lengths2, K, rr)
        nn = _CpuFrnn.frnn_gather(points2, idxs, lengths2) if return_nn else None
        return dists, idxs, nn, None

    @staticmethod
    def frnn_gather(x, idxs, lengths=None):
        m = idxs < 0
        j = idxs.clamp_min(0)
        out = torch.stack([x[n][j[n]] for n in range(x.shape[0])], 0)
        out[m] = 0
        return out


class _Transform:
    """What the reference uses of pytorch3d.transforms.Transform3d (row-vector convention)."""

    def __init__(self, m):
        self.m = m

    def get_matrix(self):
        return self.m

    def transform_points(self, points, eps=None):
        hom = torch.cat([points, torch.ones_like(points[..., :1])], -1)
        out = hom @ self.m
        return out[..., :3] / out[..., 3:]

    def transform_normals(self, normals):
        return normals @ torch.inverse(self.m)[:, :3, :3].transpose(1, 2)


class _Cameras:
    def __init__(self, w2v, proj, znear, zfar):
        self.w2v, self.proj, self.znear, self.zfar = w2v, proj, znear, zfar
        self.R = w2v[:, :3, :3]

    def __len__(self):
        return self.proj.shape[0]

    def get_full_projection_transform(self):
        return _Transform(self.proj)

    def get_world_to_view_transform(self):
        return _Transform(self.w2v)


class _Clouds:
    """The PointClouds3D accessors the per-point code touches (packed = concatenation, padded = zeros)."""

    def __init__(self, points=None, normals=None, features=None):
        self.pl, self.nl = list(points), (list(normals) if normals is not None else None)
        self.device = self.pl[0].device

    def __len__(self):
        return len(self.pl)

    def isempty(self):
        return sum(len(p) for p in self.pl) == 0

    def num_points_per_cloud(self):
        return torch.tensor([len(p) for p in self.pl], dtype=torch.int64)

    def cloud_to_packed_first_idx(self):
        n = self.num_points_per_cloud()
        return torch.cat([n.new_zeros(1), n.cumsum(0)[:-1]])

    def packed_to_cloud_idx(self):
        return torch.repeat_interleave(torch.arange(len(self.pl)), self.num_points_per_cloud())

    def points_packed(self):
        return torch.cat(self.pl, 0)

    def normals_packed(self):
        return torch.cat(self.nl, 0)

    def _padded(self, lst):
        out = lst[0].new_zeros(len(lst), max(len(p) for p in lst), 3)
        for b, p in enumerate(lst):
            out[b, :len(p)] = p
        return out

    def points_padded(self):
        return self._padded(self.pl)

    def normals_padded(self):
        return self._padded(self.nl)

    def features_padded(self):
        return None


def ewa_golden():
    R = ref_python.load_rasterizer()
    R.frnn = _CpuFrnn
    SS = R.SurfaceSplatting
    S, sigma, cutoff, znear, zfar, radius = 256, 1.0, 1.0, 2.0, 100.0, 0.2
    num = [900, 5, 700]                       # the 5-point cloud takes the `< K points` branch (:376)
    pts, nrm, first, numt = make_surface_points(num, seed=11)
    nrm[7] = 0.0                              # a degenerate normal: S_k = 0
    w2v, proj, nmat = make_cameras(3, seed=12, znear=znear, zfar=zfar)
    ras = SS.__new__(SS)
    ras.frnn_radius = radius
    ras._Vrk_h = None
    ras.raster_settings = types.SimpleNamespace(cutoff_threshold=cutoff, Vrk_invariant=False, Vrk_isotropic=True,
                                                image_size=S, antialiasing_sigma=sigma, backface_culling=True)
    ras.cameras = _Cameras(w2v, proj, znear, zfar)
    clouds = _Clouds([pts[f:f + n] for f, n in zip(first.tolist(), num)],
                     [nrm[f:f + n] for f, n in zip(first.tolist(), num)])
    torch.manual_seed(5)
    info = SS._get_per_point_info(ras, clouds)
    sq_dists = _CpuFrnn.frnn_grid_points(clouds.points_padded(), clouds.points_padded(), numt, numt, K=7, r=radius)[0]
    # the renderable filters rebuild clouds from PADDED tensors (:186-195), so a ragged batch would turn
    # surviving zero padding into points; the reference only ever feeds them equal-sized clouds
    numf = [500, 500, 500]
    fpts, fnrm, ffirst, fnumt = make_surface_points(numf, seed=13)
    fclouds = _Clouds([fpts[f:f + n] for f, n in zip(ffirst.tolist(), numf)],
                      [fnrm[f:f + n] for f, n in zip(ffirst.tolist(), numf)])
    _, mask_depth = SS._filter_points_with_invalid_depth(ras, fclouds)
    _, mask_renderable = SS.filter_renderable(ras, fclouds)
    np.savez_compressed(
        os.path.join(HERE, "ewa_point_info.npz"), image_size=S, antialiasing_sigma=sigma, cutoff=cutoff,
        znear=znear, zfar=zfar, frnn_radius=radius, points=pts.numpy(), normals=nrm.numpy(),
        first_idx=first.numpy(), num_points=numt.numpy(), w2v=w2v.numpy(), proj=proj.numpy(), nmat=nmat.numpy(),
        sq_dists=sq_dists.numpy(), vrk_h=ras._Vrk_h.view(-1).numpy(), radii=info["radii"].numpy(),
        ellipse=info["ellipse_params"].numpy(), cutoff_threshold=info["cutoff_threshold"].numpy(),
        scaler=info["scaler"].numpy(), filter_points=fpts.numpy(), filter_normals=fnrm.numpy(),
        filter_first_idx=ffirst.numpy(), filter_num_points=fnumt.numpy(), mask_depth=mask_depth.numpy(),
        mask_renderable=mask_renderable.numpy())
    print("ewa: radii px", float(info["radii"].mean()) * S / 2, "depth-valid", float(mask_depth.float().mean()),
          "renderable", float(mask_renderable.float().mean()))


def ewa_global_golden():
    """The view-invariant V_k^r setting (`Vrk_invariant=True`: `_compute_global_Vrk`, rasterizer.py:292-343 -- one
    clamped mean h per cloud) through the reference's `_get_per_point_info`, on the inputs of `ewa_golden`
    (ragged batch: the mean runs over the PADDED rows, whose distances are the -1 padding) and on an
    equal-sized batch."""
    R = ref_python.load_rasterizer()
    R.frnn = _CpuFrnn
    SS = R.SurfaceSplatting
    S, sigma, cutoff, znear, zfar, radius = 256, 1.0, 1.0, 2.0, 100.0, 0.2
    out = dict(image_size=S, antialiasing_sigma=sigma, cutoff=cutoff, frnn_radius=radius, znear=znear, zfar=zfar)
    for tag, num, seed in (("ragged", [900, 5, 700], 11), ("equal", [12000, 12000], 21)):
        pts, nrm, first, numt = make_surface_points(num, seed=seed)
        w2v, proj, nmat = make_cameras(len(num), seed=seed + 1, znear=znear, zfar=zfar)
        ras = SS.__new__(SS)
        ras.frnn_radius = radius
        ras._Vrk_h = None
        ras.raster_settings = types.SimpleNamespace(cutoff_threshold=cutoff, Vrk_invariant=True, Vrk_isotropic=True,
                                                    image_size=S, antialiasing_sigma=sigma, backface_culling=True)
        ras.cameras = _Cameras(w2v, proj, znear, zfar)
        clouds = _Clouds([pts[f:f + n] for f, n in zip(first.tolist(), num)],
                         [nrm[f:f + n] for f, n in zip(first.tolist(), num)])
        torch.manual_seed(5)
        info = SS._get_per_point_info(ras, clouds)
        sq = _CpuFrnn.frnn_grid_points(clouds.points_padded(), clouds.points_padded(), numt, numt, K=7, r=radius)[0]
        sq = sq[:, :, 1:].clone()
        sq[numt < 7] = 1e-3
        h = (0.5 * sq.max(dim=-1, keepdim=True)[0]).mean(dim=1, keepdim=True).clamp(5e-5, 1e-3).view(-1)
        # inputs are regenerated from (num, seed) by the test (tests/helpers.py); outputs every `step`-th row
        step = 8 if sum(num) > 5000 else 1
        out.update({tag + "_num": np.asarray(num), tag + "_seed": seed, tag + "_step": step, tag + "_h_cloud": h.numpy(),
                    tag + "_radii": info["radii"].numpy()[::step], tag + "_ellipse": info["ellipse_params"].numpy()[::step],
                    tag + "_scaler": info["scaler"].numpy()[::step]})
        print("ewa global %s: h per cloud %s, radii px %.2f" % (tag, h.tolist(), float(info["radii"].mean()) * S / 2))
    np.savez_compressed(os.path.join(HERE, "ewa_point_info_global.npz"), **out)


def trace_golden(LS):
    """SphereTracing.project_points (levelset_sampling.py:679-808) on CPU tensors: a blob SDF and the unit
    sphere; rays that hit, graze and leave the bounding sphere."""
    def same_shape(*args, device=None):   # pytorch3d.renderer.utils.convert_to_tensors_and_broadcast for
        assert all(torch.is_tensor(a) and a.shape == args[0].shape for a in args)   # tensors of one shape
        return list(args)
    LS.convert_to_tensors_and_broadcast = same_shape
    out = {}
    for name, net, n, tol in (("siren", TinySiren(seed=3), 3000, 5e-5), ("sphere", SphereSDF(radius=0.5), 2000, 5e-5)):
        ray0, dirs = make_rays(n, seed=21, target_radius=0.7)
        tracer = LS.SphereTracing(proj_max_iters=25, proj_tolerance=tol, alpha=1.0, radius=1.0, padding=0.1)
        res = tracer.project_points(ray0.view(2, -1, 3).clone(), dirs.view(2, -1, 3).clone(), net)
        out[name + "_ray0"], out[name + "_dirs"] = ray0.numpy(), dirs.numpy()
        out[name + "_points"] = res["levelset_points"].reshape(-1, 3).numpy()
        out[name + "_eval"] = res["network_eval_on_levelset_points"].reshape(-1).numpy()
        out[name + "_mask"] = res["mask"].reshape(-1).numpy()
        moved = (res["levelset_points"].reshape(-1, 3) - ray0).norm(dim=-1)
        print("trace %s: hit %.3f, moved mean %.3f, shapes %s %s" % (
            name, float(res["mask"].float().mean()), float(moved.mean()), tuple(res["levelset_points"].shape),
            tuple(res["mask"].shape)))
    np.savez_compressed(os.path.join(HERE, "sphere_trace.npz"), proj_max_iters=25, **out)


def rays_golden(LS):
    """RayTracing.forward (levelset_sampling.py:831-918) on CPU tensors.  Cases: the TinySiren blob (half-slope
    SDF, so the march leaves rays for the sampler + secant) and a sphere (march converges), in eval and in
    training mode (which adds the minimal-SDF search for the rays that disagree with the ground-truth mask)."""
    real_cuda, real_uniform = torch.Tensor.cuda, torch.Tensor.uniform_
    drawn = []

    def uniform_rec(self, *a, **k):
        real_uniform(self, *a, **k)
        drawn.append(self.clone())
        return self
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.uniform_ = uniform_rec
    out = {}
    try:
        for name, net, iters in (("siren", TinySiren(seed=3), 10), ("sphere", SphereSDF(radius=0.5), 10),
                                 ("siren3", TinySiren(seed=5), 3)):
            cam, dirs = make_camera_rays(2, 1500, seed=31)
            with torch.no_grad():
                inside = net(cam[:, None, :] + 2.5 * dirs).sdf  # crude ground-truth mask: correlated, not equal
            g = torch.Generator().manual_seed(7)
            object_mask = ((inside.reshape(-1) < 0.35) ^ (torch.rand(3000, generator=g) < 0.15))

            def sdf(x):
                with torch.no_grad():
                    return net(x).sdf.squeeze(-1)
            out[name + "_cam"], out[name + "_dirs"], out[name + "_object_mask"] = cam.numpy(), dirs.numpy(), object_mask.numpy()
            for mode in ("eval", "train"):
                tracer = LS.RayTracing(object_bounding_sphere=1.0, sphere_tracing_iters=iters, n_steps=64, n_secant_steps=8)
                tracer.train(mode == "train")
                del drawn[:]
                pts, net_mask, dists = tracer(sdf, cam.clone(), object_mask.clone(), dirs.clone())
                key = "%s_%s_" % (name, mode)
                out[key + "points"], out[key + "mask"], out[key + "dists"] = pts.numpy(), net_mask.numpy(), dists.numpy()
                if drawn:
                    out[key + "steps"] = drawn[-1].numpy()
                print("rays %s/%s: net mask %.3f, object mask %.3f, uniform draws %d" % (
                    name, mode, float(net_mask.float().mean()), float(object_mask.float().mean()), len(drawn)))
    finally:
        torch.Tensor.cuda, torch.Tensor.uniform_ = real_cuda, real_uniform
    np.savez_compressed(os.path.join(HERE, "ray_tracing.npz"), n_steps=64, **out)


def offsurface_golden(ref):
    import importlib
    from isopoints_b200.structures import Pointclouds
    CM = importlib.import_module("DSS.models.combined_modeling")
    U = importlib.import_module("DSS.utils")
    cams, pixels, mask_img, frontal, occluded, iso_pcl = offsurface_inputs()
    answers = [Pointclouds(frontal), Pointclouds(occluded)]
    calls = []

    def visible(points, cameras, depth_merge_threshold=0.05, return_mask=False):
        calls.append((cameras.R.clone(), cameras.T.clone(), depth_merge_threshold))
        return answers[len(calls) - 1]
    drawn = []
    real_rand_like = torch.rand_like

    def rand_like(x, *a, **k):
        drawn.append(real_rand_like(x, *a, **k))
        return drawn[-1]
    net = TinySiren(seed=3)
    model = types.SimpleNamespace(
        _points=None, decoder=net, max_points_per_pass=10000, object_bounding_sphere=1.0,
        renderer=types.SimpleNamespace(rasterizer=types.SimpleNamespace(
            raster_settings=types.SimpleNamespace(depth_merging_threshold=0.05))))
    CM.get_visible_points = visible
    CM.torch.rand_like = rand_like
    torch.manual_seed(11)
    try:
        p_off, p_ins, n_off, n_ins = CM.Model.sample_offsurface_using_isopoints(
            model, pixels.clone(), mask_img.clone(), cams, n_points_per_ray=32,
            max_insurface_per_batch=[150, 400], iso_pcl=Pointclouds(iso_pcl))
    finally:
        CM.torch.rand_like = real_rand_like
    assert len(calls) == 2 and len(drawn) == 1
    out = dict(p_off=p_off.numpy(), p_ins=p_ins.numpy(), n_off=n_off.numpy(), n_ins=n_ins.numpy(), rand=drawn[0].numpy(),
               back_R=calls[1][0].numpy(), back_T=calls[1][1].numpy(), n_points_per_ray=32,
               max_insurface=np.array([150, 400]))
    # the two helpers alone
    cam_pos = cams.get_camera_center()
    world = cams.unproject_points(torch.cat([-pixels, torch.ones_like(pixels[..., :1])], -1), scaled_depth_input=False)
    rays = torch.nn.functional.normalize(world - cam_pos[:, None, :], dim=-1)
    c0, c1, cm = U.intersection_with_unit_cube(cam_pos.view(-1, 1, 3), rays, side_length=2.0)
    vals = U.get_tensor_values(mask_img, pixels.clamp(-1, 1), squeeze_channel_dim=True)
    out.update(cube_rays=rays.numpy(), cube0=c0.numpy(), cube1=c1.numpy(), cube_mask=cm.numpy(), mask_values=vals.numpy())
    np.savez_compressed(os.path.join(HERE, "offsurface.npz"), **out)
    print("offsurface: off %s, in %s of caps [150, 400], cube hits %.3f" % (n_off.tolist(), n_ins.tolist(), float(cm.float().mean())))


def pinned_golden():
    """`--only pinned`: the C2 SDF as SURVEY 8d pins it -- the reference's own ``DSS.models.common.Siren(dim=3,
    c_dim=0, hidden_size=256, n_layers=7, first_omega_0=30, hidden_omega_0=30, outermost_linear=True)`` built right
    after ``torch.manual_seed(0)`` -- a fingerprint of its state_dict (so that tests/helpers.pinned_siren can be
    checked where /root/reference is absent), its value / autograd gradient on 512 points in float64 from the
    reference class itself, and the reference's ``UniformProjection.project_points(skip_upsampling=True)`` on the
    first 4096 points of the C2 cloud (CPU fp32 autograd SDF, frnn = the reference's frnn_bf_cpu)."""
    import importlib
    ref = ref_python.load(frnn_module=_CpuFrnn)
    common = importlib.import_module("DSS.models.common")
    torch.manual_seed(0)
    net = common.Siren(dim=3, c_dim=0, hidden_size=256, n_layers=7, first_omega_0=30, hidden_omega_0=30,
                       outermost_linear=True)
    sd = net.state_dict()
    keys = sorted(sd.keys())
    fp = np.stack([np.array([float(sd[k].double().sum()), float(sd[k].double().abs().sum()),
                             float(sd[k].reshape(-1)[0]), float(sd[k].reshape(-1)[-1])]) for k in keys])
    g = torch.Generator().manual_seed(1000)              # bench.py's C2 cloud of rank 0
    x = ((torch.rand(1, 200000, 3, generator=g) - 0.5) * 2)[:, :4096].contiguous()
    net64 = common.Siren(dim=3, c_dim=0, hidden_size=256, n_layers=7, first_omega_0=30, hidden_omega_0=30,
                         outermost_linear=True).double()
    net64.load_state_dict({k: v.double() for k, v in sd.items()})
    xe = x[0, :512].double().requires_grad_(True)
    s64 = net64(xe).sdf
    g64, = torch.autograd.grad(s64, xe, torch.ones_like(s64))
    proj = ref.levelset_sampling.UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    p0 = proj._project_points(net, x.clone(), torch.tensor([4096]))
    out = proj.project_points(x.clone(), net, skip_upsampling=True)
    np.savez_compressed(os.path.join(HERE, "pinned_siren.npz"), keys=np.array(keys), fingerprint=fp, x=x.numpy(),
                        sdf64=s64.detach().numpy().reshape(-1), grad64=g64.numpy(),
                        proj_points=p0.points.numpy(), proj_normals=p0.normals.numpy(), proj_mask=p0.mask.numpy(),
                        points=out["levelset_points"].numpy(), normals=out["levelset_normals"].numpy(),
                        mask=out["mask"].numpy())
    print("pinned: projection converged", float(p0.mask.float().mean()), "after resample", tuple(out["mask"].shape),
          float(out["mask"].float().mean()))


def insert_golden():
    """`--only insert`: the reference's ``UniformProjection.insert`` (levelset_sampling.py:172-233) and the
    ``ref_pcl`` branch of ``project_points`` (:411-424) on CPU tensors: a sphere SDF, a reference cloud whose
    per-point feature (the saliency metric) peaks around two spots."""
    from isopoints_b200.structures import Pointclouds
    ref = ref_python.load(frnn_module=_CpuFrnn)
    LS = ref.levelset_sampling
    torch.manual_seed(11)
    x = (torch.rand(1, 3000, 3) - 0.5) * 1.5
    rp = torch.nn.functional.normalize(torch.randn(2500, 3), dim=-1)
    spots = torch.nn.functional.normalize(torch.tensor([[1.0, 0.2, 0.1], [-0.3, 0.9, -0.2]]), dim=-1)
    metric = torch.exp(-((rp[:, None, :] - spots[None]) ** 2).sum(-1) / 0.02).sum(-1, keepdim=True) \
        + 0.01 * torch.rand(2500, 1)
    ref_pcl = Pointclouds([rp], features=[metric])
    proj = LS.UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    out = proj.project_points(x.clone(), SphereSDF(), ref_pcl=ref_pcl)
    # insert() alone on the projected + resampled cloud (what project_points hands it)
    base = proj.project_points(x.clone(), SphereSDF(), skip_upsampling=True)
    bp = base["levelset_points"][:, base["mask"][0]]
    num = torch.tensor([bp.shape[1]])
    pts_all, num_all, child, child_num = proj.insert(ref_pcl, bp.clone(), num)
    # second case: threshold branch (few salient points -> the `metrics > threshold` selection is kept)
    metric2 = 0.01 * torch.rand(2500, 1)
    metric2[:30] = 1.0
    ref_pcl2 = Pointclouds([rp], features=[metric2])
    _, _, child2, child_num2 = proj.insert(ref_pcl2, bp.clone(), num)
    np.savez_compressed(os.path.join(HERE, "insert.npz"), x=x.numpy(), ref_points=rp.numpy(), metric=metric.numpy(),
                        metric2=metric2.numpy(), points=out["levelset_points"].numpy(),
                        normals=out["levelset_normals"].numpy(), mask=out["mask"].numpy(), base=bp.numpy(),
                        child=child.numpy(), child_num=child_num.numpy(), child2=child2.numpy(),
                        child_num2=child_num2.numpy())
    print("insert:", tuple(out["levelset_points"].shape), "children", int(child_num[0]), int(child_num2[0]),
          "valid", float(out["mask"].float().mean()))


def _fps_sequential(x, batch, ratio, random_start=False):
    """torch_cluster.fps stand-in [third party, restated]: per cloud, start at its first point, repeatedly take
    the point farthest (squared distance, first arg-max) from the selected set; ceil(ratio * n) indices per
    cloud, in selection order, concatenated."""
    out = []
    for b in range(int(batch.max()) + 1):
        ids = (batch == b).nonzero().reshape(-1)
        pts = x[ids]
        m = int(math.ceil(ratio * len(ids)))
        sel = [0]
        md = ((pts - pts[0]) ** 2).sum(-1)
        for _ in range(m - 1):
            j = int(torch.argmax(md))
            sel.append(j)
            md = torch.minimum(md, ((pts - pts[j]) ** 2).sum(-1))
        out.append(ids[torch.tensor(sel)])
    return torch.cat(out)


def resample_uniformly_golden():
    """`--only resample_uniformly`: the reference's ``resample_uniformly`` (point_processing.py:126-166) on a
    ``Pointclouds`` (the only input type its ``wlop`` accepts, :44): WLOP to half the points (farthest sampling
    through the stand-in above, the jitter draw recorded) then ``upsample`` back."""
    from isopoints_b200.structures import Pointclouds
    ref = ref_python.load(frnn_module=_CpuFrnn)
    import pytorch3d.ops.knn as o3dk   # stubs installed by ref_python.load
    import torch_cluster
    PP = ref.point_processing
    torch_cluster.fps = _fps_sequential
    import frnn as frnn_stub            # resample_uniformly does a function-local `import frnn` (:134)
    frnn_stub.frnn_grid_points = _CpuFrnn.frnn_grid_points
    frnn_stub.frnn_gather = _CpuFrnn.frnn_gather

    def _knn_points_cpu(p1, p2, lengths1=None, lengths2=None, K=1, return_nn=False, return_sorted=True, **kw):
        outs_d, outs_i = [], []
        for n in range(p1.shape[0]):
            l1 = p1.shape[1] if lengths1 is None else int(lengths1[n])
            l2 = p2.shape[1] if lengths2 is None else int(lengths2[n])
            d = torch.cdist(p1[n, :l1].double(), p2[n, :l2].double()) ** 2
            v, i = torch.topk(d, min(K, l2), dim=1, largest=False)
            dd = torch.zeros(p1.shape[1], K); ii = torch.zeros(p1.shape[1], K, dtype=torch.long)
            dd[:l1, :v.shape[1]] = v.float(); ii[:l1, :i.shape[1]] = i
            outs_d.append(dd); outs_i.append(ii)
        dists, idx = torch.stack(outs_d), torch.stack(outs_i)
        nn = torch.stack([p2[n][idx[n]] for n in range(p1.shape[0])]) if return_nn else None
        return o3dk._KNN(dists=dists, idx=idx, knn=nn)

    PP.knn_points = _knn_points_cpu
    PP.Pointclouds = Pointclouds
    PP.estimate_pointcloud_normals = lambda p, **kw: torch.zeros_like(p)     # computed and never used (:153-158)
    torch.manual_seed(12)
    sph = torch.nn.functional.normalize(torch.randn(1, 1200, 3), dim=-1) * (1 + 0.01 * torch.randn(1, 1200, 1))
    noise = torch.randn(600, 3)
    real_randn_like = torch.randn_like
    PP.torch.randn_like = lambda x: noise.clone()
    try:
        out = PP.resample_uniformly(Pointclouds(sph.clone()), shrink_ratio=0.5, repulsion_mu=1.0)
    finally:
        PP.torch.randn_like = real_randn_like
    fps_idx = _fps_sequential(sph[0], torch.zeros(1200, dtype=torch.long), 0.5)
    np.savez_compressed(os.path.join(HERE, "resample_uniformly.npz"), P=sph.numpy(), noise=noise.numpy(),
                        fps_idx=fps_idx.numpy(), points=out.points_padded().numpy(),
                        num=out.num_points_per_cloud().numpy())
    print("resample_uniformly:", tuple(out.points_padded().shape), int(out.num_points_per_cloud()[0]))


def main():
    torch.set_num_threads(4)
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only == "pinned":
        return pinned_golden()
    if only == "insert":
        return insert_golden()
    if only == "resample_uniformly":
        return resample_uniformly_golden()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "ewa":
        ref_python.load(frnn_module=_CpuFrnn)
        return ewa_golden()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "ewa_global":
        ref_python.load(frnn_module=_CpuFrnn)
        return ewa_global_golden()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "trace":
        return trace_golden(ref_python.load(frnn_module=_CpuFrnn).levelset_sampling)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "rays":
        return rays_golden(ref_python.load(frnn_module=_CpuFrnn).levelset_sampling)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "offsurface":
        return offsurface_golden(ref_python.load(frnn_module=_CpuFrnn))
    ref = ref_python.load(frnn_module=_CpuFrnn)
    LS = ref.levelset_sampling

    # ---- C1: 4096 points, analytic unit sphere, 10 projection iterations -------------------
    torch.manual_seed(0)
    x = (torch.rand(1, 4096, 3) - 0.5) * 1.5
    proj = LS.UniformProjection(proj_max_iters=10, proj_tolerance=5e-5)
    out = proj.project_points(x.clone(), SphereSDF(), skip_resampling=True, skip_upsampling=True)
    np.savez_compressed(os.path.join(HERE, "proj_sphere.npz"), x=x.numpy(),
                        points=out["levelset_points"].numpy(), normals=out["levelset_normals"].numpy(),
                        mask=out["mask"].numpy())
    print("proj_sphere: converged", float(out["mask"].float().mean()))

    # ---- opaque nn.Module SDF (tiny Siren), ragged batch of two clouds ---------------------
    torch.manual_seed(1)
    net = TinySiren(seed=3)
    xb = (torch.rand(2, 1500, 3) - 0.5) * 1.6
    num = torch.tensor([1500, 1100])
    res = proj._project_points(net, xb.clone(), num, proj_max_iters=6)
    np.savez_compressed(os.path.join(HERE, "proj_siren.npz"), x=xb.numpy(), num=num.numpy(),
                        points=res.points.numpy(), normals=res.normals.numpy(), mask=res.mask.numpy())
    print("proj_siren: converged", float(res.mask.float().mean()))

    # ---- project -> filter -> resample (1 sample_iter, knn_k=8) on the sphere --------------
    torch.manual_seed(2)
    x2 = (torch.rand(1, 2048, 3) - 0.5) * 1.5
    proj2 = LS.UniformProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=8, sample_iters=1)
    out2 = proj2.project_points(x2.clone(), SphereSDF(), skip_upsampling=True)
    # and a 3-iteration resample directly (exercises the neighbourhood refresh on even iterations)
    p0 = proj2._project_points(SphereSDF(), x2.clone(), torch.tensor([2048]))
    r3 = proj2.resample(SphereSDF(), p0.points, p0.normals, torch.tensor([2048]), sample_iters=3)
    np.savez_compressed(os.path.join(HERE, "resample_sphere.npz"), x=x2.numpy(),
                        points=out2["levelset_points"].numpy(), normals=out2["levelset_normals"].numpy(),
                        mask=out2["mask"].numpy(), points3=r3.points.numpy(), normals3=r3.normals.numpy(),
                        mask3=r3.mask.numpy())
    print("resample_sphere: valid", float(out2["mask"].float().mean()))

    # ---- FRNN brute force -------------------------------------------------------------------
    torch.manual_seed(3)
    p = torch.rand(2, 1500, 3)
    lens = torch.tensor([1500, 1200])
    idxs, dists = ref_native.frnn_bf_cpu(p, p, lens, lens, 8, 0.1)
    p2d = torch.rand(1, 1000, 2)
    l2d = torch.tensor([1000])
    idxs2, dists2 = ref_native.frnn_bf_cpu(p2d, p2d, l2d, l2d, 5, 0.05)
    np.savez_compressed(os.path.join(HERE, "frnn_bf.npz"), p=p.numpy(), lens=lens.numpy(), K=8, r=0.1,
                        idxs=idxs.numpy(), dists=dists.numpy(), p2d=p2d.numpy(), idxs2d=idxs2.numpy(),
                        dists2d=dists2.numpy())

    # ---- wlop (ratio = 1: no FPS) and upsample through the reference's own Python ------------
    import pytorch3d.ops.knn as o3dk
    from isopoints_b200.structures import Pointclouds
    PP = ref.point_processing

    def _knn_points_cpu(p1, p2, lengths1=None, lengths2=None, K=1, return_nn=False, return_sorted=True, **kw):
        outs_d, outs_i = [], []
        for n in range(p1.shape[0]):
            l1 = p1.shape[1] if lengths1 is None else int(lengths1[n])
            l2 = p2.shape[1] if lengths2 is None else int(lengths2[n])
            d = torch.cdist(p1[n, :l1].double(), p2[n, :l2].double()) ** 2
            v, i = torch.topk(d, min(K, l2), dim=1, largest=False)
            dd = torch.zeros(p1.shape[1], K); ii = torch.zeros(p1.shape[1], K, dtype=torch.long)
            dd[:l1, :v.shape[1]] = v.float(); ii[:l1, :i.shape[1]] = i
            outs_d.append(dd); outs_i.append(ii)
        dists, idx = torch.stack(outs_d), torch.stack(outs_i)
        nn = torch.stack([p2[n][idx[n]] for n in range(p1.shape[0])]) if return_nn else None
        return o3dk._KNN(dists=dists, idx=idx, knn=nn)

    PP.knn_points = _knn_points_cpu
    PP.Pointclouds = Pointclouds
    torch.manual_seed(5)
    sph = torch.nn.functional.normalize(torch.randn(1, 1500, 3), dim=-1) * (1 + 0.01 * torch.randn(1, 1500, 1))
    noise = torch.randn(1500, 3)
    real_randn_like = torch.randn_like
    PP.torch.randn_like = lambda x: noise.clone()
    try:
        wl = PP.wlop(Pointclouds(sph.clone()), ratio=1.0, neighborhood_size=16, iters=3, repulsion_mu=0.5)
    finally:
        PP.torch.randn_like = real_randn_like
    up_pts, up_num = PP.upsample(sph[:, :1000].clone(), 1300, num_points=torch.tensor([1000]), neighborhood_size=16)
    np.savez_compressed(os.path.join(HERE, "wlop_upsample.npz"), P=sph.numpy(), noise=noise.numpy(),
                        wlop=wl.points_padded().numpy(), up_in=sph[:, :1000].numpy(), up_pts=up_pts.numpy(),
                        up_num=up_num.numpy())
    print("wlop: mean radius", float(wl.points_padded().norm(dim=-1).mean()), "upsample:", tuple(up_pts.shape))

    # ---- EdgeAwareProjection (exact K-NN neighbourhoods; pytorch3d knn_points / knn_gather stand-ins) ---
    LS.knn_points = _knn_points_cpu
    LS.knn_gather = lambda x, idx, lengths=None: torch.stack([x[n][idx[n]] for n in range(x.shape[0])])
    torch.manual_seed(6)
    xe = (torch.rand(1, 1200, 3) - 0.5) * 1.5
    ear = LS.EdgeAwareProjection(proj_max_iters=10, proj_tolerance=5e-5, knn_k=15, sample_iters=2, upsample_ratio=1.5)
    oe = ear.project_points(xe.clone(), SphereSDF())
    np.savez_compressed(os.path.join(HERE, "edge_aware.npz"), x=xe.numpy(), points=oe["levelset_points"].numpy(),
                        mask=oe["mask"].numpy())
    print("edge_aware:", tuple(oe["levelset_points"].shape), float(oe["mask"].float().mean()))

    # ---- splat forward: reference naive CPU twin --------------------------------------------
    C = ref_native.dss_C()
    S, K = 48, 4
    inp = make_splat_inputs(n_views=2, pts_per_view=[700, 500], S=S, seed=4, sigma_px=1.5)
    t = {k: torch.as_tensor(v) for k, v in inp.items()}
    idx, zbuf, qv, occ = C._splat_points_naive(t["points"], t["ellipse"], t["cutoff"], t["radii"],
                                               t["first_idx"], t["num_points"], 0.05, S, K)
    np.savez_compressed(os.path.join(HERE, "splat_naive_cpu.npz"), S=S, K=K, depth_merging_thres=0.05,
                        idx=idx.numpy(), zbuf=zbuf.numpy(), qvalue=qv.numpy(), occ=occ.numpy(), **inp)
    print("splat: occupied", float(occ.mean()))
    ewa_golden()
    trace_golden(LS)
    rays_golden(LS)
    offsurface_golden(ref)
    pinned_golden()
    insert_golden()
    resample_uniformly_golden()


if __name__ == "__main__":
    main()
