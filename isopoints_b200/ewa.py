"""Per-point EWA splat parameters and the renderable filter on libisob200.so -- the part of
``SurfaceSplatting`` (DSS/core/rasterizer.py:79-661) that runs right before the splat kernel.

Mirrors, with the reference's names and argument meaning:

* ``SurfaceSplatting._compute_isotropic_Vrk`` (:344-400) -> ``compute_isotropic_vrk_h`` (the FRNN K = 7 self
  query of isopoints_b200.frnn + one kernel; the tangent frame lives inside the parameter kernel),
* ``SurfaceSplatting._get_per_point_info`` (:514-563) -> ``get_per_point_info`` (one kernel instead of
  ~30 PyTorch ops and a batched LU of 2x2 matrices),
* ``SurfaceSplatting.filter_renderable`` (:220-255) -> ``filter_renderable`` (mask kernel + ordered
  compaction; the reference rebuilds Python lists of boolean-indexed padded tensors per view),
* ``SurfaceSplatting.forward`` (:584-661) -> ``SurfaceSplatting.forward``: filter -> parameters -> screen
  transform (PyTorch, differentiable, as in the reference) -> ``rasterize_elliptical_points``,
* ``SurfaceSplattingRenderer.forward`` (DSS/core/renderer.py:36-82) -> ``SurfaceSplattingRenderer.forward``:
  rasterise + ``splat.blend_rgba`` -> (N,S,S,4) RGBA.

Cameras are duck-typed on what the reference calls (pytorch3d's camera API; pytorch3d itself is not a
dependency): ``get_full_projection_transform().get_matrix()`` and
``get_world_to_view_transform().get_matrix()`` -> (N or 1, 4, 4) in the row-vector convention,
optional ``znear`` / ``zfar`` attributes.
There is no CPU / PyTorch fallback: CPU tensors raise.
"""
from types import SimpleNamespace

import torch

from . import _ext
from .frnn import frnn_grid_points
from .splat import (PointFragments, blend_rgba, gather_with_neg_idx, rasterize_elliptical_points,
                    visibility_mask)
from .structures import Pointclouds, packed_to_padded

MAX_VIEWS = 64   # csrc/ewa.cu: camera matrices are staged in shared memory


def _f32(t, name):
    if not t.is_cuda:
        raise TypeError("%s: for now only cuda version is supported" % name)
    return t.detach().to(torch.float32).contiguous()


def _views(first_idx, num_points, dev):
    first_idx = first_idx.to(device=dev, dtype=torch.int64).contiguous()
    num_points = num_points.to(device=dev, dtype=torch.int64).contiguous()
    if first_idx.numel() > MAX_VIEWS:
        raise ValueError("at most %d point clouds / views per call, got %d" % (MAX_VIEWS, first_idx.numel()))
    return first_idx, num_points


def _cam_matrix(m, n_views, dev, name):
    m = _f32(m.to(dev), name)
    if m.dim() == 2:
        m = m[None]
    if m.shape[0] not in (1, n_views) or tuple(m.shape[1:]) != (4, 4):
        raise ValueError("%s: expected (%d or 1, 4, 4), got %s" % (name, n_views, tuple(m.shape)))
    return m


def compute_isotropic_vrk_h(points_padded, num_points_per_cloud, frnn_radius, first_idx=None):
    """h_k of _compute_isotropic_Vrk (rasterizer.py:358-386), packed (P,): half the largest squared
    distance to the 6 nearest neighbours within ``frnn_radius``, clamped to [5e-5, 0.01]."""
    if frnn_radius <= 0:
        raise NotImplementedError("frnn_radius <= 0 selects pytorch3d's brute-force knn_points in the reference "
                                  "(rasterizer.py:365-369); only the FRNN branch (:370-373) is built")
    dev = points_padded.device
    pts = _f32(points_padded, "points_padded")
    num = num_points_per_cloud.to(device=dev, dtype=torch.int64).contiguous()
    if first_idx is None:
        first_idx = torch.cumsum(num, 0) - num
    first_idx, num = _views(first_idx, num, dev)
    K = 7
    sq_dist = frnn_grid_points(pts, pts, num, num, K=K, r=frnn_radius)[0]
    P = int(num.sum().item())       # packed rows (a single cloud may be padded beyond its length)
    h = torch.empty(P, dtype=torch.float32, device=dev)
    _ext.check(_ext.lib().isob200_ewa_vrk_h(_ext.ptr(sq_dist), _ext.ptr(first_idx), _ext.ptr(num), num.numel(),
                                            pts.shape[1], K, P, _ext.ptr(h), _ext.stream(dev)))
    return h


def global_vrk_h_from_sq_dist(sq_dist, num_points_per_cloud):
    """One h per cloud from the K = 7 self-query distances (N, P_max, 7) -- ``_compute_global_Vrk``
    (rasterizer.py:303-322): drop the self column, clouds with fewer than 7 points get 1e-3 everywhere, half the
    largest neighbour distance per row, MEAN OVER THE PADDED ROWS (their distances are the -1 padding, as in the
    reference), clamped to [5e-5, 1e-3].  Returns (N,)."""
    sq = sq_dist[:, :, 1:].clone()
    sq[num_points_per_cloud.to(sq.device) < 7] = 1e-3
    h = 0.5 * sq.max(dim=-1, keepdim=True)[0]
    return h.mean(dim=1, keepdim=True).clamp(5e-5, 1e-3).view(-1)


def compute_global_vrk_h(points_padded, num_points_per_cloud, frnn_radius):
    """h_k of the view-invariant V_k^r (``Vrk_invariant=True``, rasterizer.py:292-343), packed (P,): the per-cloud
    value of ``global_vrk_h_from_sq_dist`` repeated for the cloud's points."""
    if frnn_radius <= 0:
        raise NotImplementedError("frnn_radius <= 0 selects pytorch3d's brute-force knn_points in the reference "
                                  "(rasterizer.py:306-311); only the FRNN branch (:312-315) is built")
    pts = _f32(points_padded, "points_padded")
    num = num_points_per_cloud.to(device=pts.device, dtype=torch.int64).contiguous()
    sq_dist = frnn_grid_points(pts, pts, num, num, K=7, r=frnn_radius)[0]
    return torch.repeat_interleave(global_vrk_h_from_sq_dist(sq_dist, num), num)


def get_per_point_info(points_packed, normals_packed, cloud_to_packed_first_idx, proj_matrix, vrk_h,
                       image_size, antialiasing_sigma=1.0, cutoff_threshold=1.0):
    """_get_per_point_info (rasterizer.py:514-563) with the isotropic V_k^r: returns the reference's dict
    {"radii" (P,2), "ellipse_params" (P,3), "cutoff_threshold" (P,), "scaler" (P,)}; no gradients
    (the reference runs it under no_grad and detaches)."""
    dev = points_packed.device
    pts = _f32(points_packed, "points")
    nrm = _f32(normals_packed, "normals")
    P = pts.shape[0]
    if pts.dim() != 2 or pts.shape[1] != 3 or tuple(nrm.shape) != (P, 3):
        raise RuntimeError("expected packed points / normals of size (P, 3), got %s / %s"
                           % (tuple(pts.shape), tuple(nrm.shape)))
    first = cloud_to_packed_first_idx.to(device=dev, dtype=torch.int64).contiguous()
    if first.numel() > MAX_VIEWS:
        raise ValueError("at most %d point clouds / views per call, got %d" % (MAX_VIEWS, first.numel()))
    proj = _cam_matrix(proj_matrix, first.numel(), dev, "proj_matrix")
    h = _f32(vrk_h, "vrk_h").view(-1)
    if h.numel() != P:
        raise RuntimeError("vrk_h: expected %d values, got %d" % (P, h.numel()))
    radii = torch.empty(P, 2, dtype=torch.float32, device=dev)
    ellipse = torch.empty(P, 3, dtype=torch.float32, device=dev)
    cutoff = torch.empty(P, dtype=torch.float32, device=dev)
    scaler = torch.empty(P, dtype=torch.float32, device=dev)
    pixel_size = 2.0 / image_size
    _ext.check(_ext.lib().isob200_ewa_point_params(
        _ext.ptr(pts), _ext.ptr(nrm), _ext.ptr(first), first.numel(), P, _ext.ptr(proj), proj.shape[0],
        _ext.ptr(h), float(antialiasing_sigma) * pixel_size ** 2, float(cutoff_threshold), _ext.ptr(radii),
        _ext.ptr(ellipse), _ext.ptr(cutoff), _ext.ptr(scaler), _ext.stream(dev)))
    return {"radii": radii, "ellipse_params": ellipse, "cutoff_threshold": cutoff, "scaler": scaler}


def renderable_mask(points_packed, normals_packed, cloud_to_packed_first_idx, w2v_matrix, znear=1.0, zfar=100.0,
                    backface_culling=False):
    """The packed mask filter_renderable (rasterizer.py:220-255) ends up with, and the survivors per view
    (int32 (N,), on the device)."""
    dev = points_packed.device
    pts = _f32(points_packed, "points")
    P = pts.shape[0]
    first = cloud_to_packed_first_idx.to(device=dev, dtype=torch.int64).contiguous()
    if first.numel() > MAX_VIEWS:
        raise ValueError("at most %d point clouds / views per call, got %d" % (MAX_VIEWS, first.numel()))
    w2v = _cam_matrix(w2v_matrix, first.numel(), dev, "w2v_matrix")
    nrm = nmat = None
    if backface_culling:
        if normals_packed is None:
            raise ValueError("backface_culling needs normals")
        nrm = _f32(normals_packed, "normals")
        # Transform3d.transform_normals [pytorch3d, third party]: normals @ inverse(M)[:3,:3]^T
        nmat = torch.inverse(w2v)[:, :3, :3].transpose(1, 2).contiguous()
    mask = torch.empty(P, dtype=torch.uint8, device=dev)
    kept = torch.empty(first.numel(), dtype=torch.int32, device=dev)
    _ext.check(_ext.lib().isob200_renderable_mask(
        _ext.ptr(pts), _ext.ptr(nrm), _ext.ptr(first), first.numel(), P, _ext.ptr(w2v), _ext.ptr(nmat),
        w2v.shape[0], float(znear), float(zfar), _ext.ptr(mask), _ext.ptr(kept), _ext.stream(dev)))
    return mask.view(torch.bool), kept


def _compact_rows3(points, normals, mask_u8, n_keep):
    """Ordered compaction of (P,3) points [and normals] by a uint8 mask (isob200_compact_valid)."""
    lib = _ext.lib()
    dev = points.device
    P = points.shape[0]
    out_p = torch.empty(P, 3, dtype=torch.float32, device=dev)
    out_n = None if normals is None else torch.empty(P, 3, dtype=torch.float32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.isob200_project_step_ws_bytes(P)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _ext.check(lib.isob200_compact_valid(_ext.ptr(points), _ext.ptr(normals), _ext.ptr(mask_u8), P, _ext.ptr(out_p),
                                         _ext.ptr(out_n), _ext.ptr(cnt), _ext.ptr(ws), ws_bytes, _ext.stream(dev)))
    return out_p[:n_keep], (None if out_n is None else out_n[:n_keep])


class SurfaceSplatting:
    """DSS/core/rasterizer.py:79-661 on the kernels of this package: the isotropic V_k^r (the default of
    PointsRasterizationSettings: Vrk_isotropic=True, Vrk_invariant=False) and the view-invariant one
    (Vrk_invariant=True, one h per cloud); the anisotropic variant raises NotImplementedError."""

    def __init__(self, cameras=None, raster_settings=None, frnn_radius=0.2):
        if raster_settings is None:
            raster_settings = PointsRasterizationSettings()
        self.cameras = cameras
        self.raster_settings = raster_settings
        self.frnn_radius = frnn_radius
        self._Vrk_h = None

    # ---- pieces, named as in the reference --------------------------------------------------------
    def _compute_isotropic_Vrk_h(self, pointclouds, refresh=True):
        """rasterizer.py:358-386 incl. the ``refresh`` cache (:358-360)."""
        P = int(pointclouds.num_points_per_cloud().sum().item())
        if not refresh and self._Vrk_h is not None and P == self._Vrk_h.shape[0]:
            return self._Vrk_h
        self._Vrk_h = compute_isotropic_vrk_h(pointclouds.points_padded(), pointclouds.num_points_per_cloud(),
                                              self.frnn_radius, pointclouds.cloud_to_packed_first_idx())
        return self._Vrk_h

    def _get_per_point_info(self, pointclouds, **kwargs):
        rs = kwargs.get("raster_settings", self.raster_settings)
        cameras = kwargs.get("cameras", self.cameras)
        if getattr(rs, "Vrk_invariant", False):       # checked first, as in rasterizer.py:418-419
            h = compute_global_vrk_h(pointclouds.points_padded(), pointclouds.num_points_per_cloud(), self.frnn_radius)
        elif getattr(rs, "Vrk_isotropic", True):
            h = self._compute_isotropic_Vrk_h(pointclouds, refresh=kwargs.get("refresh", True))
        else:
            raise NotImplementedError("the anisotropic V_k^r (rasterizer.py:256-290) is not built")
        return get_per_point_info(pointclouds.points_packed(), pointclouds.normals_packed(),
                                  pointclouds.cloud_to_packed_first_idx(),
                                  cameras.get_full_projection_transform().get_matrix(), h, rs.image_size,
                                  rs.antialiasing_sigma, rs.cutoff_threshold)

    def filter_renderable(self, point_clouds, point_clouds_filter=None, **kwargs):
        """rasterizer.py:220-255: (new point clouds, packed mask over the points that entered the depth /
        back-face test).  With a ``point_clouds_filter`` (DSS/core/cloud.py PointCloudsFilters) its visibility is
        reset to all-False and its ``activation`` filter is applied first (:231-235)."""
        rs = kwargs.get("raster_settings", self.raster_settings)
        cameras = kwargs.get("cameras", self.cameras)
        n = point_clouds.num_points_per_cloud()
        if point_clouds.isempty():
            return point_clouds, torch.full((int(n.sum().item()),), True, dtype=torch.bool, device=point_clouds.device)
        if point_clouds_filter is not None:
            point_clouds_filter.set_filter(visibility=torch.full(
                (len(point_clouds), int(n.max().item())), False, dtype=torch.bool, device=point_clouds.device))
            point_clouds = point_clouds_filter.filter_with(point_clouds, ("activation",))
        w2v = cameras.get_world_to_view_transform().get_matrix()
        if w2v.shape[0] != len(point_clouds):
            point_clouds = point_clouds.extend(w2v.shape[0])     # :241-245
        znear = getattr(cameras, "znear", kwargs.get("znear", 1.0))
        zfar = getattr(cameras, "zfar", kwargs.get("zfar", 100.0))
        pts = point_clouds.points_packed()
        nrm = point_clouds.normals_packed()
        first = point_clouds.cloud_to_packed_first_idx()
        backface = bool(getattr(rs, "backface_culling", True))
        zn = znear.reshape(-1).float() if torch.is_tensor(znear) else None
        zf = zfar.reshape(-1).float() if torch.is_tensor(zfar) else None
        per_view = ((zn is not None and zn.numel() > 1 and bool((zn != zn[0]).any())) or
                    (zf is not None and zf.numel() > 1 and bool((zf != zf[0]).any())))
        if not per_view:
            znear = float(zn[0]) if zn is not None else float(znear)
            zfar = float(zf[0]) if zf is not None else float(zfar)
            mask, kept = renderable_mask(pts, nrm, first, w2v, znear, zfar, backface)
            kept = kept.tolist()                                  # one read-back: the new cloud sizes
        else:
            # per-view clip planes (camera batches built with znear / zfar tensors): the mask kernel takes one
            # pair per launch, so the views go through it one by one
            nums = point_clouds.num_points_per_cloud().tolist()
            firsts = first.tolist()
            masks, kept = [], []
            for v, (f0, nv) in enumerate(zip(firsts, nums)):
                zv = float(zn[v if zn.numel() > 1 else 0]) if zn is not None else float(znear)
                fv = float(zf[v if zf.numel() > 1 else 0]) if zf is not None else float(zfar)
                mv, kv = renderable_mask(pts[f0:f0 + nv], None if nrm is None else nrm[f0:f0 + nv],
                                         torch.zeros(1, dtype=torch.int64, device=pts.device), w2v[v:v + 1], zv, fv,
                                         backface)
                masks.append(mv)
                kept.append(int(kv.item()))
            mask = torch.cat(masks)
        total = sum(kept)
        if total == pts.shape[0]:
            return point_clouds, mask
        m8 = mask.view(torch.uint8)
        new_pts, new_nrm = _compact_rows3(_f32(pts, "points"), None if nrm is None else _f32(nrm, "normals"),
                                          m8, total)
        if pts.requires_grad:                                      # keep the autograd link to the input points
            new_pts = pts[mask]
        if nrm is not None and nrm.requires_grad:                  # ... and to the normals (boolean indexing, :250)
            new_nrm = nrm[mask]
        feats = point_clouds.features_packed()
        parts = lambda t: None if t is None else list(torch.split(t, kept))   # noqa: E731
        new = Pointclouds(points=parts(new_pts), normals=parts(new_nrm),
                          features=parts(None if feats is None else feats[mask]))
        return new, mask

    def transform(self, point_clouds, **kwargs):
        """PointsRasterizer.transform [pytorch3d, third party, restated]: NDC xy from the full projection,
        z = view-space depth; differentiable PyTorch ops as in the reference.  Returns a Pointclouds."""
        cameras = kwargs.get("cameras", self.cameras)
        pts = point_clouds.points_packed()
        b = point_clouds.packed_to_cloud_idx()
        hom = torch.cat([pts, torch.ones_like(pts[:, :1])], -1)
        full = cameras.get_full_projection_transform().get_matrix().to(pts)
        w2v = cameras.get_world_to_view_transform().get_matrix().to(pts)
        full = full[b] if full.shape[0] > 1 else full.expand(pts.shape[0], 4, 4)
        w2v = w2v[b] if w2v.shape[0] > 1 else w2v.expand(pts.shape[0], 4, 4)
        ndc = torch.bmm(hom[:, None, :], full)[:, 0]
        view = torch.bmm(hom[:, None, :], w2v)[:, 0]
        screen = torch.cat([ndc[:, :2] / ndc[:, 3:], view[:, 2:3] / view[:, 3:]], -1)
        return Pointclouds(points=list(torch.split(screen, point_clouds.num_points_per_cloud().tolist())))

    def _empty_fragments(self, batch_size, **kwargs):
        """rasterizer.py:565-582."""
        rs = kwargs.get("raster_settings", self.raster_settings)
        dev = kwargs.get("device", "cuda:%d" % torch.cuda.current_device())
        S, K = rs.image_size, rs.points_per_pixel
        idx = torch.full((batch_size, S, S, K), -1, dtype=torch.long, device=dev)
        zbuf = torch.full((batch_size, S, S, K), -1.0, dtype=torch.float, device=dev)
        qvalue = torch.full((batch_size, S, S, K), -1.0, dtype=torch.float, device=dev)
        occ = torch.full((batch_size, S, S), 0, dtype=torch.float, device=dev)
        return PointFragments(idx=idx, zbuf=zbuf, qvalue=qvalue, scaler=qvalue, occupancy=occ)

    def forward(self, point_clouds, point_clouds_filter=None, **kwargs):
        """rasterizer.py:584-661 -> (PointFragments, filtered point clouds).  A ``point_clouds_filter`` receives
        the per-point visibility of the input clouds as its padded ``visibility`` filter (:642-650)."""
        rs = kwargs.get("raster_settings", self.raster_settings)
        cameras = kwargs.get("cameras", self.cameras)
        n_cameras = int(cameras.get_world_to_view_transform().get_matrix().shape[0])
        if len(point_clouds) == 1 and n_cameras > 1:          # :597-598: one copy of the cloud per camera, so the
            point_clouds = point_clouds.extend(n_cameras)     # visibility below has one row per VIEW
        filtered, mask_filtered = self.filter_renderable(point_clouds, point_clouds_filter, **kwargs)
        if filtered.isempty():
            return self._empty_fragments(len(filtered), device=filtered.device, **kwargs), filtered
        with torch.no_grad():
            info = self._get_per_point_info(filtered, **kwargs)
        screen = self.transform(filtered, **kwargs)
        # the renderer asks for its RGBA blend (renderer.py:53-78) in the same pass over the pixels
        blend = None
        if kwargs.get("blend_rgb", False) and filtered.features_packed() is not None:
            from .splat import NORM_WEIGHT_EPS
            blend = (info["scaler"], filtered.features_packed()[:, :3], NORM_WEIGHT_EPS)
        out = rasterize_elliptical_points(
            screen, info["ellipse_params"], info["cutoff_threshold"], info["radii"],
            depth_merging_threshold=rs.depth_merging_threshold, image_size=rs.image_size,
            points_per_pixel=rs.points_per_pixel, bin_size=rs.bin_size, max_points_per_bin=rs.max_points_per_bin,
            radii_backward_scaler=rs.radii_backward_scaler, clip_pts_grad=rs.clip_pts_grad, blend=blend)
        idx, zbuf, qvalue, occ = out[:4]
        frag_scaler = gather_with_neg_idx(info["scaler"], 0, idx.view(-1).long()).view_as(qvalue)   # :634-636
        fragments = PointFragments(idx=idx, zbuf=zbuf, qvalue=qvalue, scaler=frag_scaler, occupancy=occ)
        # per-point scaler of these fragments (and the images, when blended here) for the renderer
        self._last = (fragments, info["scaler"], out[4] if blend is not None else None)
        if point_clouds_filter is not None:
            # visibility of the points that survived the renderable filter, scattered back to the points that
            # entered it, then padded with the first indices of the (extended) input clouds (:642-650): (B, max_P)
            vis = visibility_mask(idx.detach(), int(filtered.num_points_per_cloud().sum().item()), occ.detach())
            full = torch.zeros_like(mask_filtered)
            full[mask_filtered] = vis
            max_p = int(point_clouds.num_points_per_cloud().max().item())
            padded = packed_to_padded(full.float(), point_clouds.cloud_to_packed_first_idx(), max_p).bool()
            point_clouds_filter.set_filter(visibility=padded)
        return fragments, filtered

    __call__ = forward


class SurfaceSplattingRenderer:
    """DSS/core/renderer.py:14-82: rasterise, weight the fragments with scaler * exp(-Q/2), composite the point
    features with pytorch3d's NormWeightedCompositor rule and append the occupancy as alpha -- the last three
    steps in one kernel (``splat.blend_rgba``), differentiable w.r.t. the features; gradients to the points
    come from the occupancy / depth terms of the rasteriser exactly as in the reference."""

    def __init__(self, rasterizer, compositor="norm_weighted", antialiasing_sigma: float = 1.0,
                 density: float = 1e-4, frnn_radius=-1):
        if compositor is None:
            raise NotImplementedError("compositor=None selects pytorch3d's un-normalised weighted_sum "
                                      "(renderer.py:59-65); only the NormWeightedCompositor rule is built")
        self.rasterizer = rasterizer
        self.compositor = compositor
        self.cameras = rasterizer.cameras
        self.antialiasing_sigma = antialiasing_sigma
        self.density = density
        self.frnn_radius = frnn_radius

    def forward(self, point_clouds, **kwargs):
        if point_clouds.isempty():
            return None
        fragments = kwargs.get("fragments", None)
        if fragments is None:
            fragments, point_clouds = self.rasterizer(point_clouds, blend_rgb=True, **kwargs)
        last = getattr(self.rasterizer, "_last", None)
        if last is not None and last[0] is fragments and len(last) > 2 and last[2] is not None:
            images = last[2]                      # blended in the raster kernel's epilogue
            return (images, fragments) if kwargs.get("verbose", False) else images
        if last is not None and last[0] is fragments:
            scaler = last[1]
        else:
            # fragments from elsewhere: every fragment of a point carries that point's scaler (rasterizer.py:634)
            P = int(point_clouds.num_points_per_cloud().sum().item())
            m = fragments.idx >= 0
            scaler = fragments.scaler.new_zeros(P)
            scaler[fragments.idx[m].long()] = fragments.scaler[m]
        pts_rgb = point_clouds.features_packed()[:, :3]
        images = blend_rgba(fragments.idx, fragments.qvalue, fragments.occupancy, scaler, pts_rgb)
        if kwargs.get("verbose", False):
            return images, fragments
        return images

    __call__ = forward


def get_visible_points(point_clouds, cameras, depth_merge_threshold=0.05, return_mask=False):
    """DSS/utils/__init__.py:699-711: the points of ``point_clouds`` that own a fragment of an occupied pixel
    when splatted at 256 x 256 with back-face culling -- as a new Pointclouds (and, with ``return_mask``, the
    padded visibility mask)."""
    from .cloud import PointCloudsFilters
    splatter = SurfaceSplatting(raster_settings=PointsRasterizationSettings(
        depth_merging_threshold=depth_merge_threshold, image_size=256, cutoff_threshold=1.0, backface_culling=True))
    pcl_filter = PointCloudsFilters(device=point_clouds.device)
    with torch.no_grad():
        splatter(point_clouds, cameras=cameras, point_clouds_filter=pcl_filter)
    visible = pcl_filter.filter_with(point_clouds, ("visibility",))
    if return_mask:
        return visible, pcl_filter.visibility
    return visible


def PointsRasterizationSettings(**kw):
    """DSS/core/rasterizer.py:38-100, same defaults."""
    d = dict(backface_culling=True, cutoff_threshold=1.0, depth_merging_threshold=0.05, Vrk_invariant=False,
             Vrk_isotropic=True, radii_backward_scaler=10.0, image_size=256, points_per_pixel=8, bin_size=0,
             max_points_per_bin=None, clip_pts_grad=-1.0, antialiasing_sigma=1.0)
    unknown = set(kw) - set(d)
    if unknown:
        raise TypeError("unknown raster settings: %s" % sorted(unknown))
    d.update(kw)
    return SimpleNamespace(**d)


__all__ = ["SurfaceSplatting", "SurfaceSplattingRenderer", "get_visible_points", "compute_global_vrk_h",
           "global_vrk_h_from_sq_dist", "PointsRasterizationSettings", "compute_isotropic_vrk_h", "get_per_point_info",
           "renderable_mask", "visibility_mask", "packed_to_padded"]
