"""Visible iso-points, and the free-space / in-surface training samples taken from them.

Mirrors ``Model.get_visible_iso_points`` (DSS/models/combined_modeling.py:390-455; see ``get_visible_iso_points``
below) and ``Model.sample_offsurface_using_isopoints`` (:237-388), the consumers of ``get_visible_points`` in
``train_mvr.py``'s step.  The sampler:

* off-surface: for the pixels outside the ground-truth mask, a random point on the segment the camera ray cuts
  out of the bounding cube (``intersection_with_unit_cube``, DSS/utils/__init__.py:402-483) -- plus, optionally,
  the iso-points that project outside the mask;
* in-surface: for (a capped number of) pixels inside the mask, the sample with the smallest SDF among
  ``n_points_per_ray`` candidates between the front-facing and the occluded iso-point closest to the ray.

What runs on this package's kernels:

* the visibility passes (front view and the camera mirrored to the back): ``ewa.get_visible_points``,
* the closest iso-point to every ray: ``closest_point_to_rays`` -> ``isob200_ray_nearest_point`` (csrc/rays.cu), one
  warp per ray over shared-memory tiles of the points, instead of the reference's two dense (R, M) matrices and
  ``topk`` per view ("TODO: faster search", combined_modeling.py:325),
* the SDF of the R x n candidates: the forward-only fused SIREN kernel (``siren.sdf_fn``) when the decoder is the
  reference's Siren, the decoder's own forward otherwise.

The camera object is duck-typed on what the reference calls: ``R`` (B,3,3), ``T`` (B,3), ``get_camera_center()``,
``unproject_points(xy_depth, scaled_depth_input=False)``, ``transform_points(p)``, ``clone()``, optionally
``principal_point``.  ``model`` is duck-typed on the attributes the method reads: ``_points`` (Pointclouds with
normals), ``decoder``, ``max_points_per_pass``, ``object_bounding_sphere`` and
``renderer.rasterizer.raster_settings.depth_merging_threshold``; bind it with
``Model.sample_offsurface_using_isopoints = offsurface.sample_offsurface_using_isopoints``.
"""
import torch
import torch.nn.functional as F

from . import _ext, siren

__all__ = ["closest_point_to_rays", "intersection_with_unit_cube", "get_tensor_values", "insurface_segments",
           "sample_offsurface_using_isopoints", "subsample_randomly", "get_visible_iso_points"]


def closest_point_to_rays(origins, ray_directions, points, return_dist=False):
    """For every ray r (origin ``origins`` (3,) / (1,3) shared or (R,3), unit direction ``ray_directions`` (R,3))
    the point of ``points`` (M,3) with the smallest squared distance to the ray's line,
    ``|p - o|^2 - ((p - o).d)^2`` (combined_modeling.py:331-347).  Returns ``(t_sq (R,), idx (R,) int64)`` with
    ``t_sq = ((p - o).d)^2`` of that point -- the reference's ``torch.gather(ray_sq, 1, nn_idx)`` -- and, with
    ``return_dist``, the squared distance too.  M = 0 gives t_sq 0 / idx -1."""
    _ext.require_cuda(ray_directions)
    d = ray_directions.reshape(-1, 3).to(torch.float32).contiguous()
    o = origins.reshape(-1, 3).to(torch.float32).contiguous()
    p = points.reshape(-1, 3).to(torch.float32).contiguous()
    R, M = d.shape[0], p.shape[0]
    if o.shape[0] not in (1, R):
        raise ValueError("closest_point_to_rays: %d origins for %d rays" % (o.shape[0], R))
    t_sq = torch.empty((R,), dtype=torch.float32, device=d.device)
    dist = torch.empty((R,), dtype=torch.float32, device=d.device)
    idx = torch.empty((R,), dtype=torch.int32, device=d.device)
    if R:
        _ext.check(_ext.lib().isob200_ray_nearest_point(_ext.ptr(o), o.shape[0], _ext.ptr(d), R, _ext.ptr(p), M,
                                                        _ext.ptr(t_sq), _ext.ptr(dist), _ext.ptr(idx),
                                                        _ext.stream(d.device)))
    if return_dist:
        return t_sq, idx.long(), dist
    return t_sq, idx.long()


def intersection_with_unit_cube(ray0, ray_direction, side_length=1.0, padding=0.1, eps=1e-6):
    """DSS/utils/__init__.py:402-483: entry and exit point of every ray ``ray0 + d * ray_direction`` (shape
    (..., 3)) with the axis-aligned cube of half side ``side_length / 2 + padding / 2``, ordered along the ray,
    and the mask of rays that cross it (exactly two face hits).  Rays that miss get zeros."""
    ray0 = ray0.expand_as(ray_direction)
    half = side_length / 2 + padding / 2
    plane = torch.cat([torch.full_like(ray_direction, half), torch.full_like(ray_direction, -half)], dim=-1)
    d_hit = (plane - torch.cat([ray0, ray0], dim=-1)) / torch.cat([ray_direction, ray_direction], dim=-1)   # (...,6)
    p_hit = ray0.unsqueeze(-2) + d_hit.unsqueeze(-1) * ray_direction.unsqueeze(-2)                           # (...,6,3)
    on_face = ((p_hit <= half + eps) & (p_hit >= -(half + eps))).all(dim=-1)
    crosses = on_face.sum(-1) == 2
    pair = p_hit[crosses][on_face[crosses]].view(-1, 2, 3)
    ends = torch.zeros(ray_direction.shape[:-1] + (2, 3), dtype=ray_direction.dtype, device=ray_direction.device)
    ends[crosses] = pair
    along = torch.zeros(ray_direction.shape[:-1] + (2,), dtype=ray_direction.dtype, device=ray_direction.device)
    norm_ray = torch.norm(ray_direction[crosses], dim=-1)
    o = ray0[crosses]
    along[crosses] = torch.stack([torch.norm(pair[:, 0] - o, dim=-1) / norm_ray,
                                  torch.norm(pair[:, 1] - o, dim=-1) / norm_ray], dim=-1)
    along, order = along.sort()
    ends = torch.gather(ends, -2, order.unsqueeze(-1).expand_as(ends))
    first, second = ends.unbind(dim=-2)
    return first, second, crosses


def get_tensor_values(tensor, p, mode="bilinear", squeeze_channel_dim=False):
    """DSS/utils/__init__.py:325-375 (the grid_sample branch): values of ``tensor`` (B,C,H,W) at ``p`` (B,N,2) in
    [-1, 1], reflection padding -> (B,N,C)."""
    values = F.grid_sample(tensor, p.unsqueeze(1), mode=mode, padding_mode="reflection", align_corners=False)
    values = values.squeeze(2).permute(0, 2, 1)
    return values.squeeze(-1) if squeeze_channel_dim else values


def insurface_segments(cam_pos, rays, frontal_points, occluded_points):
    """combined_modeling.py:326-352 for one view: unit ``rays`` (R,3) from ``cam_pos`` (3,), the front-facing and
    the occluded visible iso-points (M,3).  Returns ``(t0_sq, t1_sq)`` (R,1): squared ray length at the closest
    frontal / occluded point -- the segment of the ray that lies inside the surface when ``t0_sq < t1_sq``."""
    t1_sq, _ = closest_point_to_rays(cam_pos, rays, occluded_points)
    t0_sq, _ = closest_point_to_rays(cam_pos, rays, frontal_points)
    return t0_sq.view(-1, 1), t1_sq.view(-1, 1)


def _eps_sqrt(x, eps=1e-17):
    """DSS/utils/mathHelper.py:20-25."""
    return torch.clamp(x.abs(), eps)


def sample_offsurface_using_isopoints(model, pixels, mask_img, cameras, n_points_per_ray=64,
                                      max_insurface_per_batch=None, iso_pcl=None, rand=None,
                                      visible_points_fn=None):
    """combined_modeling.py:237-388.  ``pixels`` (B,N,2) NDC, ``mask_img`` (B,1,H,W).  Returns
    ``(p_offsurface (N2,3), p_insurface (N1,3), num_off_per_batch (B,), num_ins_per_batch (B,))``.
    ``rand`` (B,N) in [0,1): the draw of the off-surface positions (the reference calls ``torch.rand_like``;
    injectable for reproducible tests).  ``visible_points_fn`` defaults to ``ewa.get_visible_points``."""
    if visible_points_fn is None:
        from .ewa import get_visible_points as visible_points_fn
    with torch.no_grad():
        B = cameras.R.shape[0]
        cam_pos = cameras.get_camera_center()
        sample_points = cameras.unproject_points(
            torch.cat([-pixels, pixels.new_ones(pixels.shape[:-1] + (1,))], dim=-1), scaled_depth_input=False)
        cam_ray = F.normalize(sample_points - cam_pos.unsqueeze(1), dim=-1)

        def in_gt_mask(world_points):
            screen = cameras.transform_points(world_points)
            return get_tensor_values(mask_img.float(), (-screen[..., :2]).clamp(-1.0, 1.0),
                                     squeeze_channel_dim=True).bool()
        iso_mask = in_gt_mask(sample_points)

        # off-surface: a random point between the two cube intersections of every ray outside the mask
        first, second, crosses = intersection_with_unit_cube(cam_pos.view(B, 1, 3), cam_ray,
                                                             side_length=model.object_bounding_sphere * 2)
        lengths = torch.norm(second - first, dim=-1)
        u = torch.rand_like(lengths) if rand is None else rand.to(lengths)
        off_mask = (~iso_mask) & crosses
        p_off = ((u * lengths).unsqueeze(-1) * cam_ray + first)[off_mask]
        num_off = off_mask.sum(dim=1)
        if iso_pcl is not None:          # + the iso-points that fall outside the 2-D mask
            n_iso = iso_pcl.num_points_per_cloud()
            padded = iso_pcl.points_padded()
            valid = torch.arange(padded.shape[1], device=padded.device)[None, :] < n_iso[:, None]
            outside = (~in_gt_mask(padded)) & valid
            p_off = torch.cat([p_off, padded[outside]], dim=0)
            num_off = num_off + outside.sum(dim=-1)

        # in-surface rays: the first max_insurface_per_batch[b] pixels inside the mask
        ins_mask = torch.zeros_like(iso_mask)
        if max_insurface_per_batch is not None:
            for b in range(B):
                cap = min(int(max_insurface_per_batch[b]), int(iso_mask[b].sum()))
                ins_mask[b][iso_mask[b].nonzero(as_tuple=False)[:cap]] = True
        targets = [sample_points[b][ins_mask[b]] for b in range(B)]

        thr = model.renderer.rasterizer.raster_settings.depth_merging_threshold
        frontal = visible_points_fn(model._points, cameras, depth_merge_threshold=thr)
        back = cameras.clone()           # the same camera mirrored to the far side (:307-313)
        if hasattr(cameras, "principal_point"):
            back.principal_point[:, 1] = -cameras.principal_point[:, 1]
        back.R[:, :, [0, 2]] = -cameras.R[:, :, [0, 2]]
        back.T = -torch.bmm(back.R.transpose(1, 2), -cameras.get_camera_center()[:, :, None])[:, :, 0]
        occluded = visible_points_fn(model._points, back, depth_merge_threshold=thr)

        t0_parts, t1_parts = [], []
        for b in range(B):
            rays_b = F.normalize(targets[b] - cam_pos[b].view(1, 3), dim=-1)
            t0_sq, t1_sq = insurface_segments(cam_pos[b], rays_b, frontal.points_list()[b], occluded.points_list()[b])
            ok = (t0_sq < t1_sq).view(-1)
            ins_mask[b][ins_mask[b].clone()] = ok
            t1_parts.append(_eps_sqrt(t1_sq[ok]).sqrt())
            t0_parts.append(_eps_sqrt(t0_sq[ok]).sqrt())
        num_ins = ins_mask.sum(dim=-1)
        t0, t1 = torch.cat(t0_parts, dim=0), torch.cat(t1_parts, dim=0)            # (P,)
        cam_pos_ins = cam_pos[torch.repeat_interleave(torch.arange(B, device=num_ins.device), num_ins)]
        ray_ins = F.normalize(sample_points[ins_mask] - cam_pos_ins)

        steps = torch.linspace(0, 1.0, n_points_per_ray + 2, device=lengths.device)[1:-1]
        t_cand = steps * (t1 - t0).view(-1, 1) + t0.view(-1, 1)                     # (P,n)
        cand = t_cand.unsqueeze(-1) * ray_ins.unsqueeze(-2) + cam_pos_ins.unsqueeze(-2)   # (P,n,3)
        sdf = siren.sdf_fn(model.decoder)
        flat = cand.view(-1, 3)
        vals = torch.cat([sdf(chunk) for chunk in torch.split(flat, model.max_points_per_pass, dim=0)], dim=0) \
            if flat.shape[0] else flat.new_zeros((0,))
        best = torch.argmin(vals.view(cand.shape[0], n_points_per_ray), dim=-1, keepdim=True)
        p_ins = torch.gather(cand, -2, best.unsqueeze(-1).expand(-1, -1, 3)).squeeze(-2)
    return p_off, p_ins, num_off, num_ins


def subsample_randomly(point_clouds, ratio, generator=None):
    """``PointClouds3D.subsample_randomly`` (DSS/core/cloud.py:260-283): keep ``int(ratio[b] * P_b)`` points of
    every cloud, chosen by a random permutation (``generator``: optional torch.Generator for reproducible
    tests; the reference uses the global CPU generator)."""
    n = len(point_clouds)
    if not torch.is_tensor(ratio):
        ratio = torch.full((n,), float(ratio))
    ratio = ratio.detach().float().cpu().clamp(max=1.0)
    if ratio.numel() != n:
        raise ValueError("subsample_randomly: %d ratios for %d clouds" % (ratio.numel(), n))
    if bool((ratio == 1.0).all()):
        return point_clouds.clone()
    pts, nrm, feat = point_clouds.points_list(), point_clouds.normals_list(), point_clouds.features_list()
    out_p, out_n, out_f = [], [], []
    for b, p in enumerate(pts):
        keep = torch.randperm(p.shape[0], generator=generator)[:int(ratio[b] * p.shape[0])].to(p.device)
        out_p.append(p[keep])
        if nrm is not None:
            out_n.append(nrm[b][keep])
        if feat is not None:
            out_f.append(feat[b][keep])
    return point_clouds.__class__(out_p, normals=out_n if nrm is not None else None,
                                  features=out_f if feat is not None else None)


def get_visible_iso_points(model, cameras, jitter=None, generator=None, **proj_kwargs):
    """``Model.get_visible_iso_points`` (DSS/models/combined_modeling.py:390-455): the iso-points of
    ``model._points`` that the cameras see, topped up / thinned to about ``model.max_iso_per_batch`` per view,
    jittered by +-0.025, re-projected onto the level set (``model.projection.project_points``, no resampling) and
    filtered for visibility once more.  Returns a Pointclouds with normals (one cloud per camera).

    ``model`` is duck-typed on ``_points``, ``max_iso_per_batch``, ``projection``, ``decoder``, ``device``,
    ``get_normals_from_grad`` (only called when the clouds carry no normals) and
    ``renderer.rasterizer.raster_settings.depth_merging_threshold``.  ``jitter`` (P,3) in [0,1) replaces the
    ``torch.rand_like`` draw of :441, ``generator`` seeds the random subsampling (reproducible tests).  Every
    stage runs on this package's kernels: three visibility splats, FRNN-based upsampling, the fused projection."""
    from .cloud import PointCloudsFilters
    from .ewa import get_visible_points
    from .levelset_sampling import _mask_padded_to_list
    from .point_processing import upsample
    from .structures import Pointclouds
    cap = model.max_iso_per_batch
    if cap == 0:
        return torch.zeros((1, 0, 3), device=model.device, dtype=torch.float)
    B = cameras.R.shape[0]
    thr = model.renderer.rasterizer.raster_settings.depth_merging_threshold
    if model._points.normals_packed() is None:          # the visibility test culls back faces
        model._points.update_normals_(model.get_normals_from_grad(model._points.points_packed(), requires_grad=False))
    ref_pcl = proj_kwargs.get("ref_pcl", None)
    if ref_pcl is not None:
        if len(ref_pcl) != 1:
            raise AssertionError("Currently support optimizing a single shape, with only one reference point cloud.")
        if ref_pcl.normals_packed() is None:
            ref_pcl.update_normals_(model.get_normals_from_grad(ref_pcl.points_packed(), requires_grad=False))
        _, seen = get_visible_points(ref_pcl, cameras, depth_merge_threshold=thr, return_mask=True)
        ref_pcl = PointCloudsFilters(device=ref_pcl.device, visibility=seen.any(dim=0, keepdim=True)).filter(ref_pcl)
        proj_kwargs["ref_pcl"] = ref_pcl

    start = model._points.clone()
    visible, seen = get_visible_points(start.extend(B), cameras, depth_merge_threshold=thr, return_mask=True)
    if cap > 0:       # between 0.75 cap and cap points per view (:426-440)
        hi, lo = cap, 0.75 * cap
        if ref_pcl is not None:
            lo, hi = int(0.8 * lo), int(0.8 * hi)
        counts = visible.num_points_per_cloud().tolist()
        per_view = []
        for b in range(B):
            if counts[b] > hi:
                per_view.append(subsample_randomly(visible[b], hi / float(counts[b]), generator).points_packed())
            elif counts[b] < lo:
                per_view.append(upsample(visible[b], hi).points_packed())
            else:
                per_view.append(visible[b].points_packed())
        visible = Pointclouds(per_view)
    else:
        visible = PointCloudsFilters(device=start.device, visibility=seen.any(dim=0, keepdim=True)).filter(start)

    packed = visible.points_packed()
    u = torch.rand_like(packed) if jitter is None else jitter.to(packed)
    visible.offset_(0.05 * (u - 0.5))
    res = model.projection.project_points(visible, model.decoder, skip_resampling=True,
                                          skip_upsampling=(ref_pcl is None), **proj_kwargs)
    iso = Pointclouds(list(_mask_padded_to_list(res["levelset_points"], res["mask"])),
                      normals=list(_mask_padded_to_list(res["levelset_normals"], res["mask"])))
    return get_visible_points(iso, cameras, depth_merge_threshold=thr)
