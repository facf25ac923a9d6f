"""IDR-style ray tracer: first intersection of camera rays with the zero level set of an SDF.

Mirrors ``RayTracing`` (DSS/models/levelset_sampling.py:810-1167; called at implicit_modeling.py:431) with the
same constructor arguments, ``forward(sdf, cam_loc, object_mask, ray_directions)`` signature and return value
``(points (B*P,3), network_object_mask (B*P,), dists (B*P,))``:

1. bounds: both intersections of every ray with the bounding sphere (``intersection_with_unit_sphere``,
   DSS/utils/__init__.py:484-544),
2. ``sphere_tracing`` (:920-1032): march from both ends, stepping back by half a step where the march crossed
   the surface,
3. ``ray_sampler`` (:1034-1112) for the rays that did not converge: ``n_steps`` samples per ray, first sign
   change, ``secant`` refinement (:1114-1133),
4. training only: ``minimal_sdf_points`` (:1135-1167) for rays that disagree with the ground-truth mask.

Where the time goes is the SDF: ~2 x (sphere_tracing_iters + 1) evaluations per ray plus ``n_steps`` (100) per
unfinished ray plus ``n_secant_steps``.  ``sdf`` is a callable as in the reference; pass
``isopoints_b200.siren.sdf_fn(decoder)`` to evaluate the reference's Siren decoder with the forward-only fused
tcgen05 kernel (``isob200_siren_sdf``: value only, half the tensor work of value + gradient).  The per-ray
bookkeeping between evaluations is masked PyTorch on flat per-ray tensors, like the reference's; with the fused
decoder the masked evaluations of the march run off device-side row counts (``sdf_fn.masked``), so the marching loop
is enqueued without a single read-back.
"""
import torch
from torch import nn

__all__ = ["RayTracing", "intersection_with_unit_sphere"]


def _eps_denom(x, eps=1e-17):
    """DSS/utils/mathHelper.py:14-18."""
    sgn = torch.where(x < 0, -torch.ones_like(x), torch.ones_like(x))
    return sgn * x.abs().clamp_min(eps)


def intersection_with_unit_sphere(cam_pos, cam_rays, radius=1.0):
    """DSS/utils/__init__.py:484-544.  cam_pos (N,1,3) or (N,*,3), cam_rays (N,*,3) unit directions.
    Returns the near and far intersection points (N,*,3) and the mask of rays that hit the sphere; rays that
    miss get the points where they cross the two planes orthogonal to the view axis that touch the sphere."""
    if cam_pos.ndim != cam_rays.ndim:
        cam_rays = cam_rays.view(cam_rays.shape[0], -1, 3)
        cam_pos = cam_pos.view(cam_pos.shape[0], 1, 3)
    along = (cam_pos * cam_rays).sum(dim=-1)                      # signed distance to the closest approach
    closest = cam_pos - along[..., None] * cam_rays
    dist = closest.norm(p=2, dim=-1)                              # ray to centre
    cam_dist = cam_pos.norm(dim=-1)
    hit = dist <= radius
    chord = torch.where(hit, 2 * torch.sqrt((radius ** 2 - dist ** 2).clamp_min(0)), torch.full_like(dist, 10.0))
    chord = torch.where(hit, 2 * torch.sqrt(torch.where(hit, radius ** 2 - dist ** 2, torch.ones_like(dist))), chord)
    cam_d = cam_dist.expand_as(dist)
    near_hit = torch.sqrt(torch.where(hit, cam_d ** 2 - dist ** 2, torch.ones_like(dist))) - chord / 2.0
    slope = _eps_denom(-along / cam_dist)
    near = torch.where(hit, near_hit, ((cam_dist - radius) / slope).expand_as(dist))
    p_near = near.unsqueeze(-1) * cam_rays + cam_pos
    p_far_hit = chord[..., None] * cam_rays + p_near
    far_miss = ((radius + cam_dist) / slope).expand_as(dist)
    p_far = torch.where(hit[..., None], p_far_hit, far_miss.unsqueeze(-1) * cam_rays + cam_pos)
    return p_near, p_far, hit


class RayTracing(nn.Module):
    def __init__(self, object_bounding_sphere=1.0, sdf_threshold=5.0e-5, line_search_step=0.5, line_step_iters=1,
                 sphere_tracing_iters=10, n_steps=100, n_secant_steps=8):
        super().__init__()
        self.object_bounding_sphere = object_bounding_sphere
        self.sdf_threshold = sdf_threshold
        self.sphere_tracing_iters = sphere_tracing_iters
        self.line_step_iters = line_step_iters
        self.line_search_step = line_search_step
        self.n_steps = n_steps
        self.n_secant_steps = n_secant_steps

    # ------------------------------------------------------------------------------------------------
    def forward(self, sdf, cam_loc, object_mask, ray_directions, minimal_sdf_steps=None):
        """cam_loc (B,3), object_mask (B*P,) bool, ray_directions (B,P,3) unit.  levelset_sampling.py:831-918.
        ``minimal_sdf_steps`` (n_steps,) in [0,1): the sample positions of the training-only minimal-SDF search
        (the reference draws them with uniform_(); injectable for reproducible tests)."""
        if not ray_directions.is_cuda:      # the reference allocates with .cuda() throughout
            raise TypeError("RayTracing: ray_directions must be a cuda tensor")
        return self._trace(sdf, cam_loc, object_mask, ray_directions, minimal_sdf_steps)

    def _trace(self, sdf, cam_loc, object_mask, ray_directions, minimal_sdf_steps=None):
        """The device-agnostic host sequence behind ``forward`` (the CPU tests of the host logic enter here)."""
        B, P, _ = ray_directions.shape
        near_pt, far_pt, hit = intersection_with_unit_sphere(cam_loc, ray_directions, radius=self.object_bounding_sphere)
        origin = cam_loc.view(B, 1, 3)
        dir_len = ray_directions.norm(dim=-1)
        t_near = (near_pt - origin).norm(dim=-1) / dir_len              # (B,P) ray lengths at the two bounds
        t_far = (far_pt - origin).norm(dim=-1) / dir_len
        rays = _Rays(cam_loc, ray_directions)

        start_pts, unfinished, t_start, t_end, t_min, t_max = self.sphere_tracing(
            sdf, rays, hit.reshape(-1), t_near.reshape(-1), t_far.reshape(-1))
        net_mask = t_start < t_end

        # the rays that did not converge go to the sampler
        sampler_mask = unfinished
        if bool(sampler_mask.any()):
            lo = torch.where(sampler_mask, t_start, torch.zeros_like(t_start))
            hi = torch.where(sampler_mask, t_end, torch.zeros_like(t_end))
            s_pts, s_net_mask, s_t = self.ray_sampler(sdf, rays, object_mask, lo, hi, sampler_mask)
            start_pts = torch.where(sampler_mask[:, None], s_pts, start_pts)
            t_start = torch.where(sampler_mask, s_t, t_start)
            net_mask = torch.where(sampler_mask, s_net_mask, net_mask)

        if not self.training:
            return start_pts, net_mask, t_start

        hit_flat = hit.reshape(-1)
        in_mask = ~net_mask & object_mask & ~sampler_mask      # in the ground-truth mask, no surface found
        out_mask = ~object_mask & ~sampler_mask                # outside the ground-truth mask
        left_out = (in_mask | out_mask) & ~hit_flat
        if bool(left_out.any()):     # rays that miss the sphere: the point closest to the origin
            t_close = -(rays.d * rays.o).sum(-1)
            t_start = torch.where(left_out, t_close, t_start)
            start_pts = torch.where(left_out[:, None], rays.at(t_close), start_pts)
        mask = (in_mask | out_mask) & hit_flat
        if bool(mask.any()):
            sel = net_mask & out_mask
            t_min = torch.where(sel, t_start, t_min)
            m_pts, m_t = self.minimal_sdf_points(sdf, rays, mask, t_min, t_max, steps=minimal_sdf_steps)
            idx = torch.nonzero(mask).reshape(-1)
            start_pts = start_pts.index_copy(0, idx, m_pts)
            t_start = t_start.index_copy(0, idx, m_t)
        return start_pts, net_mask, t_start

    # ------------------------------------------------------------------------------------------------
    def sphere_tracing(self, sdf, rays, hit, t_near, t_far):
        """levelset_sampling.py:920-1032 on flat per-ray tensors.  Returns the front points, the mask of rays
        still unfinished at the front, the front / back ray lengths and the initial bounds."""
        R = hit.shape[0]
        zero = torch.zeros(R, dtype=torch.float32, device=hit.device)
        un_s, un_e = hit.clone(), hit.clone()
        t_s = torch.where(hit, t_near, zero)
        t_e = torch.where(hit, t_far, zero)
        p_s = torch.where(hit[:, None], rays.at(t_near), zero[:, None].expand(R, 3))
        p_e = torch.where(hit[:, None], rays.at(t_far), zero[:, None].expand(R, 3))
        t_min, t_max = t_s.clone(), t_e.clone()
        nxt_s = _masked_eval(sdf, p_s, un_s)
        nxt_e = _masked_eval(sdf, p_e, un_e)
        # With the fused decoder the masked evaluations read their row counts from device memory, so the loop is
        # enqueued without a read-back: the two ``.any()`` exits below are skipped -- an iteration over an empty
        # active set changes nothing (every update is masked), it only costs empty launches.
        nosync = hit.is_cuda and getattr(sdf, "masked", None) is not None and getattr(sdf, "fused", lambda: False)()
        iters = 0
        while True:
            cur_s = torch.where(un_s, nxt_s, zero)
            cur_s = torch.where(cur_s <= self.sdf_threshold, zero, cur_s)
            cur_e = torch.where(un_e, nxt_e, zero)
            cur_e = torch.where(cur_e <= self.sdf_threshold, zero, cur_e)
            un_s = un_s & (cur_s > self.sdf_threshold)
            un_e = un_e & (cur_e > self.sdf_threshold)
            if iters == self.sphere_tracing_iters or (not nosync and not bool((un_s | un_e).any())):
                break
            iters += 1
            t_s = t_s + cur_s
            t_e = t_e - cur_e
            p_s, p_e = rays.at(t_s), rays.at(t_e)
            nxt_s = _masked_eval(sdf, p_s, un_s)
            nxt_e = _masked_eval(sdf, p_e, un_e)
            # step back where the march went through the surface (:997-1023)
            bad_s, bad_e = nxt_s < 0, nxt_e < 0
            k = 0
            while k < self.line_step_iters and (nosync or bool((bad_s | bad_e).any())):
                back = (1 - self.line_search_step) / (2 ** k)
                t_s = torch.where(bad_s, t_s - back * cur_s, t_s)
                p_s = torch.where(bad_s[:, None], rays.at(t_s), p_s)
                t_e = torch.where(bad_e, t_e + back * cur_e, t_e)
                p_e = torch.where(bad_e[:, None], rays.at(t_e), p_e)
                nxt_s = torch.where(bad_s, _masked_eval(sdf, p_s, bad_s), nxt_s)
                nxt_e = torch.where(bad_e, _masked_eval(sdf, p_e, bad_e), nxt_e)
                bad_s, bad_e = nxt_s < 0, nxt_e < 0
                k += 1
            un_s = un_s & (t_s < t_e)
            un_e = un_e & (t_s < t_e)
        return p_s, un_s, t_s, t_e, t_min, t_max

    # ------------------------------------------------------------------------------------------------
    def ray_sampler(self, sdf, rays, object_mask, t_lo, t_hi, sampler_mask):
        """levelset_sampling.py:1034-1112: n_steps samples on [t_lo, t_hi] of every ray in ``sampler_mask``,
        first sample with a negative SDF, secant refinement between it and its predecessor."""
        R = sampler_mask.shape[0]
        dev = sampler_mask.device
        n = self.n_steps
        out_pts = torch.zeros(R, 3, dtype=torch.float32, device=dev)
        out_t = torch.zeros(R, dtype=torch.float32, device=dev)
        idx = torch.nonzero(sampler_mask).reshape(-1)
        frac = torch.linspace(0, 1, steps=n, device=dev).view(1, n)
        t_all = t_lo[idx, None] + frac * (t_hi[idx] - t_lo[idx])[:, None]            # (S,n)
        o, d = rays.o[idx], rays.d[idx]
        pts = o[:, None, :] + t_all[..., None] * d[:, None, :]                       # (S,n,3)
        val = _eval_chunked(sdf, pts.reshape(-1, 3), 80000).reshape(-1, n)
        # first index with the smallest sign: weights n..1 make argmin return the first minimum (:1062-1064)
        order = torch.sign(val) * torch.arange(n, 0, -1, device=dev, dtype=torch.float32).view(1, n)
        first = torch.argmin(order, -1)
        rows = torch.arange(idx.shape[0], device=dev)
        out_pts[idx] = pts[rows, first]
        out_t[idx] = t_all[rows, first]
        in_gt = object_mask[idx]
        crossed = val[rows, first] < 0
        p_out = ~(in_gt & crossed)              # no surface for this pixel: take the sample with minimal SDF
        if bool(p_out.any()):
            j = torch.argmin(val[p_out], -1)
            r2 = torch.arange(j.shape[0], device=dev)
            out_pts[idx[p_out]] = pts[p_out][r2, j]
            out_t[idx[p_out]] = t_all[p_out][r2, j]
        net_mask = sampler_mask.clone()
        net_mask[idx[~crossed]] = False
        sec = (crossed & in_gt) if self.training else crossed
        if bool(sec.any()):
            f = first[sec]
            r3 = torch.arange(f.shape[0], device=dev)
            z_hi, s_hi = t_all[sec][r3, f], val[sec][r3, f]
            z_lo, s_lo = t_all[sec][r3, f - 1], val[sec][r3, f - 1]
            z = self.secant(s_lo, s_hi, z_lo, z_hi, o[sec], d[sec], sdf)
            out_pts[idx[sec]] = o[sec] + z.unsqueeze(-1) * d[sec]
            out_t[idx[sec]] = z
        return out_pts, net_mask, out_t

    def secant(self, sdf_low, sdf_high, z_low, z_high, cam_loc, ray_directions, sdf):
        """levelset_sampling.py:1114-1133."""
        z_low, z_high, sdf_low, sdf_high = z_low.clone(), z_high.clone(), sdf_low.clone(), sdf_high.clone()
        z = -sdf_low * (z_high - z_low) / (sdf_high - sdf_low) + z_low
        for _ in range(self.n_secant_steps):
            mid = sdf(cam_loc + z.unsqueeze(-1) * ray_directions)
            pos, neg = mid > 0, mid < 0
            z_low = torch.where(pos, z, z_low)
            sdf_low = torch.where(pos, mid, sdf_low)
            z_high = torch.where(neg, z, z_high)
            sdf_high = torch.where(neg, mid, sdf_high)
            z = -sdf_low * (z_high - z_low) / (sdf_high - sdf_low) + z_low
        return z

    def minimal_sdf_points(self, sdf, rays, mask, t_min, t_max, steps=None):
        """levelset_sampling.py:1135-1167: the sample with the smallest SDF among n_steps random positions on
        [t_min, t_max] of the masked rays (one set of positions shared by all rays)."""
        n = self.n_steps
        dev = mask.device
        if steps is None:
            steps = torch.empty(n, device=dev).uniform_(0.0, 1.0)
        idx = torch.nonzero(mask).reshape(-1)
        lo, hi = t_min[idx, None], t_max[idx, None]
        t = steps.to(dev).view(1, n) * (hi - lo) + lo
        pts = rays.o[idx][:, None, :] + t[..., None] * rays.d[idx][:, None, :]
        val = _eval_chunked(sdf, pts.reshape(-1, 3), 100000).reshape(-1, n)
        j = val.argmin(-1)
        rows = torch.arange(idx.shape[0], device=dev)
        return pts[rows, j], t[rows, j]


class _Rays:
    """Flat per-ray origin / direction (R,3) of a (B,P) ray bundle."""

    def __init__(self, cam_loc, ray_directions):
        B, P, _ = ray_directions.shape
        self.o = cam_loc.view(B, 1, 3).expand(B, P, 3).reshape(-1, 3)
        self.d = ray_directions.reshape(-1, 3)

    def at(self, t):
        return self.o + t.unsqueeze(-1) * self.d


def _eval_chunked(sdf, points, chunk):
    """sdf over (n,3) points in the reference's chunks (80 000 / 100 000 rows, :1056, :1158) -- or in one call when
    the callable says any size is fine (``sdf.any_size``, set by ``siren.sdf_fn`` for the fused kernel: one launch
    over full waves of tiles instead of ~4.2-wave launches per chunk)."""
    if getattr(sdf, "any_size", False) and getattr(sdf, "fused", lambda: False)():
        return sdf(points).reshape(-1)
    return torch.cat([sdf(part).reshape(-1) for part in torch.split(points, chunk, dim=0)])


def _masked_eval(sdf, points, mask):
    """zeros(R) with sdf(points[mask]) at the masked rows (the reference's ``x[mask] = sdf(p[mask])``)."""
    fast = getattr(sdf, "masked", None)
    if fast is not None and points.is_cuda:
        out = fast(points, mask)
        if out is not None:
            return out
    out = torch.zeros(points.shape[0], dtype=torch.float32, device=points.device)
    idx = torch.nonzero(mask).reshape(-1)
    if idx.numel():
        out[idx] = sdf(points[idx]).reshape(-1).to(out.dtype)
    return out
