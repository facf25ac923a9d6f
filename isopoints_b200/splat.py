"""Differentiable elliptical-splat rasteriser on libisob200.so.

Mirrors the operator surface of DSS/core/rasterizer.py + DSS/csrc (``DSS._C``) + the blend of
DSS/core/renderer.py:

* ``_C.splat_points(points, ellipse_params, cutoff_thres, radii, cloud_to_packed_first_idx,
  num_points_per_cloud, depth_merging_thres, image_size, points_per_pixel, bin_size,
  max_points_per_bin) -> (idx i32 (N,S,S,K), zbuf, qvalue f32 (N,S,S,K), occ f32 (N,S,S))``
  (rasterize_points.h:461-525), ``_C._splat_points_naive``, ``_C._backward_zbuf`` (in place),
  ``_C._splat_points_occ_backward`` and ``_C._splat_points_occ_fast_cuda_backward``;
* ``rasterize_elliptical_points(...)`` (rasterizer.py:678-740) and the autograd Function
  ``EllipticalRasterizer`` (rasterizer.py:743-973): gradients only for ``pts_screen`` -- xy from
  ``occ_grad``, z from ``zbuf_grad``; ``qvalue_grad`` is ignored exactly like the reference;
* ``blend_rgba`` = ``SurfaceSplattingRenderer.forward``'s weights + NormWeightedCompositor +
  alpha (renderer.py:53-78) in one kernel, differentiable w.r.t. the point features.

``bin_size`` / ``max_points_per_bin`` are accepted for signature compatibility; they only select
the occupancy rule (bin_size == 0 is the reference's naive kernel: occupied when z >= 0, otherwise
the fine kernel's z > 0, rasterize_points.cu:196 vs :581) and the reference's ``num_bins < 22``
argument check.  The (N,B,B,M) ``bin_points`` matrix is not materialised.
There is no CPU / PyTorch fallback: CPU tensors raise.
"""
import os
from typing import NamedTuple, Optional

import torch
from torch import autograd

from . import _ext

# raster kernel variant: 0 = v2 (record-centric hit generation, default), 2 = v2 with two CTAs per SM, 1 = v1 (pixel-centric with
# warp-level culling).  Identical results; v1 is kept as the fallback algorithm inside v2 and for A/B timing.
RASTER_VARIANT = int(os.environ.get("ISOB200_RASTER_VARIANT", "0"))

kMaxPointsPerBin = 22
kMaxPointsPerPixel = 150
NORM_WEIGHT_EPS = 1e-4   # pytorch3d norm_weighted_sum kEpsilon [third party, restated]


class PointFragments(NamedTuple):
    """DSS/core/rasterizer.py:31-36."""
    idx: torch.Tensor
    zbuf: torch.Tensor
    qvalue: torch.Tensor
    scaler: torch.Tensor
    occupancy: torch.Tensor


def _f32c(t, name):
    if not t.is_cuda:
        raise TypeError("%s: for now only cuda version is supported" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s: expected scalar type Float" % name)
    return t.contiguous()


def _i64c(t, dev):
    return t.to(device=dev, dtype=torch.int64).contiguous()


def _check_inputs(points, ellipse, cutoff, radii, first_idx, num_points):
    """Shape checks of rasterize_points.h:470-486 (torch::checkDim / checkSize -> RuntimeError)."""
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("Expected 2-dimensional tensor of size (P, 3) for points, got %s" % (tuple(points.shape),))
    P = points.shape[0]
    if ellipse.dim() != 2 or tuple(ellipse.shape) != (P, 3):
        raise RuntimeError("Expected ellipse_params of size (%d, 3), got %s" % (P, tuple(ellipse.shape)))
    if radii.dim() != 2 or tuple(radii.shape) != (P, 2):
        raise RuntimeError("Expected radii of size (%d, 2), got %s" % (P, tuple(radii.shape)))
    if cutoff.dim() != 1 or cutoff.shape[0] not in (1, P):
        raise RuntimeError("Expected cutoff_thres of size (%d,) or (1,), got %s" % (P, tuple(cutoff.shape)))
    if first_idx.dim() != 1 or num_points.dim() != 1 or first_idx.shape != num_points.shape:
        raise RuntimeError("cloud_to_packed_first_idx and num_points_per_cloud must be (N,) tensors of equal size")


def _fusable(K):
    """The raster kernel's fused epilogue (blend / visibility) exists for the default variant and K <= 16."""
    return int(K) <= 16 and (RASTER_VARIANT & 3) != 1


def _splat(points, ellipse, cutoff, radii, first_idx, num_points, depth_merging_thres, S, K, occ_inclusive,
           blend=None, want_visible=False):
    """idx, zbuf, qvalue, occ of DSS._C.splat_points.  ``blend`` = (scaler (P,) | None, feat (P,C), eps,
    want_weights): also the RGBA image (N,S,S,C+1) [and the blend weights (N,S,S,K)] from the raster kernel's
    epilogue; ``want_visible``: also the per-point visibility (P,) uint8 of rasterizer.py:851-857.  With either,
    returns (idx, zbuf, qvalue, occ, img, weights, visible) (None where not asked for)."""
    _check_inputs(points, ellipse, cutoff, radii, first_idx, num_points)
    lib = _ext.lib()
    S, K = int(S), int(K)
    if K > kMaxPointsPerPixel:
        raise RuntimeError("Must have points_per_pixel <= %d" % kMaxPointsPerPixel)
    points = _f32c(points, "points")
    dev = points.device
    ellipse = _f32c(ellipse, "ellipse_params")
    radii = _f32c(radii, "radii")
    P = points.shape[0]
    cutoff = _f32c(cutoff.expand(P) if cutoff.shape[0] == 1 and P != 1 else cutoff, "cutoff_thres")
    first_idx = _i64c(first_idx, dev)
    num_points = _i64c(num_points, dev)
    N = num_points.shape[0]
    extra = blend is not None or want_visible
    idx = torch.empty((N, S, S, K), dtype=torch.int32, device=dev)
    zbuf = torch.empty((N, S, S, K), dtype=torch.float32, device=dev)
    qvalue = torch.empty((N, S, S, K), dtype=torch.float32, device=dev)
    occ = torch.empty((N, S, S), dtype=torch.float32, device=dev)
    img = weights = vis = scaler = feat = None
    C, eps = 0, 0.0
    if blend is not None:
        scaler, feat, eps, want_w = blend
        C = feat.shape[1]
        img = torch.empty((N, S, S, C + 1), dtype=torch.float32, device=dev)
        weights = torch.empty((N, S, S, K), dtype=torch.float32, device=dev) if want_w else None
    if want_visible:
        vis = torch.zeros((P,), dtype=torch.uint8, device=dev)
    if idx.numel() == 0:
        return (idx, zbuf, qvalue, occ, img, weights, vis) if extra else (idx, zbuf, qvalue, occ)
    st = _ext.stream(dev)
    ws = _ext.workspace(lib.isob200_splat_ws_bytes(N, S), dev)
    total = torch.empty((1,), dtype=torch.int32, device=dev)
    # no view can own more than P points: P is a launch-shape bound that needs no host sync
    maxp = P
    _ext.check(lib.isob200_splat_bin(_ext.ptr(points), _ext.ptr(radii), _ext.ptr(first_idx), _ext.ptr(num_points),
                                     N, P, maxp, S, _ext.ptr(ws), ws.numel(), _ext.ptr(total), st))
    cap = int(total.item())          # the one read-back: sizes the per-tile record buffer
    recs = torch.empty((max(cap, 1) * lib.isob200_splat_record_bytes(),), dtype=torch.uint8, device=dev)
    flags = (1 if occ_inclusive else 0) | ((RASTER_VARIANT & 3) << 8)
    if not extra:
        _ext.check(lib.isob200_splat_forward(
            _ext.ptr(points), _ext.ptr(ellipse), _ext.ptr(cutoff), _ext.ptr(radii), _ext.ptr(first_idx),
            _ext.ptr(num_points), N, P, maxp, S, K, float(depth_merging_thres), flags,
            _ext.ptr(ws), ws.numel(), _ext.ptr(recs), cap, _ext.ptr(idx), _ext.ptr(zbuf), _ext.ptr(qvalue),
            _ext.ptr(occ), st))
        return idx, zbuf, qvalue, occ
    _ext.check(lib.isob200_splat_forward_fused(
        _ext.ptr(points), _ext.ptr(ellipse), _ext.ptr(cutoff), _ext.ptr(radii), _ext.ptr(first_idx),
        _ext.ptr(num_points), N, P, maxp, S, K, float(depth_merging_thres), flags,
        _ext.ptr(ws), ws.numel(), _ext.ptr(recs), cap, _ext.ptr(idx), _ext.ptr(zbuf), _ext.ptr(qvalue),
        _ext.ptr(occ), _ext.ptr(scaler), _ext.ptr(feat), 0 if feat is None else feat.stride(0), C, float(eps),
        _ext.ptr(img), _ext.ptr(weights), _ext.ptr(vis), st))
    return idx, zbuf, qvalue, occ, img, weights, vis


def bin_counts(points, radii, first_idx, num_points, image_size, bin_size):
    """``points_per_bin`` (N,B,B) int32 of the reference's coarse pass (rasterize_points.cu:353-412),
    which it computes and drops; exposed for the per-tile-count parity check."""
    lib = _ext.lib()
    points = _f32c(points, "points")
    radii = _f32c(radii, "radii")
    dev = points.device
    first_idx = _i64c(first_idx, dev)
    num_points = _i64c(num_points, dev)
    N = num_points.shape[0]
    B = 1 + (int(image_size) - 1) // int(bin_size)
    out = torch.empty((N, B, B), dtype=torch.int32, device=dev)
    _ext.check(lib.isob200_splat_bin_counts(_ext.ptr(points), _ext.ptr(radii), _ext.ptr(first_idx),
                                            _ext.ptr(num_points), N, points.shape[0], int(image_size), int(bin_size),
                                            _ext.ptr(out), _ext.stream(dev)))
    return out


def count_pixel_splats(points, ellipse, cutoff, radii, image_size):
    """Number of (pixel, point) pairs passing CheckPixelInsidePoint (rasterize_points.cu:64-98): the
    "pixel-splat" unit of the throughput metric.  Returns a python int (synchronises)."""
    points = _f32c(points, "points")
    P = points.shape[0]
    cutoff = _f32c(cutoff.expand(P) if cutoff.shape[0] == 1 and P != 1 else cutoff, "cutoff_thres")
    total = torch.zeros((1,), dtype=torch.int64, device=points.device)
    _ext.check(_ext.lib().isob200_splat_count_pairs(
        _ext.ptr(points), _ext.ptr(_f32c(ellipse, "ellipse")), _ext.ptr(cutoff), _ext.ptr(_f32c(radii, "radii")), P,
        int(image_size), _ext.ptr(total), _ext.stream(points.device)))
    return int(total.item())


def visibility_mask(idx, num_points_total, occupancy=None):
    """(P,) bool: ids in any slot of an active pixel -- active = ``occupancy != 0`` when given
    (get_per_point_visibility_mask, DSS/utils/__init__.py:378-399) else ``idx[..., 0] >= 0``
    (EllipticalRasterizer.backward, rasterizer.py:851-857).  Replaces unique() + scatter."""
    if not idx.is_cuda:
        raise TypeError("for now only cuda version is supported")
    idx = idx.contiguous()
    K = idx.shape[-1]
    vis = torch.zeros((int(num_points_total),), dtype=torch.uint8, device=idx.device)
    mask = None if occupancy is None else _f32c(occupancy, "occupancy")
    _ext.check(_ext.lib().isob200_splat_visibility(_ext.ptr(idx), _ext.ptr(mask), idx.numel() // max(K, 1), K,
                                                   int(num_points_total), _ext.ptr(vis), _ext.stream(idx.device)))
    return vis.bool()


# occupancy backward sweep: True = hybrid (per-tile lists of the non-zero gradient pixels, dense fallback
# per tile), False = plain window sweep.  Same sums up to fp32 summation order.
OCC_BACKWARD_HYBRID = True


def _occ_backward(points, radii, visible_u8, first_idx, num_points, rs, radii_s, grad_occ, mode, out, out_stride):
    N, H, W = grad_occ.shape
    lib = _ext.lib()
    ws = _ext.workspace(lib.isob200_splat_occ_backward_ws_bytes(N, H, W, points.shape[0]), points.device) if OCC_BACKWARD_HYBRID else None
    _ext.check(lib.isob200_splat_occ_backward(
        _ext.ptr(points), _ext.ptr(radii), _ext.ptr(visible_u8), _ext.ptr(first_idx), _ext.ptr(num_points),
        _ext.ptr(rs), float(radii_s), _ext.ptr(grad_occ), N, H, W, points.shape[0], mode, _ext.ptr(out),
        out_stride, _ext.ptr(ws), 0 if ws is None else ws.numel(), _ext.stream(points.device)))


class _CExt:
    """Drop-in for the bound subset of ``DSS._C`` (DSS/csrc/ext.cpp:5-18)."""

    @staticmethod
    def splat_points(points, ellipse_params, cutoff_thres, radii, cloud_to_packed_first_idx,
                     num_points_per_cloud, depth_merging_thres, image_size, points_per_pixel, bin_size,
                     max_points_per_bin):
        if bin_size != 0:
            num_bins = 1 + (int(image_size) - 1) // int(bin_size)
            if num_bins >= kMaxPointsPerBin:   # rasterize_points.cu:462-468
                raise RuntimeError("Got %d; that's too many!" % num_bins)
        return _splat(points, ellipse_params, cutoff_thres, radii, cloud_to_packed_first_idx,
                      num_points_per_cloud, depth_merging_thres, image_size, points_per_pixel,
                      occ_inclusive=(bin_size == 0))

    @staticmethod
    def _splat_points_naive(points, ellipse_params, cutoff_thres, radii, cloud_to_packed_first_idx,
                            num_points_per_cloud, depth_merging_thres, image_size, points_per_pixel):
        return _splat(points, ellipse_params, cutoff_thres, radii, cloud_to_packed_first_idx,
                      num_points_per_cloud, depth_merging_thres, image_size, points_per_pixel, occ_inclusive=True)

    @staticmethod
    def _backward_zbuf(idx, zbuf_grad, point_z_grad):
        """In place: point_z_grad (P,1) += scatter of zbuf_grad by idx (rasterize_points.h:388-419)."""
        if not idx.is_cuda:
            raise TypeError("for now only cuda version is supported")
        N, H, W, K = idx.shape
        if not point_z_grad.is_contiguous():
            raise RuntimeError("_backward_zbuf: point_z_grad must be contiguous")
        _ext.check(_ext.lib().isob200_splat_zbuf_backward(
            _ext.ptr(idx.contiguous()), _ext.ptr(_f32c(zbuf_grad, "zbuf_grad")), N, H, W, K,
            _ext.ptr(point_z_grad), point_z_grad.stride(0) if point_z_grad.dim() > 1 else 1,
            _ext.stream(idx.device)))

    @staticmethod
    def _splat_points_occ_backward(points, radii, grad_occ, cloud_to_packed_first_idx, num_points_per_cloud,
                                   radii_s, depth_merging_thres):
        """Slow-path occupancy backward (rasterize_points.h:341-386) -> (P,2)."""
        points = _f32c(points, "points")
        radii = _f32c(radii, "radii")
        out = torch.empty((points.shape[0], 2), dtype=torch.float32, device=points.device)
        if out.numel():
            _occ_backward(points, radii, None, _i64c(cloud_to_packed_first_idx, points.device),
                          _i64c(num_points_per_cloud, points.device), None, radii_s, _f32c(grad_occ, "grad_occ"),
                          1, out, 2)
        return out

    @staticmethod
    def _splat_points_occ_fast_cuda_backward(points_sorted, radii_sorted, rs, grad_occ, num_points_per_cloud,
                                             cloud_to_packed_first_idx, points_grid_off=None, grid_params=None):
        """Fast-path occupancy backward (rasterize_points.h:327-336) -> (P,2) in the order of the given
        points.  The 2-D grid arguments of the reference are accepted and ignored: the kernel is
        point-centric and needs no neighbour structure."""
        points = _f32c(points_sorted, "points")
        radii = _f32c(radii_sorted, "radii")
        out = torch.empty((points.shape[0], 2), dtype=torch.float32, device=points.device)
        if out.numel():
            _occ_backward(points, radii, None, _i64c(cloud_to_packed_first_idx, points.device),
                          _i64c(num_points_per_cloud, points.device), _f32c(rs, "rs"), 0.0,
                          _f32c(grad_occ, "grad_occ"), 0, out, 2)
        return out


_C = _CExt()


def per_view_search_radius(radii, visible, first_idx, num_points, radii_s=1.0):
    """(N,) median over the flattened (n_visible, 2) radii of every view, times ``radii_s``
    (rasterizer.py:881-884; torch.median = lower middle): exact radix select on the device, no sort and
    no host synchronisation.  Views with no visible point get 0."""
    lib = _ext.lib()
    radii = _f32c(radii, "radii")
    dev = radii.device
    N = num_points.shape[0]
    rs = torch.empty((N,), dtype=torch.float32, device=dev)
    ws = _ext.workspace(lib.isob200_splat_search_radius_ws_bytes(N), dev)
    vis = None if visible is None else visible.view(torch.uint8)
    _ext.check(lib.isob200_splat_search_radius(_ext.ptr(radii), _ext.ptr(vis), _ext.ptr(_i64c(first_idx, dev)),
                                               _ext.ptr(_i64c(num_points, dev)), N, radii.shape[0], float(radii_s),
                                               _ext.ptr(rs), _ext.ptr(ws), ws.numel(), _ext.stream(dev)))
    return rs


def per_view_median_radius(radii, visible, first_idx, num_points):
    return per_view_search_radius(radii, visible, first_idx, num_points, 1.0)


class EllipticalRasterizer(autograd.Function):
    """DSS/core/rasterizer.py:743-973."""

    @staticmethod
    def forward(ctx, pts_screen, ellipse_param, cutoff_threshold, radii, cloud_to_packed_first_idx,
                num_points_per_cloud, depth_merging_threshold, image_size, points_per_pixel, bin_size=0,
                max_points_per_bin=0, radii_backward_scaler=10.0):
        vis = None
        if ctx.needs_input_grad[0] and _fusable(points_per_pixel) and pts_screen.shape[0] > 0:
            # a backward will follow: its per-point visibility (rasterizer.py:851-857) comes out of the raster
            # kernel's epilogue instead of a second pass over idx
            _check_bins(image_size, bin_size)
            idx, zbuf, qvalue_map, occ_map, _, _, vis = _splat(
                pts_screen, ellipse_param, cutoff_threshold, radii, cloud_to_packed_first_idx, num_points_per_cloud,
                depth_merging_threshold, image_size, points_per_pixel, occ_inclusive=(bin_size == 0),
                want_visible=True)
        else:
            idx, zbuf, qvalue_map, occ_map = _C.splat_points(
                pts_screen, ellipse_param, cutoff_threshold, radii, cloud_to_packed_first_idx, num_points_per_cloud,
                depth_merging_threshold, image_size, points_per_pixel, bin_size, max_points_per_bin)
        ctx.radii_backward_scaler = radii_backward_scaler
        ctx.depth_merging_threshold = depth_merging_threshold
        ctx.has_vis = vis is not None
        saved = (pts_screen, radii, idx, cloud_to_packed_first_idx, num_points_per_cloud)
        ctx.save_for_backward(*(saved + ((vis,) if vis is not None else ())))
        ctx.mark_non_differentiable(idx)
        return idx, zbuf, qvalue_map, occ_map

    @staticmethod
    def backward(ctx, idx_grad, zbuf_grad, qvalue_grad, occ_grad):
        saved = ctx.saved_tensors
        pts_screen, radii, idx, first_idx, num_points = saved[:5]
        grads = _points_backward(pts_screen, radii, idx, first_idx, num_points, ctx.radii_backward_scaler,
                                 saved[5].view(torch.bool) if ctx.has_vis else None, zbuf_grad, occ_grad)
        return (grads,) + (None,) * 11


def _check_bins(image_size, bin_size):
    if bin_size != 0:
        num_bins = 1 + (int(image_size) - 1) // int(bin_size)
        if num_bins >= kMaxPointsPerBin:   # rasterize_points.cu:462-468
            raise RuntimeError("Got %d; that's too many!" % num_bins)


def _points_backward(pts_screen, radii, idx, first_idx, num_points, radii_s, vis, zbuf_grad, occ_grad):
    """EllipticalRasterizer.backward (rasterizer.py:784-973): (P,3) gradient of the screen-space points from the
    occupancy (xy) and depth (z) gradients.  ``vis``: per-point visibility when the forward already produced it."""
    if radii_s == 0:
        raise RuntimeError("radii_backward_scaler == 0 (WeightBackward) is not implemented in the reference "
                           "either (rasterizer.py:776-777 saves 4 tensors, :809-811 unpacks 8)")
    dev = pts_screen.device
    P = pts_screen.shape[0]
    pts = _f32c(pts_screen.detach(), "pts_screen")
    radii = _f32c(radii, "radii")
    first_idx = _i64c(first_idx, dev)
    num_points = _i64c(num_points, dev)
    grads = torch.zeros((P, 3), dtype=torch.float32, device=dev)
    if occ_grad is not None and P > 0:
        if vis is None:
            vis = visibility_mask(idx, P)                                # rasterizer.py:851-857
        rs = per_view_search_radius(radii, vis, first_idx, num_points, radii_s)             # :881-884
        _occ_backward(pts, radii, vis.view(torch.uint8), first_idx, num_points, rs, radii_s,
                      _f32c(occ_grad, "occ_grad"), 0, grads, 3)
    if zbuf_grad is not None and P > 0:
        N, H, W, K = idx.shape
        _ext.check(_ext.lib().isob200_splat_zbuf_backward(
            _ext.ptr(idx), _ext.ptr(_f32c(zbuf_grad, "zbuf_grad")), N, H, W, K,
            grads.data_ptr() + 8, 3, _ext.stream(dev)))                  # column 2 of (P,3)
    return grads


class SplatRender(autograd.Function):
    """EllipticalRasterizer + the renderer's RGBA blend (renderer.py:53-78) in ONE pass over the pixels: the blend
    runs in the raster kernel's epilogue.  Outputs (idx, zbuf, qvalue, occupancy, images); gradients: to the
    screen-space points through occupancy (incl. the image's alpha channel, which IS the occupancy) and depth as in
    EllipticalRasterizer.backward, to the features through the blend weights."""

    @staticmethod
    def forward(ctx, pts_screen, ellipse_param, cutoff_threshold, radii, first_idx, num_points,
                depth_merging_threshold, image_size, points_per_pixel, bin_size, radii_backward_scaler, scaler,
                features, eps):
        _check_bins(image_size, bin_size)
        feat = features if features.dtype == torch.float32 else features.float()
        if feat.stride(-1) != 1:
            feat = feat.contiguous()
        sc = None if scaler is None else _f32c(scaler.reshape(-1), "scaler")
        need_pts, need_feat = ctx.needs_input_grad[0], ctx.needs_input_grad[12]
        idx, zbuf, qvalue, occ, img, weights, vis = _splat(
            pts_screen, ellipse_param, cutoff_threshold, radii, first_idx, num_points, depth_merging_threshold,
            image_size, points_per_pixel, occ_inclusive=(bin_size == 0),
            blend=(sc, feat.detach(), eps, need_feat), want_visible=need_pts and pts_screen.shape[0] > 0)
        ctx.meta = (radii_backward_scaler, feat.shape, eps, vis is not None, weights is not None)
        extra = tuple(t for t in (vis, weights) if t is not None)
        ctx.save_for_backward(pts_screen, radii, idx, first_idx, num_points, *extra)
        ctx.mark_non_differentiable(idx)
        return idx, zbuf, qvalue, occ, img

    @staticmethod
    def backward(ctx, idx_grad, zbuf_grad, qvalue_grad, occ_grad, img_grad):
        radii_s, fshape, eps, has_vis, has_w = ctx.meta
        saved = ctx.saved_tensors
        pts_screen, radii, idx, first_idx, num_points = saved[:5]
        vis = saved[5].view(torch.bool) if has_vis else None
        weights = saved[5 + int(has_vis)] if has_w else None
        C = fshape[1]
        gpts = gfeat = None
        if ctx.needs_input_grad[0]:
            og = occ_grad
            if img_grad is not None:                 # alpha = occupancy (renderer.py:76-78)
                a = img_grad[..., C]
                og = a if og is None else og + a
            gpts = _points_backward(pts_screen, radii, idx, first_idx, num_points, radii_s, vis, zbuf_grad,
                                    None if og is None else og.contiguous())
        if ctx.needs_input_grad[12] and img_grad is not None and weights is not None:
            N, H, W, K = idx.shape
            gfeat = torch.zeros(fshape, dtype=torch.float32, device=idx.device)
            _ext.check(_ext.lib().isob200_splat_blend_backward(
                _ext.ptr(idx), _ext.ptr(weights), _ext.ptr(img_grad.contiguous()), N * H * W, K, C, float(eps),
                _ext.ptr(gfeat), gfeat.stride(0), _ext.stream(idx.device)))
        return (gpts,) + (None,) * 11 + (gfeat, None)


def rasterize_elliptical_points(pcls_screen, ellipse_params, cutoff_threshold, radii,
                                depth_merging_threshold: float = 0.05, image_size: int = 512,
                                points_per_pixel: int = 5, bin_size: Optional[int] = None,
                                max_points_per_bin: Optional[int] = None, radii_backward_scaler: float = 10.0,
                                clip_pts_grad: float = -1.0, blend=None):
    """DSS/core/rasterizer.py:678-740.  ``pcls_screen`` duck-types ``points_packed()``,
    ``cloud_to_packed_first_idx()``, ``num_points_per_cloud()``.  ``blend`` = (scaler (P,) | None, features (P,C),
    eps): also return the renderer's RGBA images (renderer.py:53-78) as a fifth output, computed in the raster
    kernel's epilogue (``SplatRender``) when that exists for the settings, by ``blend_rgba`` otherwise."""
    points_packed = pcls_screen.points_packed()
    cloud_to_packed_first_idx = pcls_screen.cloud_to_packed_first_idx()
    num_points_per_cloud = pcls_screen.num_points_per_cloud()
    cutoff_threshold = cutoff_threshold.expand(points_packed.shape[0])
    if bin_size is None:
        if image_size <= 64:
            bin_size = 8
        elif image_size <= 256:
            bin_size = 16
        elif image_size <= 512:
            bin_size = 32
        elif image_size <= 1024:
            bin_size = 64
    if bin_size != 0:
        num_bins = 1 + (image_size - 1) // bin_size
        if num_bins >= kMaxPointsPerBin:
            raise ValueError("bin_size too small, number of bins must be less than %d; got %d"
                             % (kMaxPointsPerBin, num_bins))
    if max_points_per_bin is None:
        max_points_per_bin = 0      # unused here: no (N,B,B,M) matrix, hence no .max() host sync
    if points_packed.requires_grad and clip_pts_grad > 0:
        def _clip(g, m=clip_pts_grad):
            norm = g.norm(dim=-1, keepdim=True)
            return torch.where(norm > m, g * (m / norm.clamp_min(1e-20)), g)
        points_packed.register_hook(_clip)
    if blend is not None and _fusable(points_per_pixel) and blend[1].shape[1] <= 4:
        return SplatRender.apply(points_packed, ellipse_params, cutoff_threshold, radii, cloud_to_packed_first_idx,
                                 num_points_per_cloud, depth_merging_threshold, image_size, points_per_pixel,
                                 bin_size, radii_backward_scaler, blend[0], blend[1], blend[2])
    out = EllipticalRasterizer.apply(points_packed, ellipse_params, cutoff_threshold, radii,
                                     cloud_to_packed_first_idx, num_points_per_cloud, depth_merging_threshold,
                                     image_size, points_per_pixel, bin_size, max_points_per_bin,
                                     radii_backward_scaler)
    if blend is not None:
        out = out + (blend_rgba(out[0], out[2], out[3], blend[0], blend[1], blend[2]),)
    return out


class _Blend(autograd.Function):
    @staticmethod
    def forward(ctx, idx, qvalue, occ, scaler, feat, eps):
        lib = _ext.lib()
        N, H, W, K = idx.shape
        C = feat.shape[1]
        dev = idx.device
        out = torch.empty((N, H, W, C + 1), dtype=torch.float32, device=dev)
        need_grad = feat.requires_grad
        weights = torch.empty((N, H, W, K), dtype=torch.float32, device=dev) if need_grad else None
        _ext.check(lib.isob200_splat_blend(_ext.ptr(idx), _ext.ptr(qvalue), _ext.ptr(occ), _ext.ptr(scaler),
                                           _ext.ptr(feat), feat.stride(0), N * H * W, K, C, float(eps),
                                           _ext.ptr(out), _ext.ptr(weights), _ext.stream(dev)))
        if need_grad:
            ctx.save_for_backward(idx, weights)
        ctx.meta = (feat.shape, eps)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        shape, eps = ctx.meta
        gfeat = None
        if ctx.needs_input_grad[4]:
            idx, weights = ctx.saved_tensors
            N, H, W, K = idx.shape
            gfeat = torch.zeros(shape, dtype=torch.float32, device=idx.device)
            _ext.check(_ext.lib().isob200_splat_blend_backward(
                _ext.ptr(idx), _ext.ptr(weights), _ext.ptr(grad_out.contiguous()), N * H * W, K, shape[1], float(eps),
                _ext.ptr(gfeat), gfeat.stride(0), _ext.stream(idx.device)))
        # alpha = occupancy is copied into the last channel (renderer.py:76-78: torch.cat with the mask), so its
        # gradient goes straight back to the rasteriser's occupancy output
        gocc = grad_out[..., shape[1]].contiguous() if ctx.needs_input_grad[2] else None
        return None, None, gocc, None, gfeat, None


def blend_rgba(idx, qvalue, occupancy, scaler, features, eps: float = NORM_WEIGHT_EPS):
    """(N,S,S,C+1) = [sum_k w_k f[idx_k] / max(sum_k w_k, eps), occupancy] with
    w_k = exp(-0.5 qvalue_k) * scaler[idx_k] over idx >= 0 (renderer.py:53-78).
    ``scaler``: per-POINT (P,) scaler (or None = 1); ``features``: (P, C<=4), e.g. rgb."""
    if not idx.is_cuda:
        raise TypeError("for now only cuda version is supported")
    feat = features if features.dtype == torch.float32 else features.float()
    if feat.stride(-1) != 1:
        feat = feat.contiguous()
    sc = None if scaler is None else _f32c(scaler.reshape(-1), "scaler")
    return _Blend.apply(idx.contiguous(), _f32c(qvalue, "qvalue"), _f32c(occupancy, "occupancy"), sc, feat, eps)


def gather_with_neg_idx(input: torch.Tensor, dim: int, index: torch.Tensor):
    """DSS/utils/__init__.py:172-185 (without mutating ``index``)."""
    mask = index >= 0
    out = torch.gather(input, dim, index.clamp_min(0))
    return out * mask.to(out.dtype)
