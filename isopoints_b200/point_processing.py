"""Point-set operators on libisob200.so: ``wlop``, ``upsample``, ``resample_uniformly``,
``farthest_sampling`` and an exact ``knn_points``.

Mirrors DSS/utils/point_processing.py (:35-122, :126-166, :281-362, :473-499) -- same
signatures, and results mirror the input type (``Pointclouds`` in -> ``pcl.__class__`` out,
tensor in -> (padded, num_points) out, :120-122, :291-296).  The neighbourhood math of every
iteration runs in one fused kernel (csrc/pointops.cu) on neighbour ids from the FRNN grid
(csrc/frnn_*.cu); the brute-force O(P^2) ``pytorch3d.ops.knn_points`` the reference calls for
``upsample`` is replaced by an exact K-nearest search on the same uniform grid (radius escalation
until every query has K hits -- a complete list inside radius r IS the K nearest).
No CPU / PyTorch fallback for the kernels.
"""
import math
from collections import namedtuple
from typing import Tuple, Union

import torch
import torch.nn.functional as F

from . import _ext
from . import frnn
from .structures import convert_pointclouds_to_tensor, is_pointclouds, list_to_padded

_KNN = namedtuple("KNN", "dists idx knn")   # pytorch3d.ops.knn._KNN


def padded_to_list(x, split_size):
    return [x[i, : split_size[i]] for i in range(x.shape[0])]


def knn_points(p1, p2, lengths1=None, lengths2=None, K: int = 1, return_nn: bool = False,
               return_sorted: bool = True, version: int = -1):
    """Exact K nearest neighbours (pytorch3d.ops.knn_points semantics [third party, restated]:
    squared distances ascending, idx int64; rows with fewer than K candidates -- lengths2 < K -- are
    padded with 0 / 0).  K <= 32."""
    _ext.require_cuda(p1, p2)
    if K > 32:
        raise RuntimeError("knn_points: K must be <= 32 on the grid path")
    N, P1, D = p1.shape
    P2 = p2.shape[1]
    dev = p1.device
    if lengths1 is None:
        lengths1 = torch.full((N,), P1, dtype=torch.int64, device=dev)
    if lengths2 is None:
        lengths2 = torch.full((N,), P2, dtype=torch.int64, device=dev)
    p1 = p1.contiguous()
    p2 = p1 if p2 is p1 else p2.contiguous()
    need = torch.clamp(lengths2, max=K)                                   # hits every live query must reach
    live = torch.arange(P1, device=dev)[None, :] < lengths1[:, None]
    # initial radius: ball holding ~2K points at the cloud's mean density, measured in the dimensions the cloud
    # really spans (volume, area or length: an axis whose extent is < 1e-6 of the largest does not count -- a
    # planar cloud would otherwise start at r ~ 0 and ask FRNN for billions of cells).  Surface-like clouds are
    # denser locally, so this is usually an over-estimate.
    ext = (p2.amax(dim=1) - p2.amin(dim=1))
    diag = ext.norm(dim=-1).clamp_min(1e-12)
    live_ax = ext > 1e-6 * ext.amax(dim=-1, keepdim=True).clamp_min(1e-30)
    ndim = live_ax.sum(dim=-1).clamp_min(1).float()
    measure = torch.where(live_ax, ext, torch.ones_like(ext)).prod(dim=-1)
    unit_ball = torch.tensor([2.0, math.pi, 4.0 * math.pi / 3.0], device=dev)[(ndim - 1).long()]
    r = (measure * (2.0 * K) / lengths2.clamp_min(1).float() / unit_ball) ** (1.0 / ndim)
    r = torch.minimum(r.float(), diag)
    dists = idx = None
    ok = None
    for _ in range(14):
        dists, idx, _, _ = frnn.frnn_grid_points(p1, p2, lengths1, lengths2, K=K, r=r, return_nn=False)
        found = (idx >= 0).sum(-1)
        ok = (found >= need[:, None]) | ~live
        if bool(ok.all()):
            break
        done = r >= diag
        if bool(done.all()):
            break
        r = torch.minimum(r * 2.0, diag * 1.0001)
    if not bool(ok.all()):
        # last resort: one search at the bounding-box diagonal finds every point; a row still short of
        # min(K, lengths2) hits after that would be returned padded as if it had real neighbours
        dists, idx, _, _ = frnn.frnn_grid_points(p1, p2, lengths1, lengths2, K=K, r=diag * 1.0001, return_nn=False)
        found = (idx >= 0).sum(-1)
        if not bool(((found >= need[:, None]) | ~live).all()):
            raise RuntimeError("knn_points: some queries found fewer than min(K, lengths2) neighbours "
                               "(non-finite coordinates?)")
    pad = idx < 0
    idx = idx.masked_fill(pad, 0)
    dists = dists.masked_fill(pad, 0.0)
    nn = frnn.frnn_gather(p2, idx) if return_nn else None
    return _KNN(dists=dists, idx=idx, knn=nn)


def farthest_sampling(point_clouds, ratio: float, random_start: bool = False):
    """Farthest point subsampling of every cloud to ceil(ratio * n) points (:473-499; the
    reference calls torch_cluster.fps [third party] whose default start point is random -- here it
    is point 0 unless ``random_start``, so results are reproducible)."""
    pts, num = convert_pointclouds_to_tensor(point_clouds)
    _ext.require_cuda(pts)
    N, P, _ = pts.shape
    dev = pts.device
    m = torch.ceil(num.double() * ratio).to(torch.int64)
    Mmax = int(math.ceil(P * ratio)) if P else 0
    out = torch.empty((N, Mmax), dtype=torch.int64, device=dev)
    lib = _ext.lib()
    ws = torch.empty((max(int(lib.isob200_fps_ws_floats(N, P)), 1),), dtype=torch.float32, device=dev)
    start = None
    if random_start:
        start = (torch.rand((N,), device=dev) * num.float()).long()
    _ext.check(lib.isob200_fps_ws(_ext.ptr(pts.contiguous()), _ext.ptr(num), _ext.ptr(m), _ext.ptr(start), N, P,
                                  Mmax, _ext.ptr(ws), ws.numel(), _ext.ptr(out), _ext.stream(dev)))
    counts = m.tolist()
    sel = [out[n, : counts[n]] for n in range(N)]
    pts_list = [pts[n][sel[n]] for n in range(N)]
    if not is_pointclouds(point_clouds):
        return list_to_padded(pts_list), m
    sampled = point_clouds.__class__(pts_list)
    normals = point_clouds.normals_padded()
    if normals is not None:
        sampled.update_normals_([normals[n][sel[n]] for n in range(N)])
    feats = point_clouds.features_padded()
    if feats is not None:
        sampled.update_features_([feats[n][sel[n]] for n in range(N)])
    return sampled


def wlop(pointclouds, ratio: float = 0.5, neighborhood_size=16, iters=3, repulsion_mu=0.5, noise=None):
    """Weighted locally optimal projection (:35-122).  ``noise``: optional (sum_X, 3) standard-normal
    tensor used for the initial jitter instead of ``torch.randn_like`` (:59) so that runs are
    repeatable / comparable."""
    P, num_points_P = convert_pointclouds_to_tensor(pointclouds)
    _ext.require_cuda(P)
    lib = _ext.lib()
    dev = P.device
    P = P.contiguous()
    N = P.shape[0]
    mn = torch.stack([P[n, : int(num_points_P[n])].amin(0) for n in range(N)]) if N else P.new_zeros((0, 3))
    mx = torch.stack([P[n, : int(num_points_P[n])].amax(0) for n in range(N)]) if N else P.new_zeros((0, 3))
    diag = torch.norm(mn - mx, dim=-1)
    h = 4 * torch.sqrt(diag / num_points_P.float())
    search_radius = min(h * neighborhood_size, 0.2)                # python min on a 1-element tensor (:48)
    if not torch.is_tensor(search_radius):
        search_radius = torch.full((N,), float(search_radius), device=dev)
    search_radius = search_radius.reshape(-1).float()
    theta_sigma_inv = (16 / h / h).float().contiguous()

    if ratio < 1.0:
        X0 = farthest_sampling(pointclouds, ratio=ratio)
    elif ratio == 1.0:
        X0 = pointclouds.clone() if is_pointclouds(pointclouds) else (P.clone(), num_points_P.clone())
    else:
        raise ValueError('ratio must be less or equal to 1.0')
    X, num_points_X = convert_pointclouds_to_tensor(X0) if is_pointclouds(X0) else X0
    X = X.contiguous()
    PX = X.shape[1]
    liveX = torch.arange(PX, device=dev)[None, :] < num_points_X[:, None]
    if noise is None:
        noise = torch.randn((int(num_points_X.sum()), 3), device=dev)
    off = torch.zeros_like(X)
    off[liveX] = noise.to(dev) * 1.0
    X = (X + off * (h * 0.1).view(-1, 1, 1)).contiguous()            # X0.offset_(randn * h * 0.1) (:59-60)

    K = neighborhood_size
    st = _ext.stream(dev)
    _, idx_pp, _, grid = frnn.frnn_grid_points(P, P, num_points_P, num_points_P, K=K + 1, r=search_radius,
                                               grid=None, return_nn=False)
    density_P = torch.empty(P.shape[:2], dtype=torch.float32, device=dev)
    _ext.check(lib.isob200_wlop_density(_ext.ptr(P), _ext.ptr(idx_pp), K + 1, 1, _ext.ptr(theta_sigma_inv), N,
                                        P.shape[1], K, _ext.ptr(density_P), st))
    for _ in range(iters):
        _, idx_xp, _, grid = frnn.frnn_grid_points(X, P, num_points_X, num_points_P, K=K, r=search_radius,
                                                   grid=grid, return_nn=False)
        _, idx_xx, _, _ = frnn.frnn_grid_points(X, X, num_points_X, num_points_X, K=K + 1, r=search_radius,
                                                grid=None, return_nn=False)
        Xn = torch.empty_like(X)
        _ext.check(lib.isob200_wlop_step(_ext.ptr(X), _ext.ptr(P), _ext.ptr(idx_xp), _ext.ptr(idx_xx), K + 1, 1,
                                         _ext.ptr(density_P), _ext.ptr(theta_sigma_inv), float(repulsion_mu), N, PX,
                                         P.shape[1], K, _ext.ptr(Xn), st))
        X = Xn
    if is_pointclouds(X0):
        return X0.update_padded(X)
    return X


def upsample(pcl, n_points: Union[int, torch.Tensor], num_points=None, neighborhood_size=16, knn_result=None):
    """Iteratively insert points into the sparsest regions until every cloud has ``n_points`` (:281-362)."""
    def _ret(points, num_points, return_pcl):
        if return_pcl:
            return pcl.__class__(padded_to_list(points, num_points.tolist()))
        return points, num_points

    return_pcl = is_pointclouds(pcl)
    points, num_points_in = convert_pointclouds_to_tensor(pcl)
    if num_points is None or return_pcl:
        num_points = num_points_in
    _ext.require_cuda(points)
    lib = _ext.lib()
    dev = points.device
    K = neighborhood_size
    if int(num_points.sum()) == 0:
        return _ret(points, num_points, return_pcl)
    n_remaining = (n_points - num_points).to(dtype=torch.long)
    if bool((n_remaining <= 0).all()):
        return _ret(points, num_points, return_pcl)
    while not bool((n_remaining == 0).all()):
        points = points.contiguous()
        B, P, _ = points.shape
        max_P = P // 8
        knn = knn_points(points, points, num_points, num_points, K=K + 1)
        sparsity = torch.empty((B, P), dtype=torch.float32, device=dev)
        child = torch.empty((B, P, 3), dtype=torch.float32, device=dev)
        _ext.check(lib.isob200_upsample_sparsity(_ext.ptr(points), None, 0.0, _ext.ptr(knn.idx), K + 1, 1,
                                                 _ext.ptr(num_points), B, P, K, _ext.ptr(sparsity), _ext.ptr(child),
                                                 _ext.stream(dev)))
        order = sparsity.sort(dim=1).indices[:, P - max_P:] if max_P > 0 else sparsity.new_zeros((B, 0)).long()
        n_new = torch.clamp(n_remaining, max=max_P)
        new_pts = torch.gather(child, 1, order.unsqueeze(-1).expand(-1, -1, 3))
        total = []
        nn_list = n_new.tolist()
        np_list = num_points.tolist()
        for b in range(B):
            total.append(torch.cat([new_pts[b][max_P - nn_list[b]:], points[b, : np_list[b]]], dim=0))
        points = list_to_padded(total)
        n_remaining = n_remaining - n_new
        num_points = n_new + num_points
        if max_P == 0:
            break
    return _ret(points, num_points, return_pcl)


def resample_uniformly(pointclouds, neighborhood_size: int = 8, knn=None, normals=None, shrink_ratio: float = 0.5,
                       repulsion_mu: float = 1.0, noise=None):
    """WLOP consolidation to ``shrink_ratio`` of the points, then ``upsample`` back (:126-166).  The
    reference also builds a K-NN and normals it never uses (:141-158); those dead steps are skipped.
    Returns a ``Pointclouds`` for a ``Pointclouds`` input (:164-165) and ``(padded points, num_points)`` for a
    tensor input -- the convention the reference documents (:131-132, :166); its own tensor path never gets there
    (``wlop`` calls ``pointclouds.get_bounding_boxes()``, :44).  ``noise``: see ``wlop``."""
    points_init, num_points = convert_pointclouds_to_tensor(pointclouds)
    wl = wlop(pointclouds, ratio=shrink_ratio, repulsion_mu=repulsion_mu, noise=noise)
    if is_pointclouds(pointclouds):
        return upsample(wl, num_points)
    x, nx = (wl, torch.ceil(num_points.double() * shrink_ratio).long()) if torch.is_tensor(wl) else wl
    return upsample(x, num_points, num_points=nx)
