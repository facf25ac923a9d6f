"""Point-sharded / view-sharded multi-GPU execution of the iso-points hot path.

The reference is single-process, single-GPU (no distributed call site anywhere); this
module is new work that shards the path the way it decomposes naturally, one process
per GPU over ``torch.distributed`` (NCCL on GPUs; the collectives are device agnostic, so the host
logic is covered by world_size-2 gloo tests on CPU):

* projection (``_project_points``): embarrassingly parallel over contiguous point ranges, SDF
  weights replicated -- no collective;
* resample: queries stay sharded, the reference cloud is replicated by ONE all-gather of
  (xyz, unit normal) = 24 B/point per sample iteration; every rank builds the full uniform grid
  locally (the build is a few % of the search) and searches only its own queries, so neighbour
  sets -- and therefore the moved points -- are identical to the single-GPU result on the
  concatenated cloud;
* splat: views are sharded across ranks (each rank rasterises its own views for the replicated
  point set); the per-point gradients of the shared points are summed with ONE all-reduce.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _ext
from . import frnn
from .levelset_sampling import ProjectionResult, UniformProjection


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n` items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def all_gather_varlen(x: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """Concatenate per-rank (n_r, C) tensors along dim 0 -> ((sum n_r, C), [n_0, ..., n_{W-1}]).

    One small all-gather of the row counts (the only host read-back) and one all-gather of the
    payload padded to the largest shard."""
    world = dist.get_world_size(group)
    n = torch.tensor([x.shape[0]], dtype=torch.int64, device=x.device)
    counts_t = torch.empty((world,), dtype=torch.int64, device=x.device)
    dist.all_gather_into_tensor(counts_t, n, group=group)
    counts = [int(c) for c in counts_t.tolist()]
    nmax = max(counts) if counts else 0
    tail = tuple(x.shape[1:])
    if nmax == 0:
        return x.new_zeros((0,) + tail), counts
    padded = x.new_zeros((nmax,) + tail)
    padded[: x.shape[0]] = x
    out = x.new_empty((world * nmax,) + tail)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    out = out.view((world, nmax) + tail)
    if all(c == nmax for c in counts):
        return out.reshape((world * nmax,) + tail), counts
    return torch.cat([out[r, : counts[r]] for r in range(world)], dim=0), counts


def all_reduce_point_grads(grad: torch.Tensor, group=None) -> torch.Tensor:
    """Sum per-point gradients of a replicated point set over the view-sharded ranks (in place)."""
    dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
    return grad


def shard_views(n_views: int, rank: int, world: int) -> List[int]:
    """Views rendered by `rank` (contiguous blocks, like the point ranges)."""
    b, e = shard_range(n_views, rank, world)
    return list(range(b, e))


class ShardedUniformProjection(UniformProjection):
    """``UniformProjection`` whose input is THIS RANK's contiguous shard of one cloud (B = 1).

    ``project_points`` / ``_project_points`` / ``resample`` keep the reference signatures
    (levelset_sampling.py:239-439); with ``skip_upsampling=True`` the results are the rank's shard of what the
    single-GPU operator returns for the concatenated cloud.  The insert / upsample branches (:411-433) need
    neighbourhoods and a sparsity ranking over the WHOLE cloud and are not sharded: ``project_points`` raises
    unless ``skip_upsampling=True`` (run them on the gathered cloud with ``UniformProjection``).

    Every rank must make the same sequence of collective calls, so the reference's data-dependent early exit
    ("nothing converged", :396-399) is decided on the GLOBAL survivor count, and a rank whose shard is empty
    still takes part in every exchange of ``resample``."""

    def __init__(self, *args, group=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.group = group

    def project_points(self, point_clouds, model, normals_init=None, skip_resampling=False,
                       skip_upsampling=False, ref_pcl=None, **kwargs):
        if not skip_upsampling:
            raise NotImplementedError(
                "ShardedUniformProjection.project_points: the insert / upsample branches are not point-sharded; "
                "pass skip_upsampling=True and upsample the gathered cloud with UniformProjection")
        return super().project_points(point_clouds, model, normals_init=normals_init,
                                      skip_resampling=skip_resampling, skip_upsampling=True, ref_pcl=ref_pcl, **kwargs)

    def _nothing_converged(self, counts) -> bool:
        """Collective version of the early exit: all ranks leave, or none does (a rank that returned alone would
        leave the others waiting in resample's all-gather)."""
        if not (dist.is_available() and dist.is_initialized()):
            return sum(counts) == 0
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(self.group) == "nccl" else "cpu"
        total = torch.tensor([sum(counts)], dtype=torch.int64, device=dev)
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return int(total.item()) == 0

    def _search_own_rows(self, pts_loc, g_pts, begin, len1, len2, radius):
        """K + 1 nearest gathered points of every own point (FRNN, the tree of levelset_sampling.py:110-140).  The
        gathered cloud holds this rank's shard as rows [begin, begin + nloc), so the cell order of the own queries
        falls out of the grid build: they are walked in that order (what a self-query gets from
        ``frnn_grid_points``; in shard order neighbouring queries touch unrelated cells: 0.73 ms instead of
        0.3 ms for 120 k of 970 k points).  Same results; no read-back."""
        nloc, dev = pts_loc.shape[0], pts_loc.device
        radius = radius.to(torch.float32).contiguous()
        grid = frnn.build_grid(g_pts[None], len2, radius)
        s = grid.sorted_points2_idxs[0]
        own = (s >= begin) & (s < begin + nloc)
        slot = torch.where(own, torch.cumsum(own, 0) - 1, nloc)            # the others share a spare slot
        order = torch.empty((nloc + 1,), dtype=torch.int32, device=dev).scatter_(0, slot, s - begin)[:nloc]
        q_pts = pts_loc[order.long()]
        hint, frnn.FAR_RADIUS_HINT = frnn.FAR_RADIUS_HINT, True     # r = knn_k point spacings, as in _create_tree
        try:
            idx, _ = frnn.find_nbrs(pts_loc[None], len1, len2, grid, self.knn_k + 1, radius,
                                    q_points=q_pts[None].contiguous(), q_order=order[None].contiguous())
        finally:
            frnn.FAR_RADIUS_HINT = hint
        return idx

    def resample(self, model, points_init, normals_init, num_points, sample_iters=None,
                 num_points_list=None, **forward_kwargs) -> ProjectionResult:
        sample_iters = sample_iters or self.sample_iters
        if points_init.shape[0] != 1:
            raise ValueError("ShardedUniformProjection: one cloud per call (B = 1)")
        if sample_iters == 0:
            return ProjectionResult(points_init, normals_init,
                                    points_init.new_full(points_init.shape[:-1], True, dtype=torch.bool))
        _ext.require_cuda(points_init)
        lib = _ext.lib()
        dev = points_init.device
        if num_points_list is not None:
            nloc = int(num_points_list[0])
        else:
            nloc = int(points_init.shape[1]) if num_points is None else int(num_points[0])
        pts_loc = points_init[0, :nloc].contiguous()
        nrm_loc = torch.empty_like(pts_loc)
        if nloc:
            _ext.check(lib.isob200_normalize_rows3(_ext.ptr(normals_init[0, :nloc].contiguous()), nloc, 1e-12,
                                                   _ext.ptr(nrm_loc), _ext.stream(dev)))
        result = None
        inv_sigma = None
        idx = None
        for it in range(sample_iters):
            # one exchange per sample iteration: xyz + unit normal of every rank's shard
            payload, counts = all_gather_varlen(torch.cat([pts_loc, nrm_loc], dim=1), self.group)
            g_pts = payload[:, :3].contiguous()
            g_nrm = payload[:, 3:].contiguous()
            ntot = g_pts.shape[0]
            if ntot * 3 < 2 * (self.knn_k + 1):      # levelset_sampling.py:251-252 on the whole cloud
                return ProjectionResult(points_init, normals_init,
                                        points_init.new_full(points_init.shape[:-1], True, dtype=torch.bool))
            if inv_sigma is None:                       # :254-256, from the initial (gathered) cloud
                diag0 = (g_pts.max(dim=0).values - g_pts.min(dim=0).values).norm()
                inv_sigma = (torch.full((1,), float(ntot), device=dev) / diag0).contiguous()
            if it % 2 == 0:                             # :261-266, neighbourhood refresh
                diag = (g_pts.max(dim=0).values - g_pts.min(dim=0).values).norm()
                radius = (torch.sqrt(diag / float(ntot)) * self.knn_k).reshape(1)
                len2 = torch.tensor([ntot], dtype=torch.int64, device=dev)
                len1 = torch.tensor([nloc], dtype=torch.int64, device=dev)
                if nloc:      # (an empty shard has nothing to search for, but stays in the exchange above)
                    idx = self._search_own_rows(pts_loc, g_pts, sum(counts[:dist.get_rank(self.group)]), len1, len2,
                                                radius)
            moved = torch.empty_like(pts_loc)
            if nloc:
                _ext.check(lib.isob200_resample_step(
                    _ext.ptr(pts_loc), _ext.ptr(g_pts), _ext.ptr(g_nrm), _ext.ptr(idx), 1, idx.shape[2], 1,
                    _ext.ptr(inv_sigma), 1, nloc, ntot, idx.shape[2] - 1, _ext.ptr(moved), _ext.stream(dev)))
            pts_loc = moved
            result = self._project_points(model, pts_loc[None], torch.tensor([nloc], device=dev),
                                          proj_max_iters=3, num_points_list=[nloc], **forward_kwargs)
        return result
