"""Drop-in for the reference's ``frnn`` package (external/FRNN/frnn/frnn.py) on libisob200.so.

Same surface, argument meaning, return conventions and error behaviour:

* ``frnn_grid_points(points1, points2, lengths1, lengths2, K, r, grid, return_nn, return_sorted,
  radius_cell_ratio) -> (dists, idxs, nn, grid)``  (frnn.py:176-301): ``dists`` squared fp32,
  ascending, padded with -1; ``idxs`` int64 padded with -1; ``grid`` a reusable ``_GRID``;
  differentiable w.r.t. both point sets through ``dists`` (frnn.py:165-173).
* ``frnn_gather(x, idxs, lengths)`` (frnn.py:304-352).
* ``_C`` exposes the in-place primitives the DSS splat backward reaches into
  (``insert_points_cuda``, ``counting_sort_cuda``: DSS/core/rasterizer.py:909-929) and
  ``prefix_sum_cuda`` mirrors ``prefix_sum.prefix_sum_cuda`` (frnn.py:11).

What differs by design (B200): grid parameters are computed on the device (no per-cloud
``.item()`` loop, frnn.py:55-71 -- one 4-byte read-back of the grid size G remains because the
offsets tensor has shape (N, G)); the grid build is a single deterministic pipeline (in-cell
order = ascending original index, where the reference's is atomic-order); queries are NOT
re-sorted through a second insert/scan/sort (frnn.py:106-130): the query kernel walks them in
the order of the cell-sorted reference array when points1 is points2, else in input order.
"""
from collections import namedtuple
from typing import Union

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _ext

_GRID = namedtuple("GRID", "sorted_points2 pc2_grid_off sorted_points2_idxs grid_params")

# traversal of the query kernel: 0 = auto, 1 = exhaustive block scan, 2 = pruned best-first, 3 = thread-per-query
# collect-then-select (same results)
QUERY_MODE = 0
# hint for the auto mode (bit 10 of the C entry's group_width): the radius spans many point spacings, so the K-th
# neighbour lies far inside it -- set by callers that choose r that way (UniformProjection._create_tree)
FAR_RADIUS_HINT = False

# When a list (set by a caller that defers its own read-backs, UniformProjection._filter_resample_run_ahead):
# build_grid sizes the cell table for SPECULATIVE_CELLS cells instead of reading the grid size back, and appends
# (device int32 with the true size, the capacity used); the caller compares them after its own synchronisation and
# redoes the search when a grid did not fit (the kernels stay in bounds meanwhile: isob200_frnn_grid_params_capped)
DEFERRED_GRID_CHECKS = None
SPECULATIVE_CELLS = 2 * 129 ** 3   # twice the largest grid of a cube-shaped box (128 cells along the shortest axis)

_MAX_CELLS = 1 << 28     # 1 GiB of int32 offsets per cloud; int cell ids stay far from overflow
_PARAMS_SIZE = {2: 6, 3: 8}
_TOTAL_IDX = {2: 5, 3: 7}


def _i64(t, device):
    return t.to(device=device, dtype=torch.int64).contiguous()


def _grid_params(points2, lengths2, r, radius_cell_ratio):
    """frnn.py:55-71 on the device. Returns params (N, 6|8) fp32 and G (python int)."""
    N, P2, D = points2.shape
    lib = _ext.lib()
    params = torch.zeros((N, _PARAMS_SIZE[D]), dtype=torch.float32, device=points2.device)
    gmax = torch.zeros((1,), dtype=torch.int32, device=points2.device)
    ws = _ext.workspace(32 * max(N, 1) + 4, points2.device)
    if DEFERRED_GRID_CHECKS is not None:
        cap = max(1, min(SPECULATIVE_CELLS, _MAX_CELLS // max(N, 1)))
        _ext.check(lib.isob200_frnn_grid_params_capped(
            _ext.ptr(points2), _ext.ptr(lengths2), _ext.ptr(r), N, P2, D, float(radius_cell_ratio), cap,
            _ext.ptr(params), _ext.ptr(gmax), _ext.ptr(ws), ws.numel(), _ext.stream(points2.device)))
        DEFERRED_GRID_CHECKS.append((gmax, cap))
        return params, cap
    _ext.check(lib.isob200_frnn_grid_params(
        _ext.ptr(points2), _ext.ptr(lengths2), _ext.ptr(r), N, P2, D, float(radius_cell_ratio),
        _ext.ptr(params), _ext.ptr(gmax), _ext.ptr(ws), ws.numel(), _ext.stream(points2.device)))
    G = int(gmax.item())
    if G > _MAX_CELLS:
        raise RuntimeError("frnn: the grid for this radius needs more than %d cells (a search radius far below "
                           "the point spacing of a cloud with a collapsed axis?); use a larger r" % _MAX_CELLS)
    return params, G


def build_grid(points2, lengths2, r, radius_cell_ratio=2.0):
    """Deterministic grid over points2: (sorted_points2, pc2_grid_off, sorted_points2_idxs, params)."""
    N, P2, D = points2.shape
    lib = _ext.lib()
    dev = points2.device
    params, G = _grid_params(points2, lengths2, r, radius_cell_ratio)
    G = max(G, 1)
    off = torch.empty((N, G), dtype=torch.int32, device=dev)
    sorted_points2 = torch.empty_like(points2)
    sorted_idxs = torch.empty((N, P2), dtype=torch.int32, device=dev)
    nbytes = lib.isob200_frnn_build_ws_bytes(N, P2, G)
    ws = _ext.workspace(nbytes, dev)
    _ext.check(lib.isob200_frnn_build(
        _ext.ptr(points2), _ext.ptr(lengths2), _ext.ptr(params), N, P2, D, G, _ext.ptr(off),
        _ext.ptr(sorted_points2), _ext.ptr(sorted_idxs), _ext.ptr(ws), ws.numel(), _ext.stream(dev)))
    return _GRID(sorted_points2, off, sorted_idxs, params)


def find_nbrs(points1, lengths1, lengths2, grid, K, r, q_points=None, q_order=None, idx_dtype=torch.int64):
    """K-bounded radius query of points1 against a built grid -> (idxs, dists)."""
    N, P1, D = points1.shape
    sorted_points2, off, sorted_idxs2, params = grid
    P2 = sorted_points2.shape[1]
    G = off.shape[1]
    dev = points1.device
    if not (1 <= K <= 32):
        print("Invalid range: K must be in [1, 32], got", K)
        raise RuntimeError("Invalid range")  # utils/dispatch.h:29-37 prints and throws
    dists = torch.empty((N, P1, K), dtype=torch.float32, device=dev)
    idxs = torch.empty((N, P1, K), dtype=idx_dtype, device=dev)
    qp = points1 if q_points is None else q_points
    _ext.check(_ext.lib().isob200_frnn_find_nbrs(
        _ext.ptr(qp), _ext.ptr(q_order), _ext.ptr(lengths1), _ext.ptr(lengths2),
        _ext.ptr(sorted_points2), _ext.ptr(off), _ext.ptr(sorted_idxs2), _ext.ptr(params), _ext.ptr(r),
        N, P1, P2, D, G, K, _ext.ptr(dists), _ext.ptr(idxs), 1 if idx_dtype == torch.int64 else 0,
        ((QUERY_MODE & 3) << 8) | ((1 << 10) if FAR_RADIUS_HINT else 0), _ext.stream(dev)))
    return idxs, dists


class _frnn_grid_points(Function):
    """autograd wrapper; mirrors frnn.py:14-173."""

    @staticmethod
    def forward(ctx, points1, points2, lengths1, lengths2, K, r, sorted_points2=None,
                pc2_grid_off=None, sorted_points2_idxs=None, grid_params_cuda=None,
                return_sorted=True, radius_cell_ratio=2.0):
        D = points1.shape[2]
        assert D == 2 or D == 3, "For now only 2D/3D is supported"
        use_cached = (sorted_points2 is not None and pc2_grid_off is not None
                      and sorted_points2_idxs is not None and grid_params_cuda is not None)
        same = (points1.data_ptr() == points2.data_ptr() and points1.shape == points2.shape
                and lengths1.data_ptr() == lengths2.data_ptr())
        if not use_cached:
            grid = build_grid(points2, lengths2, r, radius_cell_ratio)
        else:
            grid = _GRID(sorted_points2, pc2_grid_off, sorted_points2_idxs, grid_params_cuda)
        if same and not use_cached:
            # self-query: walk queries in cell-sorted order for locality (what the reference gets
            # by sorting points1 a second time, frnn.py:106-130)
            idxs, dists = find_nbrs(points1, lengths1, lengths2, grid, K, r,
                                    q_points=grid.sorted_points2, q_order=grid.sorted_points2_idxs)
        else:
            idxs, dists = find_nbrs(points1, lengths1, lengths2, grid, K, r)
        ctx.save_for_backward(points1, points2, lengths1, lengths2, idxs)
        ctx.mark_non_differentiable(idxs)
        ctx.mark_non_differentiable(grid.sorted_points2)
        ctx.mark_non_differentiable(grid.pc2_grid_off)
        ctx.mark_non_differentiable(grid.sorted_points2_idxs)
        ctx.mark_non_differentiable(grid.grid_params)
        return (idxs, dists, grid.sorted_points2, grid.pc2_grid_off, grid.sorted_points2_idxs,
                grid.grid_params)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_idxs, grad_dists, g2, g3, g4, g5):
        points1, points2, lengths1, lengths2, idxs = ctx.saved_tensors
        grad_points1, grad_points2 = _C.frnn_backward_cuda(points1, points2, lengths1, lengths2, idxs,
                                                           grad_dists.contiguous())
        return (grad_points1, grad_points2) + (None,) * 10


def frnn_grid_points(points1: torch.Tensor, points2: torch.Tensor,
                     lengths1: Union[torch.Tensor, None] = None,
                     lengths2: Union[torch.Tensor, None] = None, K: int = -1,
                     r: Union[float, torch.Tensor] = -1, grid: Union[_GRID, None] = None,
                     return_nn: bool = False, return_sorted: bool = True,
                     radius_cell_ratio: float = 2.0):
    """Fixed-radius K-nearest-neighbour search on a uniform grid (frnn.py:176-301)."""
    if points1.shape[0] != points2.shape[0]:
        raise ValueError("points1 and points2 must have the same batch  dimension")
    if points1.shape[2] != points2.shape[2]:
        raise ValueError(f"dimension mismatch: points1 of dimension {points1.shape[2]} while points2 of dimension {points2.shape[2]}")
    if points1.shape[2] != 2 and points1.shape[2] != 3:
        raise ValueError("for now only grid in 2D/3D is supported")
    if not points1.is_cuda or not points2.is_cuda:
        raise TypeError("for now only cuda version is supported")

    same_storage = points1 is points2
    points1 = points1.contiguous()
    points2 = points1 if same_storage else points2.contiguous()
    if points1.dtype != torch.float32 or points2.dtype != torch.float32:
        raise RuntimeError("expected scalar type Float")
    P1 = points1.shape[1]
    P2 = points2.shape[1]
    N = points1.shape[0]
    dev = points1.device

    same_len = lengths1 is lengths2
    if lengths1 is None:
        lengths1 = torch.full((N,), P1, dtype=torch.long, device=dev)
    else:
        lengths1 = _i64(lengths1, dev)
    if lengths2 is None:
        lengths2 = lengths1 if (same_len and P1 == P2) else torch.full((N,), P2, dtype=torch.long, device=dev)
    else:
        lengths2 = lengths1 if same_len else _i64(lengths2, dev)

    if isinstance(r, (float, int)):
        r = torch.ones((N,), dtype=torch.float32) * r
    if isinstance(r, torch.Tensor):
        assert (len(r.shape) == 1 and (r.shape[0] == 1 or r.shape[0] == N))
        if r.shape[0] == 1:
            r = r * torch.ones((N,), dtype=r.dtype, device=r.device)
    r = r.type(torch.float32)
    if r.device != dev:
        r = r.to(dev)
    r = r.contiguous()

    if grid is not None:
        out = _frnn_grid_points.apply(points1, points2, lengths1, lengths2, K, r, grid[0], grid[1],
                                      grid[2], grid[3], return_sorted, radius_cell_ratio)
    else:
        out = _frnn_grid_points.apply(points1, points2, lengths1, lengths2, K, r, None, None, None,
                                      None, return_sorted, radius_cell_ratio)
    idxs, dists, sorted_points2, pc2_grid_off, sorted_points2_idxs, grid_params_cuda = out
    grid = _GRID(sorted_points2=sorted_points2, pc2_grid_off=pc2_grid_off,
                 sorted_points2_idxs=sorted_points2_idxs, grid_params=grid_params_cuda)
    points2_nn = None
    if return_nn:
        points2_nn = frnn_gather(points2, idxs, lengths2)
    return dists, idxs, points2_nn, grid


class _frnn_gather(Function):
    @staticmethod
    def forward(ctx, x, idxs):
        N, M, U = x.shape
        _, L, K = idxs.shape
        out = torch.empty((N, L, K, U), dtype=torch.float32, device=x.device)
        _ext.check(_ext.lib().isob200_frnn_gather(
            _ext.ptr(x), _ext.ptr(idxs), 1 if idxs.dtype == torch.int64 else 0, N, M, L, K, U,
            _ext.ptr(out), _ext.stream(x.device)))
        ctx.save_for_backward(idxs)
        ctx.shape = (N, M, U)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        (idxs,) = ctx.saved_tensors
        N, M, U = ctx.shape
        _, L, K = idxs.shape
        grad_out = grad_out.contiguous()
        grad_x = torch.zeros((N, M, U), dtype=torch.float32, device=grad_out.device)
        _ext.check(_ext.lib().isob200_frnn_gather_backward(
            _ext.ptr(grad_out), _ext.ptr(idxs), 1 if idxs.dtype == torch.int64 else 0, N, M, L, K, U,
            _ext.ptr(grad_x), _ext.stream(grad_out.device)))
        return grad_x, None


def frnn_gather(x: torch.Tensor, idxs: torch.Tensor, lengths: Union[torch.Tensor, None] = None):
    """x (N,M,U), idxs (N,L,K) -> (N,L,K,U), zeros where idxs < 0 (frnn.py:304-352)."""
    N, P2, D = x.shape
    _N, P1, K = idxs.shape
    if N != _N:
        raise ValueError("x and idxs must have same batch dimension")
    _ext.require_cuda(x, idxs)
    if idxs.dtype not in (torch.int64, torch.int32):
        raise RuntimeError("frnn_gather: idxs must be int64 or int32")
    xf = x.contiguous()
    if xf.dtype != torch.float32:
        # rare non-float payloads (masks, ids): same semantics through the float kernel is lossy,
        # so use the index formulation of the reference for them
        tmp = idxs.clone()
        tmp[idxs < 0] = 0
        out = xf[:, :, None].expand(-1, -1, K, -1).gather(1, tmp[:, :, :, None].expand(-1, -1, -1, D).long())
        out[(idxs < 0)[:, :, :, None].expand(-1, -1, -1, D)] = 0
        return out
    return _frnn_gather.apply(xf, idxs.contiguous())


class _CExt:
    """The subset of ``frnn._C`` (ext.cpp:7-24) that other reference code calls directly."""

    @staticmethod
    def insert_points_cuda(points, lengths, params, grid_cnt, grid_cell, grid_idx, G):
        _ext.require_cuda(points)
        N, P, D = points.shape
        _ext.check(_ext.lib().isob200_frnn_insert_points(
            _ext.ptr(points.contiguous()), _ext.ptr(lengths), _ext.ptr(params), _ext.ptr(grid_cnt),
            _ext.ptr(grid_cell), _ext.ptr(grid_idx), N, P, D, int(G), _ext.stream(points.device)))

    @staticmethod
    def counting_sort_cuda(points, lengths, grid_cell, grid_idx, grid_off, sorted_points, sorted_idxs):
        _ext.require_cuda(points)
        N, P, D = points.shape
        G = grid_off.shape[1]
        _ext.check(_ext.lib().isob200_frnn_counting_sort(
            _ext.ptr(points.contiguous()), _ext.ptr(lengths), _ext.ptr(grid_cell), _ext.ptr(grid_idx),
            _ext.ptr(grid_off), _ext.ptr(sorted_points), _ext.ptr(sorted_idxs), N, P, D, G,
            _ext.stream(points.device)))

    @staticmethod
    def find_nbrs_cuda(points1, points2, lengths1, lengths2, pc2_grid_off, sorted_points1_idxs,
                       sorted_points2_idxs, params, K, rs, r2s):
        """grid.cu:384-440: points1/points2 are the cell-sorted arrays; results land in ORIGINAL
        query rows (sorted_points1_idxs)."""
        grid = _GRID(points2, pc2_grid_off, sorted_points2_idxs, params)
        return find_nbrs(points1, lengths1, lengths2, grid, K, rs, q_points=points1,
                         q_order=sorted_points1_idxs)

    @staticmethod
    def frnn_backward_cuda(points1, points2, lengths1, lengths2, idxs, grad_dists):
        N, P1, D = points1.shape
        P2 = points2.shape[1]
        K = idxs.shape[2]
        g1 = torch.zeros_like(points1)
        g2 = torch.zeros_like(points2)
        _ext.check(_ext.lib().isob200_frnn_backward(
            _ext.ptr(points1.contiguous()), _ext.ptr(points2.contiguous()), _ext.ptr(lengths1),
            _ext.ptr(lengths2), _ext.ptr(idxs), _ext.ptr(grad_dists), N, P1, P2, D, K, _ext.ptr(g1),
            _ext.ptr(g2), _ext.stream(points1.device)))
        return g1, g2


_C = _CExt()


def prefix_sum_cuda(grid_cnt, num_cells, grid_off):
    """prefix_sum.prefix_sum_cuda(cnt (G,), num_cells, off (G,)) -- exclusive scan, in place on
    ``grid_off`` (external/FRNN/external/prefix_sum/prefix_sum.cu:74-87); runs on the CURRENT
    stream (the reference uses the default stream)."""
    _ext.require_cuda(grid_cnt, grid_off)
    n = int(num_cells)
    lib = _ext.lib()
    ws = _ext.workspace(lib.isob200_exclusive_scan_ws_bytes(n, 1), grid_cnt.device)
    _ext.check(lib.isob200_exclusive_scan_i32(_ext.ptr(grid_cnt), _ext.ptr(grid_off), n, 1, n, n,
                                              _ext.ptr(ws), ws.numel(), _ext.stream(grid_cnt.device)))
