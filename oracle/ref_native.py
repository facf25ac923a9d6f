"""The reference's OWN native extensions (compiled unmodified by oracle/build_ref.py into
oracle/_ref/) as a live parity checker.

TEST INFRASTRUCTURE ONLY.  The .so files travel to the GPU box with the working tree; the
Python glue of the reference (frnn.py, rasterizer.py) does not, so the few host-side lines that
sequence the native calls are restated here, each citing the lines it follows.
"""
import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(HERE, "_ref")
_CACHE = {}


def _load(name):
    if name in _CACHE:
        return _CACHE[name]
    path = os.path.join(_REF, name, name + ".so")
    if not os.path.exists(path):
        _CACHE[name] = None
        return None
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _CACHE[name] = mod
    return mod


def frnn_C():
    """reference frnn._C (external/FRNN/frnn/csrc/ext.cpp:7-24) or None if not built."""
    return _load("ref_frnn_C")


def prefix_sum():
    """reference prefix_sum (external/FRNN/external/prefix_sum) or None."""
    return _load("ref_prefix_sum")


def dss_C():
    """reference DSS._C subset (DSS/csrc/ext.cpp:5-18) or None."""
    return _load("ref_dss_C")


def available():
    return all(os.path.exists(os.path.join(_REF, n, n + ".so")) for n in ("ref_frnn_C", "ref_prefix_sum", "ref_dss_C"))


# ---------------------------------------------------------------------------------------
def frnn_grid_points_cuda(points1, points2, lengths1, lengths2, K, r, radius_cell_ratio=2.0):
    """Host sequence of _frnn_grid_points.forward (frnn.py:55-162) driving the reference CUDA
    kernels: grid params loop, insert, per-cloud prefix sum, counting sort (points2 then
    points1), find_nbrs.  CUDA tensors only.  Returns (idxs i64, dists f32, sorted_points2,
    pc2_grid_off, sorted_points2_idxs, grid_params)."""
    C = frnn_C()
    ps = prefix_sum()
    N, P2, D = points2.shape
    P1 = points1.shape[1]
    dev = points1.device
    size, di, ti, max_res = (8, 3, 7, 128) if D == 3 else (6, 2, 5, 1024)
    params = torch.zeros((N, size), dtype=torch.float, device=dev)
    G = -1
    for i in range(N):
        gmin = points2[i, :lengths2[i]].min(dim=0)[0]
        gmax = points2[i, :lengths2[i]].max(dim=0)[0]
        params[i, :di] = gmin
        gsize = gmax - gmin
        cell = r[i].item() / radius_cell_ratio
        if cell < gsize.min() / max_res:
            cell = gsize.min() / max_res
        params[i, di] = 1 / cell
        params[i, di + 1:ti] = torch.floor(gsize / cell) + 1
        params[i, ti] = torch.prod(params[i, di + 1:ti])
        G = max(G, int(params[i, ti].item()))

    def build(pts, lens, P):
        cnt = torch.zeros((N, G), dtype=torch.int, device=dev)
        cell = torch.full((N, P), -1, dtype=torch.int, device=dev)
        idx = torch.full((N, P), -1, dtype=torch.int, device=dev)
        C.insert_points_cuda(pts, lens, params, cnt, cell, idx, G)
        off = torch.full((N, G), 0, dtype=torch.int, device=dev)
        pc = params.cpu()
        for i in range(N):
            ps.prefix_sum_cuda(cnt[i], pc[i, ti], off[i])
        spts = torch.zeros_like(pts)
        sidx = torch.full((N, P), -1, dtype=torch.int, device=dev)
        C.counting_sort_cuda(pts, lens, cell, idx, off, spts, sidx)
        return spts, off, sidx, cnt

    sp2, off2, sidx2, _ = build(points2, lengths2, P2)
    sp1, _, sidx1, _ = build(points1, lengths1, P1)
    idxs, dists = C.find_nbrs_cuda(sp1, sp2, lengths1, lengths2, off2, sidx1, sidx2, params, K, r, r * r)
    return idxs, dists, sp2, off2, sidx2, params


def frnn_bf_cpu(points1, points2, lengths1, lengths2, K, r):
    """reference FRNNBruteForceCPU (bruteforce_cpu.cpp:4-58): strict `dist < r2`, non-FMA g++
    arithmetic.  CPU tensors.  Returns (idxs, dists)."""
    return frnn_C().frnn_bf_cpu(points1, points2, lengths1, lengths2, K, float(r))


def splat_backward_fast_cuda(pts_screen, radii, idx, first_idx, num_points, occ_grad, zbuf_grad, radii_s):
    """Host sequence of EllipticalRasterizer.backward, fast path (DSS/core/rasterizer.py:850-968),
    driving the reference natives: visible filter (:851-863), padded views + per-view search radius
    = median(radii) * radii_s (:881-884), 2-D grid params (:887-903), frnn insert / prefix sum /
    counting sort (:909-929), re-gather by sorted index and packed offsets (:930-946), the fast
    occupancy kernel (:947), scatter back (:955-966) and the z-buffer scatter (:967).
    Returns (grad (P,3), search_radius (N,), pc_grid_off (N,G) packed-global, grid_params (N,6),
    sorted_global_idx (Pv,), visible mask (P,)).  CUDA tensors only."""
    C = dss_C()
    F = frnn_C()
    PS = prefix_sum()
    dev = pts_screen.device
    Ptot = pts_screen.shape[0]
    occupied = idx[..., 0] >= 0
    vis = torch.zeros(Ptot, dtype=torch.bool, device=dev)
    v = idx[occupied].unique().long().view(-1)
    vis[v[v >= 0]] = True
    num_v = torch.stack([x.sum() for x in torch.split(vis, num_points.tolist(), dim=0)])
    first_v = torch.zeros_like(num_v)
    first_v[1:] = num_v.cumsum(0)[:-1]
    pts_v = pts_screen[vis]
    rad_v = radii[vis]
    N = num_v.shape[0]
    Pv = pts_v.shape[0]
    maxP = int(num_v.max().item())

    def to_padded(x):
        out = x.new_zeros((N, maxP) + tuple(x.shape[1:]))
        for n in range(N):
            s, c = int(first_v[n]), int(num_v[n])
            out[n, :c] = x[s:s + c]
        return out

    def to_packed(x):
        return torch.cat([x[n, :int(num_v[n])] for n in range(N)], 0)

    pts_pad = to_padded(pts_v)
    rad_pad = to_padded(rad_v)
    search_r = torch.tensor([rad_pad[i, :num_v[i]].median() * radii_s for i in range(N)], dtype=torch.float,
                            device=dev)
    params = torch.zeros((N, 6), dtype=torch.float, device=dev)
    G = -1
    xy = pts_pad[:, :, :2].clone().contiguous()
    for i in range(N):
        gmin = xy[i, :num_v[i]].min(dim=0)[0]
        gmax = xy[i, :num_v[i]].max(dim=0)[0]
        params[i, :2] = gmin
        gsize = gmax - gmin
        cell = search_r[i].item() / 2
        if cell < gsize.min() / 1024:
            cell = gsize.min() / 1024
        params[i, 2] = 1 / cell
        params[i, 3:5] = torch.floor(gsize / cell) + 1
        params[i, 5] = torch.prod(params[i, 3:5])
        G = max(G, int(params[i, 5].item()))
    cnt = torch.zeros((N, G), dtype=torch.int, device=dev)
    cell_id = torch.full((N, maxP), -1, dtype=torch.int, device=dev)
    rank = torch.full((N, maxP), -1, dtype=torch.int, device=dev)
    F.insert_points_cuda(xy, num_v, params, cnt, cell_id, rank, G)
    pc = params.cpu()
    off = torch.full((N, G), 0, dtype=torch.int, device=dev)
    for i in range(N):
        PS.prefix_sum_cuda(cnt[i], pc[i, 5], off[i])
    sorted_xy = torch.zeros((N, maxP, 2), dtype=torch.float, device=dev)
    sorted_idx = torch.full((N, maxP), -1, dtype=torch.int, device=dev)
    F.counting_sort_cuda(xy, num_v, cell_id, rank, off, sorted_xy, sorted_idx)
    sorted_pts = torch.zeros_like(pts_pad)
    for i in range(N):
        j = sorted_idx[i, :num_v[i]].long().unsqueeze(1).expand(-1, 3)
        sorted_pts[i, :num_v[i]] = torch.gather(pts_pad[i], 0, j)
    pts_sorted_packed = to_packed(sorted_pts)
    gidx = to_packed(sorted_idx + first_v.float().unsqueeze(1)).long()       # through float32, as the reference
    gidx2 = gidx.unsqueeze(1).expand(-1, 2)
    rad_sorted = torch.gather(rad_v, 0, gidx2)
    off = off + first_v.unsqueeze(1).to(off.dtype)
    g_sorted = C._splat_points_occ_fast_cuda_backward(pts_sorted_packed, rad_sorted, search_r, occ_grad, num_v,
                                                      first_v, off, params)
    g_vis = torch.zeros_like(g_sorted).scatter_(0, gidx2, g_sorted)
    gxy = pts_screen.new_zeros(Ptot, 2)
    gz = pts_screen.new_zeros(Ptot, 1)
    gxy[vis] = g_vis
    C._backward_zbuf(idx, zbuf_grad, gz)
    return torch.cat([gxy, gz], dim=-1), search_r, off, params, gidx, vis
