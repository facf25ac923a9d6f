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
