"""CPU restatement (numpy / torch-CPU) of the reference's hot-path algorithms.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Paths cited are relative to the
yifita/iso-points tree.  Integer / index results are meant to be bit-exact with the reference's
CUDA kernels on tie-free inputs: where a comparison decides an index (radius test, K-th best,
splat coverage) the fp32 expression is evaluated exactly as nvcc contracts it in the reference
kernels (checked in the SASS of oracle/_ref/*.so: one FMUL followed by FFMAs), with fused
multiply-adds emulated through float64 (exact 24x24-bit product, one rounding).
"""
import math

import numpy as np
import torch

f32 = np.float32


def _fma(a, b, c):
    """round_f32(a*b + c) for float32 arrays (product exact in float64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


# =========================================================================================
# helpers  (DSS/utils/mathHelper.py:14-25)
# =========================================================================================
def eps_denom(x: torch.Tensor, eps: float = 1e-17) -> torch.Tensor:
    """(sign(x) + [x == 0]) * max(|x|, eps)  -- the sign of 0 is +1 (mathHelper.py:14-18)."""
    sgn = torch.where(x < 0, -torch.ones_like(x), torch.ones_like(x))
    return sgn * x.abs().clamp_min(eps)


# =========================================================================================
# FRNN  (external/FRNN/frnn/frnn.py, csrc/grid/*.cu, csrc/bruteforce/*)
# =========================================================================================
def frnn_sqdist(p2: np.ndarray, q: np.ndarray) -> np.ndarray:
    """grid.cu:331-334 as compiled (SASS, every K): d = p2 - q ; FMUL dy*dy ; FFMA dx ; FFMA dz."""
    d = (p2 - q[None, :]).astype(np.float32)
    s = (d[:, 1] * d[:, 1]).astype(np.float32)
    s = _fma(d[:, 0], d[:, 0], s)
    if p2.shape[1] == 3:
        s = _fma(d[:, 2], d[:, 2], s)
    return s


def frnn_bruteforce(points1, points2, lengths1=None, lengths2=None, K=8, r=0.1, inclusive=True):
    """Radius-bounded K nearest neighbours by exhaustive search.

    Semantics of FindNbrs{2,3}DKernel (grid.cu:186-349): candidates with sqdist <= r*r
    (``inclusive``; the brute-force kernels use a strict < : bruteforce.cu:42,
    bruteforce_cpu.cpp:44), the K smallest kept, ascending; ties by smaller index.
    Returns idxs (N,P1,K) int64 and dists (N,P1,K) float32, both padded with -1
    (grid.cu:422-423).  r: float or (N,) array.
    """
    points1 = np.ascontiguousarray(points1, dtype=np.float32)
    points2 = np.ascontiguousarray(points2, dtype=np.float32)
    N, P1, D = points1.shape
    P2 = points2.shape[1]
    rs = np.broadcast_to(np.asarray(r, dtype=np.float32).reshape(-1), (N,)) if np.ndim(r) else np.full((N,), r, np.float32)
    idxs = np.full((N, P1, K), -1, np.int64)
    dists = np.full((N, P1, K), -1, np.float32)
    for n in range(N):
        l1 = P1 if lengths1 is None else int(lengths1[n])
        l2 = P2 if lengths2 is None else int(lengths2[n])
        r2 = f32(rs[n]) * f32(rs[n])
        ref = points2[n, :l2]
        for i in range(l1):
            d = frnn_sqdist(ref, points1[n, i])
            cand = np.nonzero(d <= r2 if inclusive else d < r2)[0]
            if cand.size == 0:
                continue
            order = np.lexsort((cand, d[cand]))[:K]
            sel = cand[order]
            idxs[n, i, :sel.size] = sel
            dists[n, i, :sel.size] = d[sel]
    return idxs, dists


def frnn_grid_params(points2, lengths2, rs, radius_cell_ratio=2.0, cuda_semantics=True):
    """frnn.py:55-71 (3-D: max_res 128; 2-D: max_res 1024), fp32 like the params tensor.

    ``cuda_semantics``: `tensor / python_float` on a CUDA tensor is evaluated by ATen as
    tensor * (1.0f / (float)scalar); on a CPU tensor it is a true division.  The two can
    differ by one cell in `res` when size/cell lands on an integer.
    Returns params (N, 8|6) float32 and G.
    """
    points2 = np.asarray(points2, np.float32)
    N, P2, D = points2.shape
    max_res = 128 if D == 3 else 1024
    ps = 8 if D == 3 else 6
    params = np.zeros((N, ps), np.float32)
    G = -1
    for n in range(N):
        l2 = P2 if lengths2 is None else int(lengths2[n])
        gmin = points2[n, :l2].min(axis=0)
        gmax = points2[n, :l2].max(axis=0)
        size = (gmax - gmin).astype(np.float32)
        cell = float(f32(rs[n])) / radius_cell_ratio          # python double
        thresh = f32(size.min()) / f32(max_res)
        if f32(cell) < thresh:
            cell_t = thresh                                    # 0-dim fp32 tensor from here on
            delta = f32(1.0) / cell_t
            res = np.floor(size / cell_t).astype(np.float32) + f32(1)
        else:
            delta = f32(1.0 / cell)
            if cuda_semantics:
                res = np.floor(size * (f32(1.0) / f32(cell))).astype(np.float32) + f32(1)
            else:
                res = np.floor(size / f32(cell)).astype(np.float32) + f32(1)
        params[n, :D] = gmin
        params[n, D] = delta
        params[n, D + 1:2 * D + 1] = res
        tot = f32(res[0])
        for d in range(1, D):
            tot = f32(tot * res[d])
        params[n, 2 * D + 1] = tot
        G = max(G, int(tot))
    return params, G


def frnn_cell_ids(points, params_n):
    """grid.cu:121-129 / :83-89: (int)((p - min) * delta) truncation, clamped to the grid."""
    D = points.shape[1]
    delta = params_n[D]
    res = params_n[D + 1:2 * D + 1].astype(np.int64)
    g = ((points - params_n[None, :D]).astype(np.float32) * delta).astype(np.float32)
    g = np.trunc(g).astype(np.int64)
    g = np.clip(g, 0, res[None, :] - 1)
    cell = g[:, 0]
    for d in range(1, D):
        cell = cell * res[d] + g[:, d]
    return cell


def frnn_build_grid(points2, lengths2, params, G):
    """insert + exclusive scan + counting sort (grid.cu:95-133, prefix_sum.cu:112-121,
    counting_sort.cu:38-71) with the deterministic in-cell order 'ascending original index'
    (the reference's order inside a cell is atomicAdd arrival order)."""
    points2 = np.asarray(points2, np.float32)
    N, P2, D = points2.shape
    off = np.zeros((N, G), np.int32)
    sorted_pts = np.zeros_like(points2)
    sorted_idx = np.full((N, P2), -1, np.int32)
    for n in range(N):
        l2 = P2 if lengths2 is None else int(lengths2[n])
        cell = frnn_cell_ids(points2[n, :l2], params[n])
        cnt = np.bincount(cell, minlength=G)[:G]
        off[n] = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int32)
        order = np.argsort(cell, kind="stable")
        sorted_pts[n, :l2] = points2[n, order]
        sorted_idx[n, :l2] = order.astype(np.int32)
    return sorted_pts, off, sorted_idx


def frnn_grid_query(points1, lengths1, sorted_pts2, off2, sorted_idx2, lengths2, params, rs, K):
    """FindNbrs kernel over the grid (grid.cu:263-349): floor'd cell range, <= r2 test."""
    points1 = np.asarray(points1, np.float32)
    N, P1, D = points1.shape
    idxs = np.full((N, P1, K), -1, np.int64)
    dists = np.full((N, P1, K), -1, np.float32)
    for n in range(N):
        l1 = P1 if lengths1 is None else int(lengths1[n])
        l2 = sorted_pts2.shape[1] if lengths2 is None else int(lengths2[n])
        prm = params[n]
        r = f32(rs[n])
        r2 = f32(r * r)
        delta = prm[D]
        res = prm[D + 1:2 * D + 1].astype(np.int64)
        total = int(prm[2 * D + 1])
        ends = np.concatenate([off2[n, 1:total], [l2]]).astype(np.int64)
        for i in range(l1):
            q = points1[n, i]
            rel = (q - prm[:D]).astype(np.float32)
            lo = np.floor(((rel - r).astype(np.float32) * delta).astype(np.float32)).astype(np.int64)
            hi = np.floor(((rel + r).astype(np.float32) * delta).astype(np.float32)).astype(np.int64)
            lo = np.maximum(lo, 0)
            hi = np.minimum(hi, res - 1)
            if np.any(lo > hi):
                continue
            rngs = [np.arange(lo[d], hi[d] + 1) for d in range(D)]
            mesh = np.meshgrid(*rngs, indexing="ij")
            cell = mesh[0]
            for d in range(1, D):
                cell = cell * res[d] + mesh[d]
            cell = cell.reshape(-1)
            segs = [np.arange(off2[n, c], ends[c]) for c in cell]
            cand = np.concatenate(segs) if segs else np.zeros((0,), np.int64)
            if cand.size == 0:
                continue
            d2 = frnn_sqdist(sorted_pts2[n, cand], q)
            keep = d2 <= r2
            cand, d2 = cand[keep], d2[keep]
            orig = sorted_idx2[n, cand].astype(np.int64)
            order = np.lexsort((orig, d2))[:K]
            idxs[n, i, :order.size] = orig[order]
            dists[n, i, :order.size] = d2[order]
    return idxs, dists


def frnn_gather(x, idxs):
    """frnn.py:304-352: x (N,M,U), idxs (N,L,K) -> (N,L,K,U), zero where idx < 0."""
    x = np.asarray(x)
    N, L, K = idxs.shape
    out = np.zeros((N, L, K, x.shape[2]), x.dtype)
    for n in range(N):
        m = idxs[n] >= 0
        out[n][m] = x[n][idxs[n][m]]
    return out


def frnn_backward(points1, points2, idxs, grad_dists, lengths1=None, lengths2=None):
    """backward.cu:42-74 in float64: d dist / d p1 = 2 g (p1 - p2), and the negative for p2."""
    p1 = np.asarray(points1, np.float64)
    p2 = np.asarray(points2, np.float64)
    g1 = np.zeros_like(p1)
    g2 = np.zeros_like(p2)
    N, P1, K = idxs.shape
    for n in range(N):
        l1 = P1 if lengths1 is None else int(lengths1[n])
        l2 = p2.shape[1] if lengths2 is None else int(lengths2[n])
        for k in range(min(K, l2)):
            j = idxs[n, :l1, k]
            m = j >= 0
            rows = np.nonzero(m)[0]
            diff = 2.0 * grad_dists[n, rows, k, None].astype(np.float64) * (p1[n, rows] - p2[n, j[m]])
            g1[n, rows] += diff
            np.add.at(g2[n], j[m], -diff)
    return g1, g2


# =========================================================================================
# level-set projection + resampling  (DSS/models/levelset_sampling.py)
# =========================================================================================
def sdf_and_grad(points: torch.Tensor, model, max_points_per_pass=120000, **kw):
    """levelset_sampling.py:142-170: chunked sdf + d sdf / d x through autograd."""
    pts = points.reshape(-1, 3)
    vals, grads = [], []
    model.eval()
    for chunk in torch.split(pts, max_points_per_pass, dim=0):
        with torch.enable_grad():
            x = chunk.detach().clone().requires_grad_(True)
            s = model.forward(x, **kw).sdf
            g = torch.autograd.grad([s], [x], torch.ones_like(s))[0]
        vals.append(s.detach())
        grads.append(g)
    return torch.cat(vals, 0).reshape(points.shape[:-1]), torch.cat(grads, 0).reshape(points.shape)


def project_points_packed(model, points_packed: torch.Tensor, proj_max_iters=10, proj_tolerance=5e-5,
                          max_points_per_pass=120000, stats=None, **kw):
    """The Newton loop of UniformProjection._project_points (levelset_sampling.py:309-343) on a
    packed (M,3) tensor, written with explicit active-row lists.

    Returns (points (M,3), normals (M,3) = last raw gradient, valid (M,) bool).
    """
    with torch.no_grad():
        pts = points_packed.detach().clone()
        normals = torch.zeros_like(pts)
        active = torch.arange(pts.shape[0])
        it = 0
        while True:
            cur = pts[active]
            sdf, grad = sdf_and_grad(cur, model, max_points_per_pass, **kw)
            if stats is not None:
                stats.append(int(active.numel()))
            normals[active] = grad                                        # :322
            still = sdf.reshape(-1).abs() > proj_tolerance               # :326
            if (not bool(still.any())) or it == proj_max_iters:          # :329
                active = active[still]
                break
            it += 1
            g = grad[still]
            f = sdf.reshape(-1)[still]
            p = cur[still]
            ss = (g * g).sum(-1, keepdim=True)
            move = f[:, None] * (g / eps_denom(ss, 1.0e-17))             # :336-339
            nrm = move.norm(dim=-1, keepdim=True)
            move = move / nrm.clamp_min(1e-15) * nrm.clamp_max(0.1)      # :340-341
            active = active[still]
            pts[active] = p - move                                        # :342
        valid = torch.ones(pts.shape[0], dtype=torch.bool)
        valid[active] = False
    return pts, normals, valid


def project_points_padded(model, points: torch.Tensor, num_points, **kw):
    """_project_points on a padded (B,P,3) batch -> padded points/normals (zero padded), mask
    (levelset_sampling.py:307-308, 344-351)."""
    B, P, _ = points.shape
    nums = [int(n) for n in num_points]
    packed = torch.cat([points[b, :nums[b]] for b in range(B)], 0)
    pts, nrm, valid = project_points_packed(model, packed, **kw)
    pmax = max(nums) if nums else 0
    out_p = points.new_zeros((B, pmax, 3))
    out_n = points.new_zeros((B, pmax, 3))
    out_m = torch.zeros((B, pmax), dtype=torch.bool)
    o = 0
    for b in range(B):
        out_p[b, :nums[b]] = pts[o:o + nums[b]]
        out_n[b, :nums[b]] = nrm[o:o + nums[b]]
        out_m[b, :nums[b]] = valid[o:o + nums[b]]
        o += nums[b]
    return out_p, out_n, out_m


def resample_move(points: torch.Tensor, normals_unit: torch.Tensor, idx: torch.Tensor, inv_sigma: float):
    """One repulsion move of UniformProjection.resample (levelset_sampling.py:268-284), B = 1.

    points (P,3), normals_unit (P,3) already F.normalize'd, idx (P,K) int64 with -1 padding
    (self column already dropped).  Returns points + move.
    """
    m = idx >= 0
    j = idx.clamp_min(0)
    nn_p = points[j] * m[..., None]          # frnn_gather zero-fills (frnn.py:348-351)
    nn_n = normals_unit[j] * m[..., None]
    diff = points[:, None, :] - nn_p
    d2 = (diff * diff).sum(-1)
    w = torch.exp(-d2 * inv_sigma)
    w = w * m
    sw = w.sum(-1, keepdim=True)
    proj = diff - (diff * nn_n).sum(-1, keepdim=True) * nn_n
    move = (sw + 1.0) * (w[..., None] * proj).sum(-2) / eps_denom(sw)
    return points + move


def resample(model, points_init: torch.Tensor, normals_init: torch.Tensor, sample_iters=1, knn_k=8,
             frnn_fn=None, **proj_kw):
    """UniformProjection.resample (levelset_sampling.py:239-288) for one cloud (P,3).

    Neighbourhoods: FRNN self-query with K = knn_k+1, r = sqrt(diag/n)*knn_k, column 0 dropped as
    "self" (:126-138), refreshed on even iterations (:261).  Returns (points, normals, valid) of
    the last 3-iteration re-projection (:285-286).
    """
    P = points_init.shape[0]
    if points_init.numel() < 2 * (knn_k + 1) or sample_iters == 0:
        return points_init, normals_init, torch.ones(P, dtype=torch.bool)
    diag = float((points_init.max(0).values - points_init.min(0).values).norm())
    inv_sigma = float(P) / diag
    pts = points_init
    nrm = torch.nn.functional.normalize(normals_init, dim=-1)
    idx = None
    res = None
    for it in range(sample_iters):
        if it % 2 == 0:
            d = (pts.max(0).values - pts.min(0).values).norm()
            r = float(torch.sqrt(d / float(P)) * knn_k)
            fn = frnn_fn or (lambda p, K, r: frnn_bruteforce(p[None], p[None], K=K, r=r)[0][0])
            idx = torch.as_tensor(fn(pts.numpy(), knn_k + 1, r))[:, 1:]
        pts = resample_move(pts, nrm, idx, inv_sigma)
        res = project_points_packed(model, pts, proj_max_iters=3, **proj_kw)
    return res


# =========================================================================================
# DSS elliptical splat rasteriser  (DSS/csrc/rasterize_points.cu -- the CUDA kernels are the
# canonical semantics; the CPU twins in rasterize_points_cpu.cpp differ)
# =========================================================================================
def pix_to_ndc(i, S):
    """rasterization_utils.cuh:8-11 in fp32: -1 + (2 i + 1) / S."""
    i = np.asarray(i)
    return ((i.astype(np.float32) * f32(2) + f32(1)) / f32(S) + f32(-1)).astype(np.float32)


def splat_qvalue(a, b, c, dx, dy):
    """rasterize_points.cu:94 as compiled: fma(c*dy, dy, fma(a*dx, dx, (b*dx)*dy))."""
    t = ((b * dx).astype(np.float32) * dy).astype(np.float32)
    s = _fma((a * dx).astype(np.float32), dx, t)
    return _fma((c * dy).astype(np.float32), dy, s)


def splat_pairs(points, ellipse, cutoff, radii, first_idx, num_points, S):
    """Every (view, pixel, point) triple that passes CheckPixelInsidePoint
    (rasterize_points.cu:64-98): z >= 0, |dx| <= rx, |dy| <= ry, Q <= cutoff.

    Returns arrays (n, yi, xi, p, z, q) with yi/xi the NDC pixel indices (the output image is
    flipped: row = S-1-yi, col = S-1-xi, rasterize_points.cu:160-161).  len() of the result is
    the number of pixel-splats of the call.
    """
    points = np.asarray(points, np.float32)
    ellipse = np.asarray(ellipse, np.float32)
    radii = np.asarray(radii, np.float32)
    cutoff = np.broadcast_to(np.asarray(cutoff, np.float32).reshape(-1), (points.shape[0],))
    ndc = pix_to_ndc(np.arange(S), S)
    out = []
    for n in range(len(num_points)):
        s0 = int(first_idx[n])
        s1 = s0 + int(num_points[n])
        p = np.arange(s0, s1)
        px, py, pz = points[p, 0], points[p, 1], points[p, 2]
        rx, ry = radii[p, 0], radii[p, 1]
        ok = pz >= 0
        # conservative pixel window, then the exact fp32 tests
        lo_x = np.searchsorted(ndc, (px - rx).astype(np.float32) - f32(2.0 / S), "left")
        hi_x = np.searchsorted(ndc, (px + rx).astype(np.float32) + f32(2.0 / S), "right")
        lo_y = np.searchsorted(ndc, (py - ry).astype(np.float32) - f32(2.0 / S), "left")
        hi_y = np.searchsorted(ndc, (py + ry).astype(np.float32) + f32(2.0 / S), "right")
        wx = int(max((hi_x - lo_x).max(initial=0), 0))
        wy = int(max((hi_y - lo_y).max(initial=0), 0))
        if wx == 0 or wy == 0 or p.size == 0:
            continue
        ox, oy = np.meshgrid(np.arange(wx), np.arange(wy), indexing="xy")
        ox, oy = ox.reshape(1, -1), oy.reshape(1, -1)
        CH = max(1, 2_000_000 // (wx * wy))
        for c0 in range(0, p.size, CH):
            sl = slice(c0, c0 + CH)
            xi = lo_x[sl, None] + ox
            yi = lo_y[sl, None] + oy
            inb = (xi < hi_x[sl, None]) & (yi < hi_y[sl, None]) & (xi < S) & (yi < S) & ok[sl, None]
            xi_c = np.clip(xi, 0, S - 1)
            yi_c = np.clip(yi, 0, S - 1)
            dx = (ndc[xi_c] - px[sl, None]).astype(np.float32)
            dy = (ndc[yi_c] - py[sl, None]).astype(np.float32)
            inb &= ~((np.abs(dx) > rx[sl, None]) | (np.abs(dy) > ry[sl, None]))
            e = ellipse[p[sl]]
            q = splat_qvalue(np.broadcast_to(e[:, 0:1], dx.shape), np.broadcast_to(e[:, 1:2], dx.shape),
                             np.broadcast_to(e[:, 2:3], dx.shape), dx, dy)
            inb &= ~(q > cutoff[p[sl], None])
            r_, c_ = np.nonzero(inb)
            out.append((np.full(r_.size, n, np.int64), yi_c[r_, c_], xi_c[r_, c_], p[sl][r_],
                        np.broadcast_to(pz[sl, None], dx.shape)[r_, c_], q[r_, c_]))
    if not out:
        z = np.zeros((0,), np.int64)
        return z, z, z, z, np.zeros((0,), np.float32), np.zeros((0,), np.float32)
    return tuple(np.concatenate([o[i] for o in out]) for i in range(6))


def splat_forward(points, ellipse, cutoff, radii, first_idx, num_points, depth_merging_thres, S, K,
                  fine_occupancy=True, pairs=None):
    """RasterizePoints{Naive,Fine}CudaKernel (rasterize_points.cu:131-212, 506-597).

    Per pixel: the K covering points with the smallest z (ties -> smaller point index), ascending
    in z, cut where z - z[0] > depth_merging_thres (:203-206); occupancy = 1 where the largest
    kept z is > 0 (fine kernel, :581) or >= 0 (naive kernel, :196).
    Returns idx int32 (N,S,S,K), zbuf, qvalue float32 (N,S,S,K) (-1 padded), occ float32 (N,S,S).
    """
    N = len(num_points)
    if pairs is None:
        pairs = splat_pairs(points, ellipse, cutoff, radii, first_idx, num_points, S)
    n, yi, xi, p, z, q = pairs
    idx = np.full((N, S, S, K), -1, np.int32)
    zbuf = np.full((N, S, S, K), -1, np.float32)
    qv = np.full((N, S, S, K), -1, np.float32)
    occ = np.zeros((N, S, S), np.float32)
    if n.size == 0:
        return idx, zbuf, qv, occ
    pix = (n * S + (S - 1 - yi)) * S + (S - 1 - xi)
    order = np.lexsort((p, z, pix))
    pix, p, z, q = pix[order], p[order], z[order], q[order]
    start = np.r_[True, pix[1:] != pix[:-1]]
    seg_first = np.maximum.accumulate(np.where(start, np.arange(pix.size), 0))
    rank = np.arange(pix.size) - seg_first
    inK = rank < K
    z0 = z[seg_first]
    # q_max_z: max z among the K kept candidates (before the depth-merge cut)
    kmax = np.full(N * S * S, -1000.0, np.float32)
    np.maximum.at(kmax, pix[inK], z[inK])
    occ_flat = (kmax > 0) if fine_occupancy else (kmax >= 0)
    occ.reshape(-1)[occ_flat] = 1.0
    d = (z - z0).astype(np.float32) > f32(depth_merging_thres)
    # cut at the FIRST violating rank: z ascending => violation is monotone in rank
    keep = inK & ~d
    idx.reshape(-1, K)[pix[keep], rank[keep]] = p[keep].astype(np.int32)
    zbuf.reshape(-1, K)[pix[keep], rank[keep]] = z[keep]
    qv.reshape(-1, K)[pix[keep], rank[keep]] = q[keep]
    return idx, zbuf, qv, occ


def splat_bin_counts(points, radii, first_idx, num_points, S, bin_size):
    """points_per_bin of RasterizePointsCoarseCudaKernel (rasterize_points.cu:353-412), which the
    reference computes but never returns: points with z >= 0 whose box [p -+ r] overlaps the
    bin's NDC extent (inclusive), extents in fp32 exactly as the kernel forms them."""
    points = np.asarray(points, np.float32)
    radii = np.asarray(radii, np.float32)
    B = 1 + (S - 1) // bin_size
    half = f32(1.0) / f32(S)
    b = np.arange(B)
    lo = (pix_to_ndc(b * bin_size, S) - half).astype(np.float32)
    hi = (pix_to_ndc((b + 1) * bin_size - 1, S) + half).astype(np.float32)
    out = np.zeros((len(num_points), B, B), np.int32)
    for n in range(len(num_points)):
        s0 = int(first_idx[n])
        p = np.arange(s0, s0 + int(num_points[n]))
        ok = points[p, 2] >= 0
        x0 = (points[p, 0] - radii[p, 0]).astype(np.float32)
        x1 = (points[p, 0] + radii[p, 0]).astype(np.float32)
        y0 = (points[p, 1] - radii[p, 1]).astype(np.float32)
        y1 = (points[p, 1] + radii[p, 1]).astype(np.float32)
        ox = (x0[:, None] <= hi[None]) & (lo[None] <= x1[:, None]) & ok[:, None]   # (P,B)
        oy = (y0[:, None] <= hi[None]) & (lo[None] <= y1[:, None])
        out[n] = (oy.astype(np.int32).T @ ox.astype(np.int32))                    # [by, bx]
    return out


def blend(idx, qvalue, occ, scaler, rgb, eps=1e-4):
    """SurfaceSplattingRenderer.forward (DSS/core/renderer.py:53-78) in float64:
    w = exp(-0.5 q) * scaler[idx]; rgb = sum_k w f[idx] / max(sum_k w, eps) over idx >= 0; alpha = occ.
    The normalisation is pytorch3d's NormWeightedCompositor [third party, absent from the
    reference tree; eps = 1e-4 restated from pytorch3d/csrc/compositing/norm_weighted_sum.cu --
    parity unpinned]."""
    m = idx >= 0
    j = np.where(m, idx, 0)
    w = np.exp(-0.5 * qvalue.astype(np.float64)) * np.asarray(scaler, np.float64)[j] * m
    sw = np.maximum(w.sum(-1, keepdims=True), eps)
    col = (w[..., None] * np.asarray(rgb, np.float64)[j]).sum(-2) / sw
    return np.concatenate([col, occ[..., None].astype(np.float64)], -1)


def visibility(idx, occ, P):
    """get_per_point_visibility_mask (DSS/utils/__init__.py:378-399): ids in any slot of an
    occupied pixel."""
    vis = np.zeros(P, bool)
    v = idx[occ.astype(bool)]
    vis[v[v >= 0]] = True
    return vis


def splat_backward(points, radii, idx, first_idx, num_points, grad_occ, grad_zbuf, radii_s,
                   skip_last_cell_bug=False):
    """EllipticalRasterizer.backward, fast path (DSS/core/rasterizer.py:850-968 +
    rasterize_points_backward.cu:30-212 + rasterize_points.cu:823-846), float64 accumulation.

    grad_xy[p] for VISIBLE points (any slot of a pixel with idx[...,0] >= 0, rasterizer.py:851-857):
      sum over pixels of the same view with grad_occ != 0, |pixel - p|^2 <= r_n^2,
      r_n = median(radii of the view's visible points) * radii_s (:884), skipping z < 0 /
      |x|,|y| > 1 points and, when grad_occ > 0, pixels outside the splat's radii box:
          g * d / eps_denom(|d|^2, 1e-10)      (the kernel's eps_denom has sign(0) = 0)
    grad_z: z_grad[idx] += grad_zbuf, skipping zeros, stopping at the first -1
    (rasterize_points.cu:835-843).
    The reference additionally drops/duplicates the last 2-D grid cell of views n >= 1 (a
    packed-vs-local offset bug, rasterize_points_backward.cu:124-126); that is NOT reproduced.
    Returns grad (P,3) float64 and the per-view search radii.
    """
    points = np.asarray(points, np.float32)
    radii = np.asarray(radii, np.float32)
    P = points.shape[0]
    N, H, W, K = idx.shape
    grad = np.zeros((P, 3), np.float64)
    vis = np.zeros(P, bool)
    m0 = idx[..., 0] >= 0
    v = idx[m0]
    vis[v[v >= 0]] = True
    rs = np.zeros(N, np.float32)
    ndc_x = pix_to_ndc(np.arange(W), W)
    ndc_y = pix_to_ndc(np.arange(H), H)
    for n in range(N):
        s0 = int(first_idx[n])
        s1 = s0 + int(num_points[n])
        pv = np.nonzero(vis[s0:s1])[0] + s0
        if pv.size == 0:
            continue
        rv = np.sort(radii[pv].reshape(-1))
        rs[n] = f32(rv[(rv.size - 1) // 2]) * f32(radii_s)   # torch.median = lower middle
        r2 = f32(rs[n] * rs[n])
        gy, gx = np.nonzero(grad_occ[n] != 0)
        if gy.size == 0:
            continue
        g = grad_occ[n, gy, gx].astype(np.float32)
        xf = ndc_x[W - 1 - gx]
        yf = ndc_y[H - 1 - gy]
        for p in pv:
            px, py, pz = points[p]
            if pz < 0 or abs(py) > 1.0 or abs(px) > 1.0:
                continue
            dx = (xf - px).astype(np.float32)
            dy = (yf - py).astype(np.float32)
            d2 = _fma(dx, dx, (dy * dy).astype(np.float32))   # SASS: FMUL dy*dy ; FFMA dx
            sel = ~(d2 > r2)
            outside = (np.abs(dx) > radii[p, 0]) | (np.abs(dy) > radii[p, 1])
            sel &= ~((g > 0) & outside)
            if not sel.any():
                continue
            sgn = np.sign(d2[sel]).astype(np.float64)
            den = sgn * np.maximum(np.abs(d2[sel]), f32(1e-10)).astype(np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                grad[p, 0] += np.sum(dx[sel].astype(np.float64) / den * g[sel])
                grad[p, 1] += np.sum(dy[sel].astype(np.float64) / den * g[sel])
    # z
    flat_idx = idx.reshape(-1, K)
    flat_g = grad_zbuf.reshape(-1, K)
    alive = np.ones(flat_idx.shape[0], bool)
    for k in range(K):
        gk = flat_g[:, k]
        nz = gk != 0
        neg = flat_idx[:, k] < 0
        use = alive & nz & ~neg
        np.add.at(grad[:, 2], flat_idx[use, k], gk[use].astype(np.float64))
        alive &= ~(nz & neg)      # `break` only triggers when the gradient is non-zero
    return grad, rs


# =========================================================================================
# point-set operators  (DSS/utils/point_processing.py)
# =========================================================================================
def knn_bruteforce(p1: torch.Tensor, p2: torch.Tensor, K: int):
    """Exact K nearest neighbours of (P1,3) in (P2,3): squared dists ascending, idx (float64 ranking)."""
    d = torch.cdist(p1.double(), p2.double()) ** 2
    v, i = torch.topk(d, min(K, p2.shape[0]), dim=1, largest=False)
    return v.float(), i


def _gather0(x, idx):
    """frnn_gather for one cloud: rows of x at idx, zeros where idx < 0 (frnn.py:340-352)."""
    return x[idx.clamp_min(0)] * (idx >= 0)[..., None]


def wlop(P: torch.Tensor, noise: torch.Tensor, neighborhood_size=16, iters=3, repulsion_mu=0.5, frnn_fn=None,
         X0: torch.Tensor = None):
    """wlop for one cloud (P,3) (point_processing.py:35-122); ``X0`` = the farthest-point subsample for
    ratio < 1 (:51-52), P itself for ratio = 1.0: X0 = X0 + noise * 0.1 h,
    h = 4 sqrt(diag/n), theta(r2) = exp(-16 r2 / h^2), search radius min(h K, 0.2);
    X <- sum alpha p / sum alpha + mu sum beta delta / sum beta with
    alpha = theta(|eps|^2) / |eps| / density_P[j],  beta = density_X theta(|delta|^2) / |delta|."""
    n = P.shape[0]
    K = neighborhood_size
    diag = float((P.max(0).values - P.min(0).values).norm())
    h = 4 * math.sqrt(diag / n)
    r = min(h * K, 0.2)
    s_inv = 16 / h / h
    fn = frnn_fn or (lambda a, b, K, r: torch.as_tensor(frnn_bruteforce(a[None].numpy(), b[None].numpy(), K=K, r=r)[0][0]))
    theta = lambda r2: torch.exp(-r2 * s_inv)
    X = (P if X0 is None else X0) + noise * h * 0.1
    idx_pp = fn(P, P, K + 1, r)[:, 1:]
    dpp = (P[:, None] - _gather0(P, idx_pp)).norm(dim=-1)
    th = theta(dpp ** 2) * (idx_pp >= 0)
    density_P = th.sum(-1) + 1
    for _ in range(iters):
        idx_xp = fn(X, P, K, r)
        idx_xx = fn(X, X, K + 1, r)[:, 1:]
        nn_xp = _gather0(P, idx_xp)
        eps = X[:, None] - nn_xp
        delta = X[:, None] - _gather0(X, idx_xx)
        dxx2 = (delta ** 2).sum(-1)
        dxp2 = (eps ** 2).sum(-1)
        alpha = theta(dxp2) / eps_denom(eps.norm(dim=-1))
        beta = theta(dxx2) / eps_denom(delta.norm(dim=-1))
        density_X = theta(dxx2).sum(-1) + 1
        na = alpha / _gather0(density_P[:, None], idx_xp).squeeze(-1)
        na = torch.where(idx_xp < 0, torch.zeros_like(na), na)
        nb = density_X[:, None] * beta
        nb = torch.where(idx_xx < 0, torch.zeros_like(nb), nb)
        X = (na[..., None] * nn_xp).sum(-2) / eps_denom(na.sum(-1, keepdim=True)) + \
            repulsion_mu * (nb[..., None] * delta).sum(-2) / eps_denom(nb.sum(-1, keepdim=True))
    return X


def upsample(points: torch.Tensor, n_points: int, neighborhood_size=16):
    """upsample for one cloud (P,3) -> (n_points,3) (point_processing.py:281-362): per round, every
    point proposes the mid-point (nn_k + 2p)/3 that is farthest from all its K neighbours; the P//8
    sparsest proposals (at most the remaining count) are PREPENDED; K-NN is rebuilt each round."""
    K = neighborhood_size
    pts = points
    remaining = n_points - pts.shape[0]
    while remaining > 0:
        Pn = pts.shape[0]
        max_P = Pn // 8
        _, idx = knn_bruteforce(pts, pts, K + 1)
        nn = pts[idx[:, 1:]]                                   # (P,K,3)
        mid = (nn + 2 * pts[:, None]) / 3
        dist = (mid[:, :, None] - nn[:, None]).norm(dim=-1)     # (P,K,K)
        sparsity, father = dist.min(-1).values.max(-1)
        order = sparsity.sort().indices[Pn - max_P:]
        n_new = min(remaining, max_P)
        new = mid[torch.arange(Pn), father][order]
        pts = torch.cat([new[max_P - n_new:], pts], 0)
        remaining -= n_new
        if max_P == 0:
            break
    return pts


def fps(points: torch.Tensor, m: int, start: int = 0):
    """Farthest point sampling of one cloud (P,3) to m indices in selection order -- what
    ``torch_cluster.fps`` computes for ``farthest_sampling`` (point_processing.py:473-499) [third party,
    restated from its documentation; upstream draws the start point at random, here it is ``start``]:
    repeatedly the point with the largest squared distance to the selected set (first arg-max)."""
    sel = [start]
    md = ((points - points[start]) ** 2).sum(-1)
    for _ in range(m - 1):
        j = int(torch.argmax(md))
        sel.append(j)
        md = torch.minimum(md, ((points - points[j]) ** 2).sum(-1))
    return torch.tensor(sel, dtype=torch.int64)


def resample_uniformly(P: torch.Tensor, noise: torch.Tensor, shrink_ratio=0.5, repulsion_mu=1.0, frnn_fn=None):
    """resample_uniformly for one cloud (point_processing.py:126-166): wlop(ratio=shrink_ratio, repulsion_mu)
    (:160; defaults neighborhood_size=16, iters=3) then upsample back to the input count (:161).  The K-NN and
    normals of :141-158 are computed and never used."""
    n = P.shape[0]
    sub = fps(P, int(math.ceil(n * shrink_ratio)))
    X = wlop(P, noise, repulsion_mu=repulsion_mu, frnn_fn=frnn_fn, X0=P[sub])
    return upsample(X, n)


def insert(ref_points: torch.Tensor, metrics: torch.Tensor, points: torch.Tensor, frnn_fn=None):
    """UniformProjection.insert for one cloud (levelset_sampling.py:172-233): children 2f/3 + m/3 of every
    "father" f = a point within 2 * avg_spacing of a high-saliency reference point (and not exactly on it),
    m = its 8 FRNN neighbours (zeros where it has fewer, like frnn_gather).  ``metrics`` (R,1) is the
    reference cloud's feature; salient = above min(2 median, max / 2) (:189), replaced by the top
    max(min(50, R // 20), 1) when that set is empty or larger than min(50, R // 20) (:195-197).
    Returns (child (C,3), n_children)."""
    P = points.shape[0]
    diag = float((points.max(0).values - points.min(0).values).norm())
    avg_spacing = math.sqrt(diag / ref_points.shape[0])
    patch = 8
    r = min(avg_spacing * patch, 0.2)
    fn = frnn_fn or (lambda a, b, K, r: frnn_bruteforce(a[None].numpy(), b[None].numpy(), K=K, r=r))
    idx, _ = fn(points, points, patch + 1, r)
    idx = torch.as_tensor(idx[0])[:, 1:]
    R = metrics.shape[0]
    threshold = min(metrics.median() * 2, metrics.max() * 0.5)
    ref = ref_points[(metrics > threshold).squeeze(-1)]
    cap = min(50, int(R / 20))
    if ref.shape[0] == 0 or ref.shape[0] > cap:
        ref = ref_points[metrics.sort(dim=0).indices[-max(cap, 1):, 0]]
    _, d = fn(points, ref, 1, r * 4)
    d = torch.as_tensor(d[0]).reshape(P)
    father = (d < 4 * avg_spacing ** 2) & (d > 0)
    mother = _gather0(points, idx[:, -patch:])[father]
    child = 2 * points[father].unsqueeze(-2) / 3 + mother / 3
    return child.reshape(-1, 3), int(father.sum()) * patch


# =========================================================================================
# EWA per-point splat parameters + renderable filter  (DSS/core/rasterizer.py:124-255, 344-563)
# =========================================================================================
def _to_packed_views(data: torch.Tensor, first_idx, num_points):
    """gather_batch_to_packed (DSS/utils/__init__.py:218-250): per-view data -> per packed point;
    a single-view tensor is broadcast."""
    if data.shape[0] == 1:
        return data.expand(int(sum(int(n) for n in num_points)), *data.shape[1:])
    b = torch.repeat_interleave(torch.arange(len(num_points)), torch.as_tensor(num_points, dtype=torch.int64))
    return data[b]


def ewa_vrk_h(sq_dists: torch.Tensor, num_points) -> torch.Tensor:
    """_compute_isotropic_Vrk, first half (rasterizer.py:358-386).  sq_dists (N, P1, K) padded result of
    the K = 7 self query (slot 0 = the point itself); returns packed h_k (P,)."""
    d = sq_dists[:, :, 1:].clone()
    n = torch.as_tensor(num_points, dtype=torch.int64)
    d[n < sq_dists.shape[2]] = 1e-3                                     # :376
    packed = torch.cat([d[b, :int(n[b])] for b in range(d.shape[0])], 0)  # padded_to_packed :378-380
    h = 0.5 * packed.max(dim=-1)[0]                                     # :382
    return h.clamp(5e-5, 0.01)                                          # :385


def ewa_point_params(points: torch.Tensor, normals: torch.Tensor, first_idx, num_points, proj: torch.Tensor,
                     vrk_h: torch.Tensor, image_size: int, antialiasing_sigma: float, cutoff: float,
                     rand: torch.Tensor = None, dtype=torch.float64):
    """_get_per_point_info (rasterizer.py:514-563) with the isotropic V_k^r (:388-400), operation by
    operation, in `dtype` (float64 = the checker; float32 = the reference's own arithmetic).
    points / normals packed (P,3); proj (N or 1, 4, 4) = get_full_projection_transform().get_matrix().
    `rand` (P,3) stands in for torch.rand_like (:393); the results do not depend on it beyond rounding.
    Returns radii (P,2), ellipse (P,3), cutoff (P,), scaler (P,), det_mk (P,)."""
    import torch.nn.functional as F
    pts = points.to(dtype)
    nrm = normals.to(dtype)
    M44 = proj.to(dtype)
    P = pts.shape[0]
    if rand is None:
        rand = torch.rand(P, 3, generator=torch.Generator().manual_seed(0)).to(dtype)
    # ---- _compute_WJk (:438-487) ----
    W = _to_packed_views(M44[..., :3, :], first_idx, num_points)               # (P,3,4)
    hom = torch.cat([pts, torch.ones(P, 1, dtype=dtype)], -1)                    # to_homogen
    denom = (hom[:, None, :] @ _to_packed_views(M44[..., 3:], first_idx, num_points)).view(-1)
    denom_sqr = eps_denom(denom ** 2)
    Jk = torch.zeros(P, 4, 2, dtype=dtype)
    denom = eps_denom(denom)
    Jk[:, 0, 0] = 1 / denom
    Jk[:, 1, 1] = 1 / denom
    xy_view = hom[:, None, :] @ _to_packed_views(M44[..., :2], first_idx, num_points)   # (P,1,2)
    Jk[:, 3, 0] = -1 / denom_sqr * xy_view[:, :, 0].view(-1)
    Jk[:, 3, 1] = -1 / denom_sqr * xy_view[:, :, 1].view(-1)
    WJk = W @ Jk                                                                 # (P,3,2)
    # ---- _compute_isotropic_Vrk, second half (:388-400) ----
    u0 = F.normalize(torch.cross(nrm, nrm + rand.to(dtype), dim=-1), dim=-1)
    u1 = F.normalize(torch.cross(nrm, u0, dim=-1), dim=-1)
    Sk = torch.stack([u0, u1], dim=1)                                            # (P,2,3)
    Vrk = vrk_h.to(dtype).view(-1, 1, 1) * Sk.transpose(1, 2) @ Sk
    # ---- _compute_variance_and_detMk (:423-436) ----
    Mk = Sk @ WJk
    Vk = WJk.transpose(1, 2) @ Vrk @ WJk
    pixel_size = 2.0 / image_size
    variance = Vk + antialiasing_sigma * torch.eye(2, dtype=dtype).expand(P, 2, 2) * (pixel_size ** 2)
    det_mk = torch.det(Mk)
    # ---- _get_per_point_info (:526-557) ----
    gv_det = torch.det(variance)
    gv_inv = torch.inverse(variance)
    ellipse = torch.stack([gv_inv[:, 0, 0], gv_inv[:, 0, 1] + gv_inv[:, 1, 0], gv_inv[:, 1, 1]], -1)
    a, b, c = ellipse[:, 0], ellipse[:, 1], ellipse[:, 2]                        # :502-511
    den = eps_denom(4 * a * c - b ** 2)
    y = torch.sqrt((4 * a * cutoff / den).abs().clamp_min(1e-17))
    x = torch.sqrt((4 * c * cutoff / den).abs().clamp_min(1e-17))
    radii = torch.stack([x, y], -1)
    scaler = torch.sqrt((gv_det * 4 * np.pi * np.pi).abs().clamp_min(1e-17))
    scaler = det_mk.abs() / eps_denom(scaler)
    return radii, ellipse, torch.full_like(a, cutoff), scaler, det_mk


def renderable_mask(points: torch.Tensor, normals, first_idx, num_points, w2v: torch.Tensor,
                    nmat: torch.Tensor = None, znear: float = 1.0, zfar: float = 100.0):
    """filter_renderable (rasterizer.py:220-255) as one packed mask: _filter_points_with_invalid_depth
    (:163-218: znear <= z_view <= zfar, z_view from Transform3d.transform_points = p_hom @ w2v, xyz / w)
    and, when `nmat` is given, _filter_backface_points (:124-161: (normals @ nmat).z < 0, nmat =
    inverse(w2v)[:3,:3]^T as pytorch3d's Transform3d.transform_normals builds it).  float32."""
    P = points.shape[0]
    hom = torch.cat([points, torch.ones(P, 1, dtype=points.dtype)], -1)
    V = _to_packed_views(w2v, first_idx, num_points)
    out = (hom[:, None, :] @ V)[:, 0]
    zv = out[:, 2] / out[:, 3]
    mask = (zv >= znear) & (zv <= zfar)
    if nmat is not None:
        nv = (normals[:, None, :] @ _to_packed_views(nmat, first_idx, num_points))[:, 0]
        mask = mask & (nv[:, 2] < 0)
    n = [int(x) for x in num_points]
    kept = [int(mask[int(f):int(f) + c].sum()) for f, c in zip(first_idx, n)]
    return mask, kept


# =========================================================================================
# SphereTracing.project_points  (DSS/models/levelset_sampling.py:679-808)
# =========================================================================================
def sphere_trace(model, ray0: torch.Tensor, ray_direction: torch.Tensor, proj_max_iters=10, proj_tolerance=5e-5,
                 alpha=1.0, radius=1.0, padding=0.1, **kw):
    """Packed rays (M,3).  Per iteration (:733-786): evaluate the SDF (+ gradient) at the active rays;
    active = |sdf| > 0.1 tol and still inside the sphere; active rays advance by alpha * sdf * dir with
    the step length clamped to 0.1; a ray whose new position leaves |x| < radius + padding keeps its old
    position and retires.  Returns points (M,3), last sdf (M,), last gradient (M,3), mask = |sdf| <= tol."""
    pts = ray0.clone()
    M = pts.shape[0]
    active = torch.ones(M, dtype=torch.bool)
    inside = torch.ones(M, dtype=torch.bool)
    sdf = torch.zeros(M)
    grad = torch.zeros(M, 3)
    trials = 0
    while True:
        if bool(active.any()):
            s, g = sdf_and_grad(pts[active], model, **kw)
            sdf[active] = s.reshape(-1)
            grad[active] = g
        active = (sdf.abs() > 1e-1 * proj_tolerance) & inside                      # :760-761
        if not (bool(active.any()) and trials < proj_max_iters):                     # :764
            break
        move = alpha * sdf[active, None] * ray_direction[active]                    # :769-770
        nrm = move.norm(dim=-1, keepdim=True)
        move = move / nrm.clamp_min(1e-15) * nrm.clamp_max(0.1)                     # :771-772
        new = pts[active] + move
        ok = new.norm(dim=-1) < (padding + radius)                                  # :774-775
        idx = torch.nonzero(active).reshape(-1)
        inside[idx] = ok
        pts[idx[ok]] = new[ok]                                                      # :776-777
        trials += 1
    return pts, sdf, grad, sdf.abs() <= proj_tolerance


# =========================================================================================
# in-surface sampler: closest iso-point to a ray  (DSS/models/combined_modeling.py:326-352)
# =========================================================================================
def ray_nearest_point(cam_pos: torch.Tensor, ray0: torch.Tensor, points: torch.Tensor):
    """The dense computation of combined_modeling.py:331-347 for one view, in the reference's own op sequence
    (float32, torch-CPU): pC = p - C; ray_sq = ((pC * ray0).sum(-1))^2 (R,M); dist_to_ray = |pC|^2 - ray_sq;
    the smallest dist per ray.  Returns (ray_sq at the arg-min (R,), idx (R,), dist_to_ray (R,M), ray_sq (R,M))."""
    pC = points - cam_pos.view(1, 3)
    ray_sq = (pC[None, :, :] * ray0[:, None, :]).sum(-1) ** 2
    dist_to_ray = (pC ** 2).sum(-1).unsqueeze(0) - ray_sq
    _, nn_idx = torch.topk(dist_to_ray, k=1, dim=1, largest=False)
    return torch.gather(ray_sq, 1, nn_idx).view(-1), nn_idx.view(-1), dist_to_ray, ray_sq
