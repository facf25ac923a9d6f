"""Iso-point projection / resampling operators on libisob200.so.

Mirrors the operator surface of DSS/models/levelset_sampling.py (``LevelSetProjection``,
``UniformProjection``, ``EdgeAwareProjection``, ``sample_uniform_iso_points``,
``ProjectionResult``): same constructor arguments, method names, argument meaning (including the
``x or default`` idiom, :304-305, :377-378), return containers and early-exit behaviour, so the
reference's callers (combined_modeling.py:449, implicit_modeling.py:156, trainer.py:227) can
switch imports and run unchanged.

Everything between two SDF evaluations runs in hand-written sm_100a kernels:
  * one fused Newton-step + in-order compaction kernel per iteration (csrc/project.cu)
    instead of ~15 boolean-mask / elementwise PyTorch kernels with 2 host syncs;
  * one resample kernel per sample_iter (csrc/resample.cu) instead of two (N,P,K,3)
    ``frnn_gather`` materialisations + ~20 elementwise kernels;
  * the FRNN grid build / query kernels (csrc/frnn_*.cu).
The SDF itself stays the caller's opaque ``nn.Module`` (evaluated through autograd exactly like
levelset_sampling.py:142-170).  There is no CPU / PyTorch fallback for the kernels.
"""
from collections import namedtuple
import math
import os
from typing import Callable, Optional, Tuple, Union

import torch
import torch.autograd as autograd
import torch.nn.functional as F

from . import _ext
from . import frnn
from . import siren
from .structures import (convert_pointclouds_to_tensor, is_pointclouds, packed_to_padded,
                         padded_to_packed_idx, reduce_mask_padded,
                         num_points_2_cloud_to_packed_first_idx)

# one cloud + fused SIREN: enqueue filter + resample without reading the survivor count back first
# (UniformProjection._filter_resample_run_ahead); off = the read-back right after the projection
RUN_AHEAD = os.environ.get("ISOB200_RUN_AHEAD", "1") != "0"

ProjectionResult = namedtuple('ProjectionResult', ('points', 'normals', 'mask'))
_KNN = namedtuple("KNN", "dists idx knn")  # pytorch3d.ops.knn._KNN


def _filter_projection_result(result: ProjectionResult) -> ProjectionResult:
    """levelset_sampling.py:59-65."""
    points, normals, mask = result
    points = reduce_mask_padded(points, mask)
    normals = reduce_mask_padded(normals, mask)
    mask = reduce_mask_padded(mask, mask)
    return ProjectionResult(points, normals, mask)


def _filter_projection_result_counted(result: ProjectionResult):
    """``_filter_projection_result`` plus the survivor counts as a host list and as a device int64
    tensor, with ONE read-back for all of it (the reference pays one in ``.any()`` and more inside
    the boolean-mask indexing)."""
    points, normals, mask = result
    if mask.shape[0] == 1 and points.is_cuda and points.dtype == torch.float32:
        # single cloud: one order-preserving compaction kernel; the count it leaves on the device is the
        # output shape and the only thing read back
        lib = _ext.lib()
        dev = points.device
        M = int(mask.shape[1])
        pts, nrm = points.reshape(-1, 3).contiguous(), normals.reshape(-1, 3).contiguous()
        valid = mask.reshape(-1).contiguous().view(torch.uint8)
        out_p, out_n = torch.empty_like(pts), torch.empty_like(nrm)
        count = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws = _ext.workspace(lib.isob200_project_step_ws_bytes(max(M, 1)), dev)
        _ext.check(lib.isob200_compact_valid(_ext.ptr(pts), _ext.ptr(nrm), _ext.ptr(valid), M, _ext.ptr(out_p),
                                             _ext.ptr(out_n), _ext.ptr(count), _ext.ptr(ws), ws.numel(),
                                             _ext.stream(dev)))
        n = int(count.item())
        return ProjectionResult(out_p[:n][None], out_n[:n][None], mask.new_ones((1, n))), [n], count.long()
    num = mask.sum(dim=-1)
    return _filter_projection_result(result), [int(c) for c in num.tolist()], num


class LevelSetProjection(object):
    """levelset_sampling.py:67-76."""

    def __init__(self, proj_max_iters=10, proj_tolerance=5.0e-5, max_points_per_pass=120000):
        self.proj_max_iters = proj_max_iters
        self.proj_tolerance = proj_tolerance
        self.max_points_per_pass = max_points_per_pass

    def project_points(self, points_init, network, latent, levelset):
        raise NotImplementedError


class UniformProjection(LevelSetProjection):
    """Project points onto the zero level set and spread them uniformly (levelset_sampling.py:79-439)."""

    def __init__(self, proj_max_iters=10, proj_tolerance=5e-5, max_points_per_pass=120000,
                 sample_iters=1, knn_k=8, resampling_clip=0.02, **kwargs):
        super().__init__(proj_max_iters=proj_max_iters, proj_tolerance=proj_tolerance,
                         max_points_per_pass=max_points_per_pass)
        self.knn_k = knn_k
        self.sample_iters = sample_iters
        self.resampling_clip = resampling_clip
        self._knn_idx = None
        self._knn_dists = None
        self._knn_nn_cache = None
        self._knn_src = None

    # `_knn_nn` is only materialised if somebody reads it (the resample kernel gathers in registers)
    @property
    def _knn_nn(self):
        if self._knn_nn_cache is None and self._knn_idx is not None and self._knn_src is not None:
            self._knn_nn_cache = frnn.frnn_gather(self._knn_src, self._knn_idx)
        return self._knn_nn_cache

    def _create_tree(self, points_padded: torch.Tensor, refresh_tree=True, num_points_per_cloud=None):
        """levelset_sampling.py:110-140: FRNN with K = knn_k + 1, r = sqrt(diag / n) * knn_k; the
        first column is dropped as "self"."""
        if not refresh_tree and getattr(self, '_knn_idx', None) is not None:
            return self._knn_idx
        assert (points_padded.ndim == 3)
        if num_points_per_cloud is None:
            num_points_per_cloud = torch.tensor([points_padded.shape[1]] * points_padded.shape[0],
                                                device=points_padded.device, dtype=torch.long)
        diag = self.__dict__.pop("_diag_hint", None)   # resample() just computed it for this very tensor
        if diag is None:
            mn, mx = torch.aminmax(points_padded, dim=1)
            diag = (mx - mn).norm(dim=-1)
        search_radius = torch.sqrt(diag / num_points_per_cloud.float()) * self.knn_k
        # r = knn_k estimated point spacings: ~knn_k^2 neighbours inside it on a surface, the K-th one far inside
        hint, frnn.FAR_RADIUS_HINT = frnn.FAR_RADIUS_HINT, True
        try:
            dists, idxs, _, grid = frnn.frnn_grid_points(
                points_padded, points_padded, num_points_per_cloud, num_points_per_cloud,
                K=self.knn_k + 1, r=search_radius, grid=None, return_nn=False)
        finally:
            frnn.FAR_RADIUS_HINT = hint
        self._knn_gather = frnn.frnn_gather
        self._knn_full_idx = idxs
        self._knn_idx = idxs[..., 1:]
        self._knn_dists = dists[..., 1:]
        self._knn_src = points_padded
        self._knn_nn_cache = None
        self.knn_gather = frnn.frnn_gather
        return self._knn_idx

    def _compute_sdf_and_grad(self, points, model, latent=None, **forward_kwargs) -> Tuple[torch.Tensor]:
        """Chunked SDF value + input gradient through autograd (levelset_sampling.py:142-170)."""
        shp = points.shape
        points_packed = points.reshape(-1, 3)
        if points_packed.is_cuda:
            # the reference's own Siren decoder (common.py:90-165): one fused tcgen05 kernel
            fused = None if latent is not None else siren.sdf_and_grad(model, points_packed, forward_kwargs)
            if fused is not None:
                return fused[0].view(shp[:-1]), fused[1].view(shp)
        grad_packed = []
        eval_packed = []
        with autograd.no_grad():
            model.eval()
            for sub_points in torch.split(points_packed, self.max_points_per_pass, dim=0):
                with autograd.enable_grad():
                    net_input = sub_points.detach().requires_grad_(True)
                    network_eval = model.forward(net_input, **forward_kwargs).sdf
                    input_grad = autograd.grad([network_eval], [net_input], torch.ones_like(network_eval),
                                               retain_graph=False)[0]
                grad_packed.append(input_grad)
                eval_packed.append(network_eval.detach())
            if len(grad_packed) == 1:
                grad_packed, eval_packed = grad_packed[0], eval_packed[0]
            else:
                grad_packed = torch.cat(grad_packed, dim=0)
                eval_packed = torch.cat(eval_packed, dim=0)
            grad_packed = grad_packed.view(shp)
            eval_packed = eval_packed.view(shp[:-1])
        return eval_packed, grad_packed

    # ------------------------------------------------------------------------------------
    def _project_points(self, model: Callable, points: torch.Tensor, num_points: torch.Tensor,
                        proj_max_iters: int = None, proj_tolerance: float = None, num_points_list=None,
                        _live=None, **forward_kwargs) -> ProjectionResult:
        """Newton projection of the live rows of ``points`` (B,P,3) (levelset_sampling.py:290-351).

        Returns padded points (B,Pmax,3), the last SDF gradient as normals (B,Pmax,3) and the
        converged mask (B,Pmax) bool.  Non-converged points keep their last position.
        ``num_points_list``: the host copy of ``num_points`` when the caller already has it
        (saves one read-back).  ``_live`` (internal; one cloud, the fused SIREN path): an int32 device scalar --
        only the first ``_live`` rows are projected, the rest come back unconverged; the host never learns
        the count here.
        """
        proj_max_iters = proj_max_iters or self.proj_max_iters
        proj_tolerance = proj_tolerance or self.proj_tolerance
        _ext.require_cuda(points)
        lib = _ext.lib()
        dev = points.device
        B, P = points.shape[0], points.shape[1]
        points = points.contiguous()
        if points.dtype != torch.float32:
            raise RuntimeError("expected scalar type Float")
        if _live is not None:
            assert B == 1 and _live.dtype == torch.int32
            num_points_list = [P]
        num_list = [int(x) for x in (num_points_list if num_points_list is not None else num_points.tolist())]
        full = all(n == P for n in num_list)
        if full:
            points_packed = points.reshape(-1, 3).clone()
        else:
            live, _ = padded_to_packed_idx(num_points, P)
            points_packed = points.reshape(-1, 3)[live].contiguous()
        M = points_packed.shape[0]
        normals_packed = torch.zeros_like(points_packed)
        not_converged = torch.ones((M,), dtype=torch.bool, device=dev)

        analytic = getattr(model, "isob200_analytic_sdf", None)
        fused = siren.match(model, forward_kwargs) if M > 0 else None
        if _live is not None and (fused is None or analytic is not None):
            raise RuntimeError("a device-side row count needs the fused SIREN path")
        if analytic is not None and not forward_kwargs and M > 0:
            kind, radius = analytic
            if kind != "sphere":
                raise ValueError("unknown built-in SDF %r" % (kind,))
            valid_u8 = torch.empty((M,), dtype=torch.uint8, device=dev)
            _ext.check(lib.isob200_project_sphere(
                _ext.ptr(points_packed), _ext.ptr(normals_packed), _ext.ptr(valid_u8), M, float(radius),
                float(proj_tolerance), 0.1, int(proj_max_iters), _ext.stream(dev)))
            valid_packed = valid_u8.bool()
        elif fused is not None:
            # The reference's Siren decoder: one kernel per Newton iteration -- fused tcgen05 SDF +
            # gradient, the update of :333-342 and the compaction of the still-active rows -- reading
            # the live row count from device memory: the whole loop is enqueued without a single
            # read-back (the early exit of :329 becomes launches over an empty active set).
            act = (torch.empty((M,), dtype=torch.int32, device=dev), torch.empty((M,), dtype=torch.int32, device=dev))
            nxt = (torch.empty((M, 3), dtype=torch.float32, device=dev),
                   torch.empty((M, 3), dtype=torch.float32, device=dev))
            cnt = torch.zeros((proj_max_iters + 2,), dtype=torch.int32, device=dev)  # live rows entering iteration it
            nc_u8 = not_converged.view(torch.uint8)
            st = _ext.stream(dev)
            blob, scratch, n_hidden = siren.packed(model, fused)   # operand images, current parameter version
            for it in range(proj_max_iters + 1):
                last = (it == proj_max_iters)
                cur = points_packed if it == 0 else nxt[it & 1]
                _ext.check(lib.isob200_siren_project_step(
                    _ext.ptr(cur), M, _ext.ptr(_live) if it == 0 else _ext.ptr(cnt[it:]), _ext.ptr(blob), n_hidden,
                    _ext.ptr(scratch), scratch.numel(), _ext.ptr(points_packed), _ext.ptr(normals_packed),
                    _ext.ptr(nc_u8), None if it == 0 else _ext.ptr(act[it & 1]), float(proj_tolerance), 0.1,
                    0 if last else 1, _ext.ptr(act[(it + 1) & 1]),
                    None if last else _ext.ptr(nxt[(it + 1) & 1]), _ext.ptr(cnt[it + 1:]), st))
            siren.STATS["calls"] += proj_max_iters + 1
            if _live is None:
                siren.STATS["rows"] += M
            if siren.RECORD is not None:   # bench.py: live row counts, resolved after the timed region
                siren.RECORD.append((M if _live is None else _live, cnt))
            valid_packed = ~not_converged
        else:
            if M > 0:
                act_a = torch.empty((M,), dtype=torch.int32, device=dev)
                act_b = torch.empty((M,), dtype=torch.int32, device=dev)
                count = torch.zeros((1,), dtype=torch.int32, device=dev)
                ws = _ext.workspace(lib.isob200_project_step_ws_bytes(M), dev)
                nc_u8 = not_converged.view(torch.uint8)
                act_in, act_out = None, act_a
                nxt_a = torch.empty((M, 3), dtype=torch.float32, device=dev)   # compacted next SDF inputs
                nxt_b = torch.empty((M, 3), dtype=torch.float32, device=dev)
                nxt_in, nxt_out = None, nxt_a
                A = M
                it = 0
                while True:
                    curr_points = points_packed if act_in is None else nxt_in[:A]
                    curr_sdf, curr_grad = self._compute_sdf_and_grad(curr_points, model, **forward_kwargs)
                    curr_sdf = curr_sdf.reshape(-1).contiguous().float()
                    curr_grad = curr_grad.reshape(-1, 3).contiguous().float()
                    last = (it == proj_max_iters)
                    _ext.check(lib.isob200_project_step(
                        _ext.ptr(points_packed), _ext.ptr(normals_packed), _ext.ptr(nc_u8),
                        _ext.ptr(act_in), A, None, _ext.ptr(curr_sdf), _ext.ptr(curr_grad),
                        float(proj_tolerance), 0.1, 0 if last else 1, _ext.ptr(act_out),
                        None if last else _ext.ptr(nxt_out), _ext.ptr(count), _ext.ptr(ws), ws.numel(),
                        _ext.stream(dev)))
                    if last:
                        break
                    A = int(count.item())  # the opaque SDF callback needs the batch shape
                    if A == 0:
                        break
                    it += 1
                    act_in = act_out
                    act_out = act_b if act_in is act_a else act_a
                    nxt_in = nxt_out
                    nxt_out = nxt_b if nxt_in is nxt_a else nxt_a
            valid_packed = ~not_converged

        if full:
            return ProjectionResult(points_packed.view(B, P, 3), normals_packed.view(B, P, 3),
                                    valid_packed.view(B, P))
        first = num_points_2_cloud_to_packed_first_idx(num_points)
        pmax = max(num_list) if num_list else 0
        points_out = packed_to_padded(points_packed, first, pmax)
        normals_out = packed_to_padded(normals_packed, first, pmax)
        valid_mask = packed_to_padded(valid_packed.view(-1, 1).float(), first, pmax).squeeze(-1).bool()
        return ProjectionResult(points_out, normals_out, valid_mask)

    # ------------------------------------------------------------------------------------
    def resample(self, model, points_init, normals_init, num_points, sample_iters=None,
                 num_points_list=None, _live=None, **forward_kwargs) -> ProjectionResult:
        """Repulse along neighbours' tangent planes then re-project, ``sample_iters`` times
        (levelset_sampling.py:239-288)."""
        sample_iters = sample_iters or self.sample_iters
        batch_size = points_init.shape[0]
        if num_points is None:
            num_points = torch.full((batch_size,), points_init.shape[1], dtype=torch.long,
                                    device=points_init.device)
        if sample_iters == 0:
            return ProjectionResult(points_init, normals_init,
                                    points_init.new_full(points_init.shape[:-1], True, dtype=torch.bool))
        if _live is None and points_init.nelement() < 2 * (self.knn_k + 1):
            return ProjectionResult(points_init, normals_init,
                                    points_init.new_full(points_init.shape[:-1], True, dtype=torch.bool))
        _ext.require_cuda(points_init)
        lib = _ext.lib()
        dev = points_init.device
        flat = points_init.reshape(-1, 3)
        if _live is None:
            mn, mx = torch.aminmax(flat, dim=0)
            diag = (mx - mn).norm()  # stays on the device (:254)
        else:
            # one cloud whose first `_live` rows count (device scalar; `num_points` is its int64 copy): same
            # bounding box from a kernel that reads the count on the device; an empty cloud gets diag = 1 so
            # that the speculative grid stays finite (the caller drops the result when it learns the count)
            box = torch.empty((1, 2, 3), dtype=torch.float32, device=dev)
            ws = _ext.workspace(64, dev)
            _ext.check(lib.isob200_points_bbox(_ext.ptr(points_init.contiguous()), _ext.ptr(num_points), 1,
                                               int(points_init.shape[1]), 3, _ext.ptr(box), _ext.ptr(ws), ws.numel(),
                                               _ext.stream(dev)))
            diag = (box[0, 1] - box[0, 0]).norm()
            diag = torch.where(_live[0] > 0, diag, torch.ones_like(diag))
        inv_sigma_spatial = (num_points.float() / diag).contiguous()  # (B,)

        points = points_init.contiguous()
        B, P = points.shape[0], points.shape[1]
        normals = torch.empty_like(points)
        _ext.check(lib.isob200_normalize_rows3(_ext.ptr(normals_init.contiguous()), B * P, 1e-12,
                                               _ext.ptr(normals), _ext.stream(dev)))
        projection_result = None
        for sample_iter in range(sample_iters):
            if sample_iter % 2 == 0:
                # assume repulsion doesn't change neighborhood
                if sample_iter == 0 and B == 1 and type(self)._create_tree is UniformProjection._create_tree:
                    self._diag_hint = diag.reshape(1)   # same tensor, same bounding box (:129 == :254 for B = 1)
                self._create_tree(points, refresh_tree=True, num_points_per_cloud=num_points)
            idx_full = self._knn_full_idx  # (B,P,knn_k+1) int64, column 0 = self
            moved = torch.empty_like(points)
            _ext.check(lib.isob200_resample_step(
                _ext.ptr(points), _ext.ptr(points), _ext.ptr(normals), _ext.ptr(idx_full), 1, idx_full.shape[2],
                1, _ext.ptr(inv_sigma_spatial), B, P, P, idx_full.shape[2] - 1, _ext.ptr(moved),
                _ext.stream(dev)))
            # NB (:284-286): the reference carries the MOVED, un-projected points into the next
            # sample_iter (`points = points + move`); the projection result is only returned.
            points = moved
            projection_result = self._project_points(model, points, num_points, proj_max_iters=3,
                                                     num_points_list=num_points_list, _live=_live,
                                                     **forward_kwargs)
        return projection_result

    # ------------------------------------------------------------------------------------
    def insert(self, ref_pcl, points, num_points, current_knn_result=None):
        """Insert children around points close to high-saliency ``ref_pcl`` points
        (levelset_sampling.py:172-233)."""
        batch_size = points.shape[0]
        diag = (points.view(-1, 3).max(dim=0).values - points.view(-1, 3).min(0).values).norm().item()
        avg_spacing = math.sqrt(diag / ref_pcl.num_points_per_cloud().item())
        patch_size = 8
        knn_k = patch_size
        search_radius = min(avg_spacing * knn_k, 0.2)
        if current_knn_result is None:
            dists, idxs, _, _ = frnn.frnn_grid_points(points, points, num_points, num_points, K=knn_k + 1,
                                                      r=search_radius, grid=None, return_nn=False)
            current_knn_result = _KNN(dists=dists[..., 1:], idx=idxs[..., 1:], knn=None)
        try:
            metrics = ref_pcl.features_packed()
            num_ref = metrics.shape[0]
            threshold = min(metrics.median() * 2, metrics.max() * 0.5)
            assert (len(ref_pcl) == 1), "Support only 1 point cloud"
            ref_pts = ref_pcl.points_packed()[(metrics > threshold).squeeze(-1)].view(1, -1, 3)
            if ref_pts.shape[1] == 0 or ref_pts.shape[1] > min(50, int(num_ref / 20)):
                ref_pts = ref_pcl.points_packed()[metrics.sort(dim=0).indices[
                    -max(min(50, int(num_ref / 20)), 1):, 0]].view(1, -1, 3).expand(batch_size, -1, -1)
            dists_to_ref, idxs_to_ref, _, _ = frnn.frnn_grid_points(
                points, ref_pts.expand(batch_size, -1, -1).contiguous(), lengths1=num_points, lengths2=None,
                K=1, return_nn=False, grid=None, r=search_radius * 4)
            dists_to_ref = dists_to_ref.view(batch_size, -1)
            dist_threshold = avg_spacing ** 2
            father_pts_mask = (dists_to_ref < 4 * dist_threshold) & (dists_to_ref > 0)
            father_pts = points[father_pts_mask]
            mother_pts = frnn.frnn_gather(points, current_knn_result.idx[..., -patch_size:].contiguous())
            mother_pts = mother_pts[father_pts_mask]
            child_pts = 2 * father_pts.unsqueeze(-2) / 3 + mother_pts / 3
            child_per_batch = father_pts_mask.sum(-1) * mother_pts.shape[-2]
            child_pts = child_pts.view(-1, 3)
            first_idx = F.pad(child_per_batch, (1, 0), 'constant', 0).cumsum(0)
            child_pts = packed_to_padded(child_pts, first_idx[:-1], int(child_per_batch.max().item()))
        except Exception as e:  # same recovery as the reference (:224-228)
            import logging
            logging.getLogger("isopoints_b200").error("Error occurred during insertion {}".format(e))
            child_pts = points.new_zeros((batch_size, 0, 3))
            child_per_batch = num_points.new_zeros((batch_size,))
        points = torch.cat((points, child_pts), dim=1)
        num_points = num_points + child_per_batch
        return points, num_points, child_pts, child_per_batch

    def upsample(self, points, n_points, model, num_points=None, **forward_kwargs):
        """levelset_sampling.py:235-237."""
        from .point_processing import upsample
        points, num_points = upsample(points, n_points, num_points=num_points, neighborhood_size=31)
        return points, num_points

    def _run_ahead_ok(self, model, mask, points, forward_kwargs) -> bool:
        """Whether filter + resample can be enqueued without the host knowing the survivor count: one cloud, the
        fused SIREN decoder, and none of the steps overridden (the point-sharded subclass decides collectively)."""
        cls, base = type(self), UniformProjection
        return (RUN_AHEAD and mask.shape[0] == 1 and mask.shape[1] > 0 and points.is_cuda
                and points.dtype == torch.float32
                and all(getattr(cls, f) is getattr(base, f)
                        for f in ("_nothing_converged", "_create_tree", "resample", "_project_points"))
                and getattr(model, "isob200_analytic_sdf", None) is None
                and siren.match(model, forward_kwargs) is not None)

    def _filter_resample_run_ahead(self, model, result: ProjectionResult, sample_iters, **forward_kwargs):
        """``_filter_projection_result`` + ``resample`` (:401-409) with the survivor count left on the device: the
        compaction keeps its M-row buffers, every later kernel takes the count from device memory (FRNN lengths,
        the bounding box, the Newton kernels' row count), and the one read-back happens after the whole step is
        enqueued -- the host runs ahead of the GPU instead of waiting for the projection to drain.  Returns None
        when nothing converged (:396-399)."""
        points, normals, mask = result
        lib = _ext.lib()
        dev = points.device
        M = int(mask.shape[1])
        pts, nrm = points.reshape(-1, 3).contiguous(), normals.reshape(-1, 3).contiguous()
        valid = mask.reshape(-1).contiguous().view(torch.uint8)
        out_p, out_n = torch.zeros_like(pts), torch.zeros_like(nrm)   # rows past the count: finite
        count = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws = _ext.workspace(lib.isob200_project_step_ws_bytes(M), dev)
        _ext.check(lib.isob200_compact_valid(_ext.ptr(pts), _ext.ptr(nrm), _ext.ptr(valid), M, _ext.ptr(out_p),
                                             _ext.ptr(out_n), _ext.ptr(count), _ext.ptr(ws), ws.numel(),
                                             _ext.stream(dev)))
        tree = {k: self.__dict__.get(k) for k in ("_knn_idx", "_knn_dists", "_knn_full_idx", "_knn_src",
                                                  "_knn_nn_cache")}
        outer, checks = frnn.DEFERRED_GRID_CHECKS, []
        frnn.DEFERRED_GRID_CHECKS = checks   # no grid-size read-back either: cell tables sized ahead, checked below
        try:
            res = self.resample(model, out_p[None], out_n[None], count.long(), sample_iters=sample_iters,
                                _live=count, **forward_kwargs)
        finally:
            frnn.DEFERRED_GRID_CHECKS = outer
        n = int(count.item())
        if n == 0 or sample_iters == 0 or 3 * n < 2 * (self.knn_k + 1):
            self.__dict__.update(tree)   # the read-back path builds no tree in these cases
            if n == 0:
                return None
            # resample's own early returns (:242-245)
            return ProjectionResult(out_p[:n][None], out_n[:n][None], mask.new_ones((1, n)))
        if any(int(g.item()) > cap for g, cap in checks):
            # a grid needed more cells than the table sized ahead of time (a strongly anisotropic box): its search
            # found nothing; redo on the compacted survivors with the sizes read back
            self.__dict__.update(tree)
            return self.resample(model, out_p[:n][None], out_n[:n][None], count.long(), sample_iters=sample_iters,
                                 num_points_list=[n], **forward_kwargs)
        if self._knn_idx is not None and self._knn_idx.shape[1] == M:
            self._knn_full_idx = self._knn_full_idx[:, :n]
            self._knn_idx, self._knn_dists = self._knn_idx[:, :n], self._knn_dists[:, :n]
            self._knn_src, self._knn_nn_cache = self._knn_src[:, :n], None
        return ProjectionResult(res.points[:, :n], res.normals[:, :n], res.mask[:, :n])

    def _nothing_converged(self, counts) -> bool:
        """The early exit of :396-399 (``not valid_projection.any()``) from the survivor counts already on the
        host.  A hook: the point-sharded subclass has to take this decision collectively."""
        return sum(counts) == 0

    # ------------------------------------------------------------------------------------
    def project_points(self, point_clouds, model, normals_init: Optional[torch.Tensor] = None,
                       skip_resampling: bool = False, skip_upsampling: bool = False,
                       ref_pcl=None, proj_max_iters: Optional[int] = None,
                       sample_iters: Optional[int] = None, **forward_kwargs):
        """project -> (filter, resample) -> (insert | upsample, re-project); levelset_sampling.py:353-439.

        Returns {'levelset_points', 'levelset_normals', 'mask'}; 'levelset_normals' is absent when
        nothing converged (:396-399).
        """
        points_init, num_points = convert_pointclouds_to_tensor(point_clouds)
        num_points_init = num_points
        proj_max_iters = proj_max_iters or self.proj_max_iters
        sample_iters = sample_iters or self.sample_iters
        if normals_init is None and is_pointclouds(point_clouds):
            normals_init = point_clouds.normals_padded()

        num_list = None
        if torch.is_tensor(point_clouds):   # dense tensor input: every cloud has all P rows
            num_list = [int(points_init.shape[1])] * int(points_init.shape[0])
        with autograd.no_grad():
            points_projected, normals_projected, valid_projection = self._project_points(
                model, points_init, num_points, proj_max_iters=proj_max_iters, num_points_list=num_list,
                **forward_kwargs)
            if skip_resampling:
                if not valid_projection.any():
                    return {'levelset_points': points_projected, 'mask': valid_projection}
            else:
                unfiltered = (points_projected, valid_projection)
                if self._run_ahead_ok(model, valid_projection, points_projected, forward_kwargs):
                    res = self._filter_resample_run_ahead(
                        model, ProjectionResult(points_projected, normals_projected, valid_projection),
                        sample_iters, **forward_kwargs)
                    if res is None:   # (:396-399)
                        return {'levelset_points': unfiltered[0], 'mask': unfiltered[1]}
                    points_projected, normals_projected, valid_projection = res
                else:
                    (points_projected, normals_projected, valid_projection), counts, num_points = \
                        _filter_projection_result_counted(
                            ProjectionResult(points_projected, normals_projected, valid_projection))
                    if self._nothing_converged(counts):   # (:396-399)
                        return {'levelset_points': unfiltered[0], 'mask': unfiltered[1]}
                    points_projected, normals_projected, valid_projection = self.resample(
                        model, points_projected, normals_projected, num_points, sample_iters=sample_iters,
                        num_points_list=counts, **forward_kwargs)
                num_points = valid_projection.sum(dim=-1)

            if not skip_upsampling and ref_pcl is not None:
                points_projected, normals_projected, valid_projection = _filter_projection_result(
                    ProjectionResult(points_projected, normals_projected, valid_projection))
                num_points = valid_projection.sum(dim=-1)
                _, _, new_points, num_new_points = self.insert(ref_pcl, points_projected, num_points)
                new_points_projected, new_normals_projected, new_valid_projection = self._project_points(
                    model, new_points, num_new_points, proj_max_iters=10, **forward_kwargs)
                points_projected = torch.cat([points_projected, new_points_projected], dim=1)
                normals_projected = torch.cat([normals_projected, new_normals_projected], dim=1)
                valid_projection = torch.cat([valid_projection, new_valid_projection], dim=1)
            elif not skip_upsampling:
                points_projected, normals_projected, valid_projection = _filter_projection_result(
                    ProjectionResult(points_projected, normals_projected, valid_projection))
                num_points = valid_projection.sum(dim=-1)
                points_projected, num_points = self.upsample(
                    points_projected, num_points_init, model, num_points, **forward_kwargs)
                points_projected, normals_projected, valid_projection = self._project_points(
                    model, points_projected, num_points, proj_max_iters=10, **forward_kwargs)

            return {'levelset_points': points_projected,
                    'levelset_normals': normals_projected,
                    'mask': valid_projection}


class SphereTracing(LevelSetProjection):
    """Ray marching onto the level set (levelset_sampling.py:663-808): every ray advances by
    ``alpha * sdf`` along its direction (step clamped to 0.1) until ``|sdf| <= 0.1 * proj_tolerance``, it
    would leave the sphere of radius ``radius + padding``, or ``proj_max_iters`` steps are used.

    Same machinery as ``UniformProjection._project_points``: one SDF evaluation on the compacted active
    rays + one fused update / compaction kernel (``isob200_trace_step``) per iteration; with the reference's
    Siren decoder the loop runs off device-side ray counts without a read-back."""

    def __init__(self, proj_max_iters=10, proj_tolerance=5e-5, max_points_per_pass=120000,
                 alpha=1.0, radius=1.0, padding=0.1, **kwargs):
        super().__init__(proj_max_iters=proj_max_iters, proj_tolerance=proj_tolerance,
                         max_points_per_pass=max_points_per_pass)
        self.alpha = alpha
        self.radius = radius
        self.padding = padding

    def _eval(self, points, model, latent, forward_kwargs):
        """SDF + input gradient of the active rays through autograd (:742-756), chunked."""
        sdfs, grads = [], []
        lat = [None] * (1 + points.shape[0] // self.max_points_per_pass) if latent is None else \
            torch.split(latent, self.max_points_per_pass, dim=0)
        for sub, c in zip(torch.split(points, self.max_points_per_pass, dim=0), lat):
            with autograd.enable_grad():
                net_input = sub.detach().requires_grad_(True)
                network_eval = model.forward(net_input, c=c, **forward_kwargs).sdf
                g = autograd.grad([network_eval], [net_input], torch.ones_like(network_eval))[0]
            sdfs.append(network_eval.detach().reshape(-1))
            grads.append(g.detach())
        return torch.cat(sdfs).float().contiguous(), torch.cat(grads).float().contiguous()

    def project_points(self, ray0: torch.Tensor, ray_direction: torch.Tensor, model,
                       latent: Optional[torch.Tensor] = None, **forward_kwargs):
        """Returns {'levelset_points' (shape of ray0), 'network_eval_on_levelset_points' (shape[:-1]),
        'levelset_points_Dx' (the reference returns the points again under this key, :806), 'mask'}."""
        shp = ray0.shape
        _ext.require_cuda(ray0, ray_direction)
        ray0, ray_direction = torch.broadcast_tensors(ray0, ray_direction)
        if ray0.dtype != torch.float32:
            raise RuntimeError("expected scalar type Float")
        dev = ray0.device
        points = ray0.reshape(-1, 3).clone()
        dirs = ray_direction.reshape(-1, 3).float().contiguous()
        M = points.shape[0]
        if latent is not None and latent.nelement() > 0:
            latent = latent.squeeze()
            assert latent.ndim == 2
            if len(shp) > 2:   # (N, C) per cloud -> one row per ray (:44-52)
                latent = torch.repeat_interleave(latent, M // shp[0], dim=0) if latent.shape[0] > 1 \
                    else latent.expand(M, -1)
        else:
            latent = None
        lib = _ext.lib()
        network_eval = torch.zeros((M,), dtype=torch.float32, device=dev)
        grad = torch.zeros((M, 3), dtype=torch.float32, device=dev)
        iters = int(self.proj_max_iters)
        args = (float(0.1 * self.proj_tolerance), float(self.alpha), 0.1, float(self.padding + self.radius))
        with autograd.no_grad():
            model.eval()
            fused = siren.match(model, forward_kwargs) if (M > 0 and latent is None) else None
            if M > 0:
                act = (torch.empty((M,), dtype=torch.int32, device=dev), torch.empty((M,), dtype=torch.int32, device=dev))
                nxt = (torch.empty((M, 3), dtype=torch.float32, device=dev),
                       torch.empty((M, 3), dtype=torch.float32, device=dev))
                ws = _ext.workspace(lib.isob200_project_step_ws_bytes(M), dev)
                st = _ext.stream(dev)
            if fused is not None:
                # The reference's Siren decoder: one kernel per iteration -- the FORWARD half of the fused
                # tcgen05 SDF kernel (the march never uses the gradient the reference computes at :745-757),
                # the update of :764-777 and the compaction of the still-active rays -- off device-side ray
                # counts: the whole loop is enqueued without a read-back.
                cnt = torch.zeros((iters + 2,), dtype=torch.int32, device=dev)   # active rays entering iteration it
                blob, scratch, n_hidden = siren.packed(model, fused)
                for it in range(iters + 1):
                    last = (it == iters)
                    cur = points if it == 0 else nxt[it & 1]
                    _ext.check(lib.isob200_siren_trace_step(
                        _ext.ptr(cur), M, None if it == 0 else _ext.ptr(cnt[it:]), _ext.ptr(blob), n_hidden,
                        _ext.ptr(scratch), scratch.numel(), _ext.ptr(points), _ext.ptr(dirs), _ext.ptr(network_eval),
                        None if it == 0 else _ext.ptr(act[it & 1]), *args, 0 if last else 1,
                        _ext.ptr(act[(it + 1) & 1]), None if last else _ext.ptr(nxt[(it + 1) & 1]),
                        _ext.ptr(cnt[it + 1:]), st))
                siren.STATS["calls"] += iters + 1
                siren.STATS["rows"] += M
                grad = None
            elif M > 0:
                count = torch.zeros((1,), dtype=torch.int32, device=dev)
                A, it = M, 0
                while True:
                    cur = points if it == 0 else nxt[it & 1][:A]
                    lat = latent if (latent is None or it == 0) else latent[act[it & 1][:A].long()]
                    curr_sdf, curr_grad = self._eval(cur, model, lat, forward_kwargs)
                    last = (it == iters)
                    _ext.check(lib.isob200_trace_step(
                        _ext.ptr(points), _ext.ptr(dirs), _ext.ptr(network_eval), _ext.ptr(grad),
                        None if it == 0 else _ext.ptr(act[it & 1]), A, None, _ext.ptr(curr_sdf), _ext.ptr(curr_grad),
                        *args, 0 if last else 1, _ext.ptr(act[(it + 1) & 1]),
                        None if last else _ext.ptr(nxt[(it + 1) & 1]), _ext.ptr(count), _ext.ptr(ws), ws.numel(), st))
                    if last:
                        break
                    A = int(count.item())   # the opaque SDF callback needs the batch shape
                    if A == 0:
                        break
                    it += 1
        valid_projection = network_eval.abs() <= self.proj_tolerance
        levelset_points = points.view(shp)
        # d sdf / d x at the last evaluation of each ray (:745-757; not part of the reference's return value):
        # available on the autograd path only, the fused path evaluates the forward half of the network
        self.last_gradient = None if grad is None else grad.view(shp)
        return {'levelset_points': levelset_points,
                'network_eval_on_levelset_points': network_eval.view(shp[:-1]),
                'levelset_points_Dx': levelset_points,
                'mask': valid_projection.view(shp[:-1])}


def _mask_padded_to_list(values, mask):
    """DSS/utils/__init__.py:119-134: per-cloud tensors of the mask == True rows."""
    return [values[b][mask[b]] for b in range(values.shape[0])]


def sample_uniform_iso_points(model, n_points: int, init_points: Optional[torch.Tensor] = None,
                              bounding_sphere_radius: float = 1.0, pointclouds_cls=None):
    """Uniformly distributed iso-points of ``model`` (levelset_sampling.py:1405-1445): project 4n random
    points, keep those inside the bounding sphere, WLOP-consolidate to ~n/2.., upsample + re-project
    twice.  Returns a ``Pointclouds`` with one cloud of ~``n_points`` points."""
    from .point_processing import upsample, wlop
    from .structures import Pointclouds as _Local
    Pointclouds = pointclouds_cls or _Local
    projector = UniformProjection(max_points_per_pass=16000, proj_max_iters=10, proj_tolerance=5e-5, knn_k=8)
    if init_points is None:
        init_points = (torch.rand((1, n_points * 4, 3)) - 0.5) * 2 * bounding_sphere_radius
        init_points = init_points.cuda()
    proj_results = projector.project_points(init_points, model, skip_resampling=True, skip_upsampling=True)
    boundary_mask = proj_results['levelset_points'].norm(dim=-1) < bounding_sphere_radius
    proj_pcl = Pointclouds(_mask_padded_to_list(proj_results['levelset_points'],
                                                proj_results['mask'] & boundary_mask.view_as(proj_results['mask'])))
    wlop_result = wlop(proj_pcl, min(0.5, n_points / proj_pcl.num_points_per_cloud().item()))
    proj_results = projector.project_points(wlop_result, model, skip_resampling=True, skip_upsampling=False)
    proj_pcl = Pointclouds(_mask_padded_to_list(proj_results['levelset_points'], proj_results['mask']))
    upsampled_pcl = upsample(proj_pcl, n_points)
    proj_results = projector.project_points(upsampled_pcl, model, skip_resampling=True, skip_upsampling=False)
    return Pointclouds(_mask_padded_to_list(proj_results['levelset_points'], proj_results['mask']))


class EdgeAwareProjection(UniformProjection):
    """Edge-aware resampling variant (levelset_sampling.py:442-660): exact K-NN neighbourhoods
    (``knn_points``, K = knn_k + 1 = 32 by default), bilateral normal denoising, a LOP-style move and
    insertion of mid-points scored by spacing x normal deviation.  The K x K mid-point / neighbour
    comparison runs in one kernel (csrc/pointops.cu, ``upsample_sparsity_kernel`` with normals)
    instead of the reference's (N,P,K,K,3) tensor; the remaining per-neighbour math is PyTorch like
    in the reference.  Effectively one cloud per call (``inv_sigma_spatial`` broadcasting, :552)."""

    def __init__(self, proj_max_iters=10, proj_tolerance=5e-5, max_points_per_pass=120000, knn_k=31,
                 repulsion_mu=0.5, sample_iters=5, total_iters=1, sharpness_angle=15, edge_sensitivity=1,
                 resampling_clip=0.02, upsample_ratio=1.5, **kwargs):
        super().__init__(sample_iters=sample_iters, total_iters=total_iters, resampling_clip=resampling_clip,
                         knn_k=knn_k, proj_max_iters=proj_max_iters, proj_tolerance=proj_tolerance,
                         max_points_per_pass=max_points_per_pass)
        self.sharpness_sigma = 1 - math.cos(sharpness_angle / 180 * math.pi)
        self.repulsion_mu = repulsion_mu
        self.edge_sensitivity = edge_sensitivity
        self.upsample_ratio = upsample_ratio

    def _create_tree(self, points_padded: torch.Tensor, refresh_tree=True, num_points_per_cloud=None):
        """:472-498 -- exact K-NN instead of the radius search of the base class."""
        from .point_processing import knn_points
        if not refresh_tree and getattr(self, '_knn_idx', None) is not None:
            return self._knn_idx
        assert (points_padded.ndim == 3)
        if num_points_per_cloud is None:
            num_points_per_cloud = torch.tensor([points_padded.shape[1]] * points_padded.shape[0],
                                                device=points_padded.device, dtype=torch.long)
        knn_result = knn_points(points_padded, points_padded, num_points_per_cloud, num_points_per_cloud,
                                K=self.knn_k + 1, return_nn=True, return_sorted=True)
        self._knn_full_idx = knn_result.idx
        self._knn_idx = knn_result.idx[..., 1:]
        self._knn_dists = knn_result.dists[..., 1:]
        self._knn_src = points_padded
        self._knn_nn_cache = knn_result.knn[..., 1:, :]
        self.knn_gather = frnn.frnn_gather
        return self._knn_idx

    def denoise_normals(self, points, normals, num_points, **kwargs):
        """Bilateral smoothing of the normals over the cached neighbourhood (:500-525):
        w = exp(-d^2 n/2) [d^2 <= 32/n] * exp(-((1 - <n, n_j>) / sigma_n)^2)."""
        normals = F.normalize(normals, dim=-1)
        knn_normals = self.knn_gather(normals, self._knn_idx, num_points)
        self.sharpness_sigma = kwargs.get('sharpness_sigma', self.sharpness_sigma)
        weights_n = ((1 - torch.sum(knn_normals * normals[:, :, None, :], dim=-1)) / self.sharpness_sigma) ** 2
        weights_n = torch.exp(-weights_n)
        inv_sigma_spatial = num_points / 2.0
        spatial_dist = 16 / inv_sigma_spatial
        deltap = self._knn_nn - points[:, :, None, :]
        deltap = torch.sum(deltap * deltap, dim=-1)
        weights_p = torch.exp(-deltap * inv_sigma_spatial)
        weights_p = torch.where(deltap > spatial_dist, torch.zeros_like(weights_p), weights_p)
        weights = weights_p * weights_n
        sw = torch.sum(weights, dim=-1, keepdim=True)
        sgn = torch.where(sw < 0, -torch.ones_like(sw), torch.ones_like(sw))
        normals_denoised = torch.sum(knn_normals * weights[:, :, :, None], dim=-2) / (sgn * sw.abs().clamp_min(1e-17))
        normals_denoised = F.normalize(normals_denoised, dim=-1)
        return normals_denoised, weights_p, weights_n

    def upsample(self, points, n_points, model, num_points=None, **forward_kwargs):
        """LOP move along denoised normals + edge-aware mid-point insertion up to
        ``n_points * upsample_ratio`` points (:527-660)."""
        from .point_processing import padded_to_list
        from .structures import list_to_padded

        def _eps_denom(x, eps=1e-17):
            sgn = torch.where(x < 0, -torch.ones_like(x), torch.ones_like(x))
            return sgn * x.abs().clamp_min(eps)

        upsample_ratio = forward_kwargs.pop('upsample_ratio', self.upsample_ratio)
        n_points = n_points * upsample_ratio
        n_points = n_points.ceil().long() if isinstance(n_points, torch.Tensor) else int(math.ceil(n_points))
        batch_size = points.shape[0]
        dev = points.device
        if num_points is None:
            num_points = torch.full((batch_size,), points.shape[1], dtype=torch.long, device=dev)
        self._create_tree(points, refresh_tree=True, num_points_per_cloud=num_points)
        inv_sigma_spatial = num_points / 2.0
        spatial_dist = 16 / inv_sigma_spatial
        _, normals = self._compute_sdf_and_grad(points, model, **forward_kwargs)
        normals = F.normalize(normals, dim=-1, eps=1e-15)
        normals, _, _ = self.denoise_normals(points, normals, num_points)

        knn_d, knn_nn = self._knn_dists, self._knn_nn
        move_clip = knn_d[..., 0].mean().sqrt()
        diff = points[:, :, None, :] - knn_nn
        weight_lop = torch.exp(-torch.sum(normals[:, :, None, :] * diff, dim=-1) ** 2 * inv_sigma_spatial)
        weight_lop = torch.where(knn_d > spatial_dist, torch.zeros_like(weight_lop), weight_lop)
        spatial_w = torch.exp(-knn_d * inv_sigma_spatial)
        spatial_w = torch.where(knn_d > spatial_dist, torch.zeros_like(spatial_w), spatial_w)
        density_w = torch.sum(spatial_w, dim=-1) + 1.0
        move_data = torch.sum(weight_lop[..., None] * diff, dim=-2) / _eps_denom(torch.sum(weight_lop, dim=-1, keepdim=True))
        move_repul = self.repulsion_mu * density_w[..., None] * torch.sum(spatial_w[..., None] * (-diff), dim=-2) / \
            _eps_denom(torch.sum(spatial_w, dim=-1, keepdim=True))
        # F.normalize without dim normalises along dim=1 in the reference (:583-586) -- reproduced
        move_repul = F.normalize(move_repul) * move_repul.norm(dim=-1, keepdim=True).clamp_max(move_clip)
        move_data = F.normalize(move_data) * move_data.norm(dim=-1, keepdim=True).clamp_max(move_clip)
        points = points - (move_data + move_repul)

        n_remaining = n_points - num_points
        lib = _ext.lib()
        K = self.knn_k
        idx_full = self._knn_full_idx
        max_P = points.shape[1] // 10
        while not bool((n_remaining == 0).all()):
            points = points.contiguous()
            B, P, _ = points.shape
            sparsity = torch.empty((B, P), dtype=torch.float32, device=dev)
            child = torch.empty((B, P, 3), dtype=torch.float32, device=dev)
            _ext.check(lib.isob200_upsample_sparsity(
                _ext.ptr(points), _ext.ptr(normals.contiguous()), float(self.edge_sensitivity), _ext.ptr(idx_full),
                K + 1, 1, _ext.ptr(num_points), B, P, K, _ext.ptr(sparsity), _ext.ptr(child), _ext.stream(dev)))
            order = sparsity.sort(dim=1).indices[:, P - max_P:] if max_P > 0 else sparsity.new_zeros((B, 0)).long()
            n_new = torch.clamp(n_remaining, max=max_P)
            new_pts = torch.gather(child, 1, order.unsqueeze(-1).expand(-1, -1, 3))
            nn_list, np_list = n_new.tolist(), num_points.tolist()
            total = [torch.cat([new_pts[b][max_P - nn_list[b]:], points[b, : np_list[b]]], dim=0) for b in range(B)]
            points = list_to_padded(total)
            n_remaining = n_remaining - n_new
            num_points = n_new + num_points
            if max_P == 0:
                break
            self._create_tree(points, num_points_per_cloud=num_points, refresh_tree=True)
            idx_full = self._knn_full_idx
            _, normals = self._compute_sdf_and_grad(points, model, **forward_kwargs)
            normals = F.normalize(normals, dim=-1)
        return points, num_points
