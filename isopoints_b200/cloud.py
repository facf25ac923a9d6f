"""Per-point boolean filters over a batch of point clouds.

Mirrors ``PointCloudsFilters`` (DSS/core/cloud.py:289-367): three padded (N, P_max) masks -- ``inmask``,
``activation``, ``visibility`` -- each defaulting to a single broadcastable ``True``.  ``SurfaceSplatting``
writes ``visibility`` (rasterizer.py:231-235, 647-650); ``get_visible_points`` (DSS/utils/__init__.py:699-711)
reads it back through ``filter_with``.
"""
from typing import Sequence

import torch

from .structures import Pointclouds, convert_pointclouds_to_tensor, is_pointclouds

__all__ = ["PointCloudsFilters"]

_NAMES = ("inmask", "activation", "visibility")


class PointCloudsFilters:
    """Filters are padded 2-D boolean masks (N, P_max); (1, 1) broadcasts over clouds and points."""

    def __init__(self, device="cpu", inmask=None, activation=None, visibility=None):
        self.device = torch.device(device)
        for name, v in zip(_NAMES, (inmask, activation, visibility)):
            if v is None:
                v = torch.tensor([[True]], dtype=torch.bool)
            setattr(self, name, torch.as_tensor(v, dtype=torch.bool).to(self.device))

    def set_filter(self, **kwargs):
        """cloud.py:304-314: replace the named filters (2-D tensors); the device follows the new tensors."""
        for k, v in kwargs.items():
            if k not in _NAMES:
                raise AttributeError("unknown filter %r" % (k,))
            if v.ndim != 2:
                raise ValueError("filter should be a 2-dim tensor (padded values)")
            self.device = v.device
            setattr(self, k, v.to(torch.bool))
        for k in _NAMES:
            setattr(self, k, getattr(self, k).to(self.device))

    def filter(self, point_clouds):
        """Filter with all the existing filters (cloud.py:316-320)."""
        return self.filter_with(point_clouds, _NAMES)

    def filter_with(self, point_clouds, filter_names: Sequence[str]):
        """cloud.py:322-367: the points (normals, features) every named filter keeps, as a new Pointclouds with
        max(N_clouds, N_filter) clouds; padded positions never pass."""
        if is_pointclouds(point_clouds) and point_clouds.isempty():
            return point_clouds
        points, num = convert_pointclouds_to_tensor(point_clouds)
        filters = [getattr(self, k).to(points.device) for k in filter_names]
        N = max([points.shape[0]] + [f.shape[0] for f in filters])
        P = max([points.shape[1]] + [f.shape[1] for f in filters])
        if P != points.shape[1]:
            raise ValueError("filter of %d points on clouds padded to %d" % (P, points.shape[1]))
        keep = torch.ones((N, P), dtype=torch.bool, device=points.device)
        for f in filters:
            keep = keep & f                                      # (1|N, 1|P) broadcasts
        num = num.expand(N) if num.shape[0] == 1 else num
        keep = keep & (torch.arange(P, device=points.device)[None, :] < num[:, None])
        counts = keep.sum(dim=1).tolist()                       # one read-back: the new cloud sizes

        def pick(padded):
            if padded is None:
                return None
            padded = padded.expand(N, -1, -1)
            return list(torch.split(padded[keep], counts))
        if not is_pointclouds(point_clouds):
            return Pointclouds(pick(points))
        return Pointclouds(pick(points), normals=pick(point_clouds.normals_padded()),
                           features=pick(point_clouds.features_padded()))
