"""Minimal ``Pointclouds``-compatible container + the packed/padded helpers the hot path uses.

pytorch3d is not installable here, and the operators in this package never ``isinstance``-check
against this class: they duck-type on the methods below, so real
``pytorch3d.structures.Pointclouds`` / ``DSS.core.cloud.PointClouds3D`` objects work too and
``pcl.__class__(...)`` is used to build results like the reference does
(DSS/utils/point_processing.py:294).  Semantics follow pytorch3d.structures.Pointclouds
[third party, restated from its documented behaviour]: padded tensors are zero padded, packed
tensors are the concatenation of the per-cloud lists.
"""
from typing import List, Optional, Sequence, Union

import torch
import torch.nn.functional as F


def is_pointclouds(pcl) -> bool:
    return hasattr(pcl, "points_padded") and hasattr(pcl, "num_points_per_cloud")


def convert_pointclouds_to_tensor(pcl):
    """pytorch3d.ops.utils.convert_pointclouds_to_tensor: -> (padded (N,P,D), num_points (N,) i64)."""
    if is_pointclouds(pcl):
        return pcl.points_padded(), pcl.num_points_per_cloud()
    if torch.is_tensor(pcl):
        X = pcl
        num_points = X.shape[1] * torch.ones(X.shape[0], device=X.device, dtype=torch.int64)
        return X, num_points
    raise ValueError("The inputs X, Y should be either Pointclouds objects or tensors.")


def num_points_2_cloud_to_packed_first_idx(num_points):
    """DSS/utils/__init__.py:26-29."""
    first = F.pad(num_points, (1, 0), "constant", 0).cumsum(0)
    return first[:-1]


def padded_to_packed_idx(num_points, P):
    """Flat indices (into N*P rows) of the live rows of a padded tensor, cloud-major."""
    N = num_points.shape[0]
    ar = torch.arange(P, device=num_points.device)
    mask = ar[None, :] < num_points[:, None]
    return mask.view(-1).nonzero(as_tuple=False).squeeze(1), mask


def list_to_padded(xs: Sequence[torch.Tensor], pad_size=None):
    N = len(xs)
    P = max([x.shape[0] for x in xs], default=0) if pad_size is None else pad_size
    tail = xs[0].shape[1:] if N else (3,)
    out = xs[0].new_zeros((N, P) + tuple(tail)) if N else torch.zeros((0, P) + tuple(tail))
    for i, x in enumerate(xs):
        out[i, : x.shape[0]] = x
    return out


def packed_to_padded(packed, first_idx, max_size):
    """pytorch3d.ops.packed_to_padded: packed (F,D) or (F,), first_idx (N,) -> (N,max_size,D)."""
    squeeze = packed.ndim == 1
    if squeeze:
        packed = packed[:, None]
    N = first_idx.shape[0]
    F_ = packed.shape[0]
    ends = torch.cat([first_idx[1:], first_idx.new_tensor([F_])])
    lens = ends - first_idx
    ar = torch.arange(max_size, device=packed.device)
    mask = ar[None, :] < lens[:, None]
    src = (first_idx[:, None] + ar[None, :])[mask]
    out = packed.new_zeros((N, max_size, packed.shape[1]))
    out[mask] = packed[src]
    return out.squeeze(-1) if squeeze else out


def padded_to_packed(padded, first_idx, num_packed):
    """pytorch3d.ops.padded_to_packed: (N,M,D) -> (num_packed,D)."""
    squeeze = padded.ndim == 2
    if squeeze:
        padded = padded[:, :, None]
    N, M, D = padded.shape
    ends = torch.cat([first_idx[1:], first_idx.new_tensor([num_packed])])
    lens = ends - first_idx
    ar = torch.arange(M, device=padded.device)
    mask = ar[None, :] < lens[:, None]
    out = padded[mask]
    return out.squeeze(-1) if squeeze else out


def reduce_mask_padded(values, mask):
    """DSS/utils/__init__.py:149-169: keep mask==True entries, re-pad to the largest survivor count."""
    batch_size = values.shape[0]
    value_packed = values[mask]
    num_true = mask.view(batch_size, -1).sum(dim=1)
    first_idx = num_points_2_cloud_to_packed_first_idx(num_true)
    dtype = value_packed.dtype
    pmax = int(num_true.max().item()) if batch_size else 0
    return packed_to_padded(value_packed.float(), first_idx, pmax).to(dtype=dtype)


class Pointclouds:
    """List/padded/packed point cloud batch (the subset of pytorch3d's API used on the hot path)."""

    def __init__(self, points, normals=None, features=None):
        if torch.is_tensor(points):
            if points.ndim != 3:
                raise ValueError("Points tensor has incorrect dimensions.")
            self._points_list = [points[i] for i in range(points.shape[0])]
            self._points_padded = points
        else:
            self._points_list = list(points)
            self._points_padded = None
        self._N = len(self._points_list)
        self.device = self._points_list[0].device if self._N else torch.device("cpu")
        self._num = torch.tensor([p.shape[0] for p in self._points_list], dtype=torch.int64,
                                 device=self.device)
        self._normals_list = self._aux(normals)
        self._features_list = self._aux(features)

    def _aux(self, a):
        if a is None:
            return None
        if torch.is_tensor(a):
            if a.ndim != 3:
                raise ValueError("Auxiliary tensor has incorrect dimensions.")
            return [a[i, : self._points_list[i].shape[0]] for i in range(self._N)]
        a = list(a)
        if len(a) != self._N:
            raise ValueError("Points and auxiliary input must be the same length.")
        return a

    # ---- size / access ----
    def __len__(self):
        return self._N

    def __getitem__(self, index):
        if isinstance(index, int):
            index = [index]
        elif isinstance(index, slice):
            index = list(range(self._N))[index]
        elif torch.is_tensor(index):
            if index.dtype == torch.bool:
                index = index.nonzero().squeeze(1)
            index = index.tolist()
        pts = [self._points_list[i] for i in index]
        nrm = [self._normals_list[i] for i in index] if self._normals_list is not None else None
        ft = [self._features_list[i] for i in index] if self._features_list is not None else None
        return self.__class__(pts, normals=nrm, features=ft)

    def isempty(self) -> bool:
        return self._N == 0 or int(self._num.sum().item()) == 0

    def num_points_per_cloud(self):
        return self._num

    def points_list(self):
        return self._points_list

    def normals_list(self):
        return self._normals_list

    def features_list(self):
        return self._features_list

    def points_padded(self):
        if self._points_padded is None:
            self._points_padded = list_to_padded(self._points_list)
        return self._points_padded

    def normals_padded(self):
        return None if self._normals_list is None else list_to_padded(self._normals_list, self.points_padded().shape[1])

    def features_padded(self):
        return None if self._features_list is None else list_to_padded(self._features_list, self.points_padded().shape[1])

    def points_packed(self):
        return torch.cat(self._points_list, dim=0) if self._N else torch.zeros((0, 3))

    def normals_packed(self):
        return None if self._normals_list is None else torch.cat(self._normals_list, dim=0)

    def features_packed(self):
        return None if self._features_list is None else torch.cat(self._features_list, dim=0)

    def cloud_to_packed_first_idx(self):
        return num_points_2_cloud_to_packed_first_idx(self._num)

    def packed_to_cloud_idx(self):
        return torch.repeat_interleave(torch.arange(self._N, device=self.device), self._num, dim=0)

    def get_bounding_boxes(self):
        """(N, 3, 2): min and max per axis."""
        out = []
        for p in self._points_list:
            out.append(torch.stack([p.min(dim=0)[0], p.max(dim=0)[0]], dim=1))
        return torch.stack(out, dim=0)

    # ---- functional updates ----
    def clone(self):
        return self.__class__([p.clone() for p in self._points_list],
                              normals=None if self._normals_list is None else [x.clone() for x in self._normals_list],
                              features=None if self._features_list is None else [x.clone() for x in self._features_list])

    def extend(self, N: int):
        if not isinstance(N, int):
            raise ValueError("N must be an integer.")
        if N <= 0:
            raise ValueError("N must be > 0.")
        rep = lambda l: None if l is None else [x.clone() for x in l for _ in range(N)]
        return self.__class__(rep(self._points_list), normals=rep(self._normals_list),
                              features=rep(self._features_list))

    def update_padded(self, new_points_padded, new_normals_padded=None, new_features_padded=None):
        n = self._num.tolist()
        pts = [new_points_padded[i, : n[i]] for i in range(self._N)]
        nrm = self._normals_list if new_normals_padded is None else [new_normals_padded[i, : n[i]] for i in range(self._N)]
        ft = self._features_list if new_features_padded is None else [new_features_padded[i, : n[i]] for i in range(self._N)]
        return self.__class__(pts, normals=nrm, features=ft)

    def offset_(self, offsets_packed):
        off = torch.split(offsets_packed, self._num.tolist(), dim=0)
        self._points_list = [p + o for p, o in zip(self._points_list, off)]
        self._points_padded = None
        return self

    def offset(self, offsets_packed):
        return self.clone().offset_(offsets_packed)

    def update_normals_(self, normals):
        self._normals_list = self._aux(normals)
        return self

    def update_features_(self, features):
        self._features_list = self._aux(features)
        return self

    def to(self, device):
        mv = lambda l: None if l is None else [x.to(device) for x in l]
        return self.__class__(mv(self._points_list), normals=mv(self._normals_list), features=mv(self._features_list))

    def cuda(self):
        return self.to("cuda")

    def cpu(self):
        return self.to("cpu")
