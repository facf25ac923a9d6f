"""Load the reference's own Python hot path from /root/reference, unmodified on disk.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the authoring
container): tests/golden/make_golden.py uses it to produce the committed golden vectors.

* Missing third-party packages (pytorch3d, trimesh, frnn, ...) are auto-stubbed by a
  sys.meta_path finder: every attribute resolves to a dummy class so that module-level
  `class X(PytorchPointClouds)` / `from a.b import c` statements still execute.
* The handful of pytorch3d symbols the hot path really executes get functional pure-torch
  stand-ins (semantics restated from pytorch3d's documentation: zero padding, packed =
  concatenation).
* Two idioms PyTorch 1.6 accepted (four sites) raise on torch >= 2: they are patched in the SOURCE STRING at
  load time, each pattern asserted to match (PATCHES below).  Nothing else is changed.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_ref():
    """The reference tree: $ISO_REFERENCE, /root/reference (authoring container), or the git-ignored staged copy of
    its Python files under baseline/_ref/ (oracle/stage_ref.py; it travels to the GPU box with the snapshot)."""
    for c in (os.environ.get("ISO_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if c and os.path.isdir(os.path.join(c, "DSS")):
            return c
    return "/root/reference"


REF = _find_ref()


def available():
    return os.path.isdir(os.path.join(REF, "DSS"))

STUB_PACKAGES = (
    "pytorch3d", "trimesh", "matplotlib", "skimage", "torch_batch_svd", "torch_cluster", "frnn",
    "prefix_sum", "plotly", "imageio", "plyfile", "easydict", "git", "pymeshlab",
    "point_cloud_utils", "im2mesh", "fvcore", "cv2", "open3d", "tensorboardX", "mcubes", "kornia",
)

# (file suffix, old, new, expected count)
PATCHES = (
    ("DSS/models/levelset_sampling.py",
     "net_input.detach_().requires_grad_(True)",
     "net_input = net_input.detach().requires_grad_(True)", None),
    ("DSS/models/levelset_sampling.py",
     "not_converged[not_converged] = curr_not_converged",
     "not_converged[not_converged.clone()] = curr_not_converged", 1),
    ("DSS/core/rasterizer.py",
     "valid_depth_mask[valid_depth_mask] = frontface_mask",
     "valid_depth_mask[valid_depth_mask.clone()] = frontface_mask", 1),
    ("DSS/models/combined_modeling.py",
     "mask_insurface[b][mask_insurface[b]] = (ray_len0 < ray_len1).view(-1)",
     "mask_insurface[b][mask_insurface[b].clone()] = (ray_len0 < ray_len1).view(-1)", 1),
)


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _DummyMeta(name, (_Dummy,), {})


class _Dummy(metaclass=_DummyMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if full in sys.modules:
            return sys.modules[full]
        if name.islower():  # `from pkg import submodule`
            try:
                return importlib.import_module(full)
            except ImportError:
                pass
        cls = _DummyMeta(name, (_Dummy,), {"__module__": self.__name__})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUB_PACKAGES:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


class _PatchingLoader(importlib.machinery.SourceFileLoader):
    def get_data(self, path):
        data = super().get_data(path)
        if not path.endswith(".py"):
            return data
        src = data.decode("utf-8")
        for suffix, old, new, count in PATCHES:
            if path.endswith(suffix):
                n = src.count(old)
                assert n >= 1 and (count is None or n == count), (path, old, n)
                src = src.replace(old, new)
        return src.encode("utf-8")

    def get_code(self, fullname):  # never trust a stale .pyc for a patched file
        path = self.get_filename(fullname)
        return compile(self.get_data(path), path, "exec", dont_inherit=True)


class _RefFinder(importlib.abc.MetaPathFinder):
    """Resolves `DSS` and its submodules to files under REF through the patching loader."""

    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] != "DSS":
            return None
        rel = fullname.replace(".", "/")
        pkg = os.path.join(REF, rel, "__init__.py")
        mod = os.path.join(REF, rel + ".py")
        if os.path.exists(pkg):
            return importlib.util.spec_from_file_location(
                fullname, pkg, loader=_PatchingLoader(fullname, pkg),
                submodule_search_locations=[os.path.dirname(pkg)])
        if os.path.exists(mod):
            return importlib.util.spec_from_file_location(fullname, mod, loader=_PatchingLoader(fullname, mod))
        return None


# ---- functional stand-ins for the pytorch3d symbols the hot path executes ------------------
def _list_to_packed(x):
    n = torch.tensor([len(t) for t in x], dtype=torch.int64)
    first = torch.zeros_like(n)
    first[1:] = n.cumsum(0)[:-1]
    packed = torch.cat(x, 0)
    p2l = torch.repeat_interleave(torch.arange(len(x)), n)
    return packed, n, first, p2l


def _padded_to_list(x, split_size=None):
    xs = list(x.unbind(0))
    if split_size is None:
        return xs
    return [t[:s] if isinstance(s, int) else t[tuple(slice(0, i) for i in s)] for t, s in zip(xs, split_size)]


def _list_to_padded(x, pad_size=None, pad_value=0.0, equisized=False):
    P = max(len(t) for t in x) if pad_size is None else (pad_size if isinstance(pad_size, int) else pad_size[0])
    out = x[0].new_full((len(x), P) + tuple(x[0].shape[1:]), pad_value)
    for i, t in enumerate(x):
        out[i, :len(t)] = t
    return out


def _packed_to_padded(inputs, first_idxs, max_size):
    sq = inputs.ndim == 1
    if sq:
        inputs = inputs[:, None]
    N = first_idxs.shape[0]
    ends = torch.cat([first_idxs[1:], first_idxs.new_tensor([inputs.shape[0]])])
    out = inputs.new_zeros((N, int(max_size), inputs.shape[1]))
    for n in range(N):
        s, e = int(first_idxs[n]), int(ends[n])
        out[n, :e - s] = inputs[s:e]
    return out.squeeze(-1) if sq else out


def _padded_to_packed(inputs, first_idxs, num_inputs):
    sq = inputs.ndim == 2
    if sq:
        inputs = inputs[:, :, None]
    N = first_idxs.shape[0]
    ends = torch.cat([first_idxs[1:], first_idxs.new_tensor([num_inputs])])
    out = torch.cat([inputs[n, :int(ends[n]) - int(first_idxs[n])] for n in range(N)], 0)
    return out.squeeze(-1) if sq else out


def _is_pointclouds(p):
    return hasattr(p, "points_padded") and hasattr(p, "num_points_per_cloud")


def _convert_pointclouds_to_tensor(p):
    if _is_pointclouds(p):
        return p.points_padded(), p.num_points_per_cloud()
    if torch.is_tensor(p):
        return p, p.shape[1] * torch.ones((p.shape[0],), dtype=torch.int64, device=p.device)
    raise ValueError("The inputs X, Y should be either Pointclouds objects or tensors.")


_LOADED = {}


def load(frnn_module=None):
    """Install the hooks and import the reference's DSS.models.levelset_sampling and
    DSS.utils.point_processing.  `frnn_module`: object providing frnn_grid_points / frnn_gather
    (the reference's frnn refuses CPU tensors, frnn.py:255-256; tests inject a CPU stand-in
    built on the reference's own frnn_bf_cpu).  Returns a namespace of the loaded modules."""
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s not present" % REF)
    if "mods" not in _LOADED:
        sys.meta_path.insert(0, _RefFinder())
        sys.meta_path.append(_StubFinder())
        import pytorch3d.structures as s3d   # stubs
        import pytorch3d.ops as o3d
        import pytorch3d.ops.utils as o3du
        import pytorch3d.ops.knn as o3dk
        from collections import namedtuple
        s3d.list_to_packed = _list_to_packed
        s3d.padded_to_list = _padded_to_list
        s3d.list_to_padded = _list_to_padded
        o3d.packed_to_padded = _packed_to_padded
        o3d.padded_to_packed = _padded_to_packed
        o3d.convert_pointclouds_to_tensor = _convert_pointclouds_to_tensor
        o3d.is_pointclouds = _is_pointclouds
        o3du.convert_pointclouds_to_tensor = _convert_pointclouds_to_tensor
        o3du.is_pointclouds = _is_pointclouds
        o3dk._KNN = namedtuple("KNN", "dists idx knn")
        ls = importlib.import_module("DSS.models.levelset_sampling")
        pp = importlib.import_module("DSS.utils.point_processing")
        mh = importlib.import_module("DSS.utils.mathHelper")
        _LOADED["mods"] = types.SimpleNamespace(levelset_sampling=ls, point_processing=pp, mathHelper=mh)
    mods = _LOADED["mods"]
    if frnn_module is not None:
        mods.levelset_sampling.frnn = frnn_module
        mods.point_processing.frnn = frnn_module
    elif not isinstance(sys.modules.get("frnn"), _StubModule) and "frnn" in sys.modules:
        mods.levelset_sampling.frnn = sys.modules["frnn"]      # use_natives() after an earlier load()
        mods.point_processing.frnn = sys.modules["frnn"]
    return mods


def use_natives(which):
    """Choose the native code under the reference's Python BEFORE load():
      "reference": the reference's own `frnn` Python package (external/FRNN/frnn/{__init__,frnn}.py) on the
                   reference's own compiled extensions from oracle/_ref (frnn._C, prefix_sum, DSS._C) -- the
                   reference-GPU arm;
      "isob200":   isopoints_b200.install() -- the same three import names backed by libisob200.so (the drop-in)."""
    if which == "isob200":
        from isopoints_b200 import install
        return install.install()
    assert which == "reference", which
    from oracle import ref_native
    if not ref_native.available():
        raise RuntimeError("oracle/_ref natives not built")
    sys.modules["frnn._C"] = ref_native.frnn_C()
    sys.modules["prefix_sum"] = ref_native.prefix_sum()
    sys.modules["DSS._C"] = ref_native.dss_C()
    pkg = os.path.join(REF, "external", "FRNN", "frnn", "__init__.py")
    spec = importlib.util.spec_from_file_location("frnn", pkg, submodule_search_locations=[os.path.dirname(pkg)])
    mod = importlib.util.module_from_spec(spec)
    mod._C = sys.modules["frnn._C"]
    sys.modules["frnn"] = mod
    spec.loader.exec_module(mod)
    return {"frnn": mod, "prefix_sum": sys.modules["prefix_sum"], "DSS._C": sys.modules["DSS._C"]}


def load_rasterizer():
    """Import the reference's DSS.core.rasterizer (for SurfaceSplatting._get_per_point_info, the renderable
    filters and -- when use_natives() registered a real DSS._C -- rasterize_elliptical_points /
    EllipticalRasterizer).  Without natives DSS._C is stubbed: the per-point parameter code never calls it.
    pytorch3d.ops.knn_points / eyes get pure-torch stand-ins (documented semantics: K smallest squared distances
    ascending; a batch of identity matrices), kMaxPointsPerBin its upstream value 22."""
    load()
    if "rasterizer" not in _LOADED:
        import DSS
        if isinstance(sys.modules.get("DSS._C"), _StubModule) or "DSS._C" not in sys.modules:
            sys.modules["DSS._C"] = _StubModule("DSS._C")
        DSS._C = sys.modules["DSS._C"]
        import pytorch3d.ops as o3d
        import pytorch3d.renderer.points.rasterize_points as rp3d
        rp3d.kMaxPointsPerBin = 22

        def knn_points(p1, p2, lengths1=None, lengths2=None, K=1, **kw):
            from collections import namedtuple
            d = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
            if lengths2 is not None:
                col = torch.arange(p2.shape[1])[None, None, :]
                d = d.masked_fill(col >= lengths2[:, None, None], float("inf"))
            dist, idx = torch.topk(d, K, dim=-1, largest=False, sorted=True)
            return namedtuple("KNN", "dists idx knn")(dist, idx, None)

        def eyes(dim, N, device=None, dtype=torch.float32):
            return torch.eye(dim, device=device, dtype=dtype)[None].repeat(N, 1, 1)

        o3d.knn_points = knn_points
        o3d.eyes = eyes
        _LOADED["rasterizer"] = importlib.import_module("DSS.core.rasterizer")
    return _LOADED["rasterizer"]
