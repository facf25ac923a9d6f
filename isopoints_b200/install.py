"""Zero-line drop-in: register this package's natives under the module names the reference imports.

The reference reaches its native code through three import names:

    import frnn                                  DSS/models/levelset_sampling.py:9, DSS/utils/point_processing.py,
                                                 DSS/core/rasterizer.py:17 (also ``frnn._C.insert_points_cuda`` /
                                                 ``counting_sort_cuda``, :909-929)
    from prefix_sum import prefix_sum_cuda       external/FRNN/frnn/frnn.py:11, DSS/core/rasterizer.py:873
    from .. import _C   (= ``DSS._C``)           DSS/core/rasterizer.py:21 (``splat_points``, ``_backward_zbuf``,
                                                 ``_splat_points_occ_backward``, ``_splat_points_occ_fast_cuda_backward``)

``install()`` puts module objects with exactly those names and attributes into ``sys.modules`` -- backed by
``libisob200.so`` -- so that the reference's Python (``train_mvr.py`` and everything under ``DSS/``) runs unchanged
on this library instead of its own three compiled extensions: call it once before ``import DSS``.  Nothing of the
reference is edited; ``uninstall()`` removes the entries again.

``install(operators=True)`` additionally replaces, after ``DSS`` has been imported, the Python operators whose
whole loop lives in kernels here (``UniformProjection``, ``EdgeAwareProjection``, ``SphereTracing``,
``sample_uniform_iso_points``, ``wlop``, ``upsample``, ``resample_uniformly``, ``rasterize_elliptical_points``,
``EllipticalRasterizer``) by attribute assignment on the reference's own modules.
"""
import sys
import types

_NAMES = ("frnn", "frnn.frnn", "prefix_sum", "DSS._C")
_saved = {}


def _frnn_modules():
    from . import frnn as f
    top = types.ModuleType("frnn")
    top.__doc__ = "isopoints_b200 drop-in for external/FRNN/frnn (frnn/__init__.py:2)"
    top.__path__ = []          # a package: `from frnn.frnn import ...` resolves the entry below
    sub = types.ModuleType("frnn.frnn")
    for m in (top, sub):
        m.frnn_grid_points = f.frnn_grid_points
        m.frnn_gather = f.frnn_gather
        m._C = f._C
    sub._GRID = f._GRID
    sub.prefix_sum_cuda = f.prefix_sum_cuda
    # constants other code imports from frnn.frnn (frnn.py:55-60; DSS/core/rasterizer.py:871)
    sub.GRID_PARAMS_SIZE = 8
    sub.MAX_RES = 128
    top.frnn = sub
    return top, sub


def _prefix_sum_module():
    from . import frnn as f
    m = types.ModuleType("prefix_sum")
    m.__doc__ = "isopoints_b200 drop-in for external/FRNN/external/prefix_sum"
    m.prefix_sum_cuda = f.prefix_sum_cuda
    return m


def _dss_c_module():
    from . import splat
    m = types.ModuleType("DSS._C")
    m.__doc__ = "isopoints_b200 drop-in for the bound subset of DSS._C (DSS/csrc/ext.cpp:5-18)"
    for name in ("splat_points", "_splat_points_naive", "_backward_zbuf", "_splat_points_occ_backward",
                 "_splat_points_occ_fast_cuda_backward"):
        setattr(m, name, getattr(splat._C, name))
    return m


def install(operators=False):
    """Register ``frnn``, ``frnn.frnn``, ``prefix_sum`` and ``DSS._C`` in ``sys.modules`` (idempotent).  Fails
    loudly (ImportError) when ``libisob200.so`` is missing: there is no fallback to fall back to."""
    from . import _ext
    _ext.lib()
    top, sub = _frnn_modules()
    mods = {"frnn": top, "frnn.frnn": sub, "prefix_sum": _prefix_sum_module(), "DSS._C": _dss_c_module()}
    for name, m in mods.items():
        if name not in _saved:
            _saved[name] = sys.modules.get(name)
        sys.modules[name] = m
    dss = sys.modules.get("DSS")
    if dss is not None:           # DSS already imported: `from .. import _C` reads the package attribute first
        dss._C = mods["DSS._C"]
    if operators:
        install_operators()
    return mods


def install_operators():
    """Swap the reference's Python operators for this package's (same names, same signatures) on the already
    imported reference modules."""
    import importlib
    from . import levelset_sampling as ls, point_processing as pp, splat
    ref_ls = importlib.import_module("DSS.models.levelset_sampling")
    for name in ("UniformProjection", "EdgeAwareProjection", "SphereTracing", "sample_uniform_iso_points",
                 "ProjectionResult"):
        setattr(ref_ls, name, getattr(ls, name))
    ref_pp = importlib.import_module("DSS.utils.point_processing")
    for name in ("wlop", "upsample", "resample_uniformly", "farthest_sampling"):
        setattr(ref_pp, name, getattr(pp, name))
    ref_ls.upsample, ref_ls.wlop = pp.upsample, pp.wlop
    try:
        ref_r = importlib.import_module("DSS.core.rasterizer")
        ref_r.rasterize_elliptical_points = splat.rasterize_elliptical_points
        ref_r.EllipticalRasterizer = splat.EllipticalRasterizer
    except Exception:      # pytorch3d absent: the rasteriser module of the reference does not import
        pass


def uninstall():
    for name in _NAMES:
        if name in _saved:
            old = _saved.pop(name)
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


class _CallableModule(types.ModuleType):
    """``isopoints_b200.install()`` is the spelling the documentation uses: calling the module runs ``install``."""

    def __call__(self, *args, **kwargs):
        return install(*args, **kwargs)


sys.modules[__name__].__class__ = _CallableModule
