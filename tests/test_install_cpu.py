"""isopoints_b200.install(): the three import names the reference reaches its native code through resolve to
this package (CPU: names / attributes only, no compute)."""
import sys

import pytest


def test_install_registers_the_reference_import_names_and_uninstall_restores():
    from isopoints_b200 import frnn as ours, install, splat
    before = {k: sys.modules.get(k) for k in ("frnn", "frnn.frnn", "prefix_sum", "DSS._C")}
    try:
        install.install()
        import frnn                                   # levelset_sampling.py:9, rasterizer.py:17
        from frnn.frnn import _GRID, prefix_sum_cuda as ps2   # noqa: F401  (rasterizer.py:871)
        from prefix_sum import prefix_sum_cuda        # frnn.py:11, rasterizer.py:873
        assert frnn.frnn_grid_points is ours.frnn_grid_points and frnn.frnn_gather is ours.frnn_gather
        for name in ("insert_points_cuda", "counting_sort_cuda", "find_nbrs_cuda", "frnn_backward_cuda"):
            assert callable(getattr(frnn._C, name))      # ext.cpp:7-24, what rasterizer.py:909-929 reaches into
        assert prefix_sum_cuda is ours.prefix_sum_cuda
        c = sys.modules["DSS._C"]
        for name in ("splat_points", "_splat_points_naive", "_backward_zbuf", "_splat_points_occ_backward",
                     "_splat_points_occ_fast_cuda_backward"):          # DSS/csrc/ext.cpp:5-18
            assert getattr(c, name) == getattr(splat._C, name)
        install()                                     # idempotent; the module itself is callable
    finally:
        install.uninstall()
    for k, v in before.items():
        assert sys.modules.get(k) is v


def test_reference_python_binds_to_the_installed_natives():
    """The reference's own modules, unmodified on disk, import on top of the drop-in."""
    from oracle import ref_python
    if not ref_python.available():
        pytest.skip("reference tree not present")
    from isopoints_b200 import install
    try:
        mods = install.install()
        ref = ref_python.load()
        assert ref.levelset_sampling.frnn is mods["frnn"] and ref.point_processing.frnn is mods["frnn"]
        rast = ref_python.load_rasterizer()
        if not isinstance(rast._C, ref_python._StubModule):     # an earlier test may have loaded it with the stub
            assert rast._C is mods["DSS._C"]
        assert rast.frnn is mods["frnn"] or isinstance(rast.frnn, ref_python._StubModule)
    finally:
        install.uninstall()
