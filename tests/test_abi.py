"""The C-ABI library loads and exports every symbol include/isob200.h declares (no GPU needed)."""
import ctypes
import os
import re

from isopoints_b200 import _ext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "isob200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(isob200_\w+)\s*\(", txt)))


def test_header_and_loader_agree():
    assert _declared() == _ext.exported_symbols()


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_ext.LIB_PATH), "build with `python -m isopoints_b200.build`"
    h = ctypes.CDLL(_ext.LIB_PATH)
    for name in _declared():
        assert hasattr(h, name), name
    lib = _ext.lib()
    assert lib.isob200_abi_version() >= 1
    assert lib.isob200_compiled_arch() == 1000      # sm_100a only
    assert lib.isob200_exclusive_scan_ws_bytes(1 << 20, 2) > 0


def test_no_cpu_fallback():
    """Product code never imports the oracle and refuses CPU tensors."""
    import pytest
    import torch
    from isopoints_b200 import frnn
    with pytest.raises(TypeError):
        frnn.frnn_grid_points(torch.rand(1, 8, 3), torch.rand(1, 8, 3), K=2, r=0.1)
    pkg = os.path.join(ROOT, "isopoints_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
