"""CPU tests of the C-ABI library: it loads, exports every declared symbol and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import re
from pathlib import Path

import pytest

from casmcode_clexmonte_b200 import _capi

ROOT = Path(__file__).resolve().parents[1]


def test_header_symbols_exported():
    header = (ROOT / "include" / "cmx_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    declared = set(re.findall(r"\b(cmx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_capi.EXPORTED_SYMBOLS)
    L = ctypes.CDLL(str(_capi.LIB_PATH))
    for name in declared:
        assert hasattr(L, name), name


def test_version_and_device_count():
    assert _capi.lib().cmx_version() >= 100
    assert _capi.device_count() >= 0


def test_no_cpu_fallback(load_tables):
    """Without a GPU the product must refuse, not compute on the host."""
    if _capi.device_count() > 0:
        pytest.skip("CUDA device present")
    with pytest.raises(_capi.CmxError) as e:
        _capi.Tables(load_tables("fcc_default"))
    assert e.value.code == _capi.CMX_ERR_CUDA


def test_tables_validation_without_gpu(load_tables):
    """Argument validation happens before any device work."""
    import copy
    t = copy.copy(load_tables("fcc_default"))
    t.factor_n = t.factor_n.copy()
    t.factor_n[0] = 10 ** 6
    with pytest.raises(_capi.CmxError) as e:
        _capi.Tables(t)
    assert e.value.code == _capi.CMX_ERR_INVALID


def test_product_does_not_import_oracle():
    pkg = ROOT / "casmcode_clexmonte_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + \
            list((ROOT / "include").glob("*.h")):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f
