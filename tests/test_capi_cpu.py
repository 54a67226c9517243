"""CPU tests of the C-ABI library: it loads, exports every declared symbol and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import re
from pathlib import Path

import numpy as np
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


def test_header_constants_match_the_binding():
    """Every status code, sweep flag and state option of include/cmx_b200.h has the same value
    in the ctypes binding (the plugin and the Python host must mean the same bits)."""
    header = (ROOT / "include" / "cmx_b200.h").read_text()
    defines = {m.group(1): int(m.group(2)) for m in
               re.finditer(r"^#define\s+(CMX_[A-Z0-9_]+)\s+(\d+)u?\b", header, flags=re.M)}
    names = [n for n in defines if n.startswith(("CMX_SWEEP_", "CMX_ERR_", "CMX_STATE_")) or n == "CMX_OK"]
    assert {"CMX_SWEEP_DE_SUM", "CMX_SWEEP_FORCE_GENERIC", "CMX_SWEEP_STREAM", "CMX_SWEEP_THREAD_GENERIC",
            "CMX_SWEEP_PAIR_SUM", "CMX_STATE_LINEAR_ROWS", "CMX_OK", "CMX_ERR_CUDA"} <= set(names)
    for n in names:
        assert getattr(_capi, n) == defines[n], n
    flags = [defines[n] for n in names if n.startswith("CMX_SWEEP_")]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags)   # distinct single bits


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


def test_supercell_box_is_the_same_lattice():
    """Hermite normal form of transformation matrices (host arithmetic): lower triangular,
    reduced, and T^-1 H unimodular -- H spans the lattice T spans.  10 * fcc_conventional
    (tests/unit/teststructures.hh:12-16 times 10) is the box of the reference's KMC tests."""
    from casmcode_clexmonte_b200 import _capi
    conv = np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1]])
    assert _capi.supercell_box(conv * 10) == (10, 20, 20, 10, 10, 0)
    assert _capi.supercell_box(np.diag([4, 6, 8])) == (4, 6, 8, 0, 0, 0)
    rng = np.random.default_rng(5)
    n = 0
    while n < 200:
        T = rng.integers(-6, 7, (3, 3))
        det = int(round(np.linalg.det(T)))
        if det == 0:
            with pytest.raises(_capi.CmxError):
                _capi.supercell_box(T)
            continue
        n += 1
        N0, N1, N2, s10, s20, s21 = _capi.supercell_box(T)
        assert N0 * N1 * N2 == abs(det)
        assert 0 <= s10 < N1 and 0 <= s20 < N2 and 0 <= s21 < N2
        H = np.array([[N0, 0, 0], [s10, N1, 0], [s20, s21, N2]], dtype=float)
        U = np.linalg.solve(T.astype(float), H)
        assert np.allclose(U, np.round(U), atol=1e-9) and abs(abs(np.linalg.det(U)) - 1) < 1e-9
