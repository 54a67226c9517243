import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running")


def _has_gpu():
    try:
        from casmcode_clexmonte_b200 import _capi
        return _capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def systems():
    return json.loads((GOLDEN / "systems.json").read_text())


@pytest.fixture(scope="session")
def load_tables():
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    cache = {}

    def _load(name):
        if name not in cache:
            cache[name] = ClexulatorTables.load(GOLDEN / "tables" / f"{name}.npz")
        return cache[name]

    return _load


@pytest.fixture(scope="session")
def load_vectors():
    def _load(case):
        return dict(np.load(GOLDEN / f"vectors_{case}.npz"))
    return _load


@pytest.fixture(scope="session")
def oracle():
    """The live oracle (harness + oracle/_ref), or None when it was not built."""
    try:
        from oracle import oracle as O
        O.lib()
        if not O.available("fcc_default"):
            return None
        return O
    except Exception:
        return None


CASES = {
    "fcc_sparse": ("fcc", "eci_sparse"),
    "fcc_full": ("fcc", "eci_full"),
    "zro": ("zro", "eci"),
    # the synthetic FCC binary pair + triplet basis (tests/golden/make_synthetic_clexulator.py)
    "fcc_synthetic": ("fcc_syn", "eci"),
}
