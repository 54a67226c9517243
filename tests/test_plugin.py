"""The reference-side plugins (plugin/B200SemiGrandCanonicalCalculator.cc,
plugin/B200CanonicalCalculator.cc): C++ host code against the reference's plugin interface,
calling the CUDA library through the C ABI only.  They are compiled with g++ against the
stand-in headers of plugin/shim (libcasm is not in this image) and driven by
plugin/test_plugin*.cc the way MonteCalculator drives a plugin: dlopen, make_<Name>(),
reset(params, system), run(state, occ_location, run_manager)."""
import subprocess
from pathlib import Path

import pytest

from conftest import GOLDEN

ROOT = Path(__file__).resolve().parents[1]
PLUGIN = ROOT / "plugin"


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    from casmcode_clexmonte_b200 import build
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    build.build()
    r = subprocess.run(["make", "-C", str(PLUGIN)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    flat = tmp_path_factory.mktemp("plugin") / "fcc_default.cmxt"
    ClexulatorTables.load(GOLDEN / "tables" / "fcc_default.npz").save_flat(flat)
    return flat


def test_plugin_builds_and_exports_the_factory(built):
    """CPU: the plugin source compiles, exports the C-linkage factory the reference looks up
    ("make_" + name, MonteCalculator.cc:163-166), and the calculator it returns declares its
    requirements and rejects params without the tables (no GPU needed)."""
    so = PLUGIN / "_build" / "libB200SemiGrandCanonicalCalculator.so"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(so)], capture_output=True, text=True).stdout
    assert " T make_B200SemiGrandCanonicalCalculator" in nm
    # the only undefined symbols that are not libc / libstdc++ are the C ABI's
    und = subprocess.run(["nm", "-D", "--undefined-only", str(so)], capture_output=True, text=True).stdout
    cmx = sorted({line.split()[-1] for line in und.splitlines() if " cmx_" in line})
    assert "cmx_sgc_sweep" in cmx and "cmx_delta_e" in cmx and "cmx_tables_create_from_file" in cmx
    # ... and the reference's own calculator of the ensemble (libcasm_clexmonte; a stand-in here),
    # from which the standard sampling / analysis functions are handed through
    assert " U make_SemiGrandCanonicalCalculator" in und
    r = subprocess.run([str(PLUGIN / "_build" / "test_plugin"), str(so), str(built), "--no-gpu"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "plugin ok" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["", "_bulk"])
def test_plugin_run_equals_the_c_abi(built, variant):
    """GPU: one run() through the plugin (sampling fixture every 2 passes, counters through the
    RunManager calls of occupation_metropolis.hh:109-116 -- one by one, or with the proposed
    bulk call) leaves the occupation, the acceptance count and the potential the C ABI gives
    for the same seed; the potential's occ_delta / per_supercell are served by the device."""
    so = PLUGIN / "_build" / f"libB200SemiGrandCanonicalCalculator{variant}.so"
    r = subprocess.run([str(PLUGIN / "_build" / "test_plugin"), str(so), str(built)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "plugin ok: 10 passes" in r.stdout
    # a skewed supercell whose unit cells the caller numbers in its own order: the plugin asks
    # Conversions::l_to_ijk; a B-Va pair interacts exactly when it is a first-neighbour pair
    assert "general supercell ok" in r.stdout


def test_canonical_plugin_builds_and_exports_the_factory(built):
    """CPU: the canonical calculator compiles, exports make_B200CanonicalCalculator and binds
    the pair-exchange entry points of the C ABI."""
    so = PLUGIN / "_build" / "libB200CanonicalCalculator.so"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(so)], capture_output=True, text=True).stdout
    assert " T make_B200CanonicalCalculator" in nm
    und = subprocess.run(["nm", "-D", "--undefined-only", str(so)], capture_output=True, text=True).stdout
    cmx = sorted({line.split()[-1] for line in und.splitlines() if " cmx_" in line})
    for sym in ("cmx_canonical_sweep", "cmx_canonical_default_swaps", "cmx_canonical_set_swaps", "cmx_delta_e",
                "cmx_state_set_site_order"):
        assert sym in cmx
    r = subprocess.run([str(PLUGIN / "_build" / "test_plugin_canonical"), str(so), str(built), "--no-gpu"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "canonical plugin ok" in r.stdout


@pytest.mark.gpu
def test_canonical_plugin_run_equals_the_c_abi(built):
    """GPU: the canonical calculator's potential is the formation energy of the C ABI, a
    two-site delta equals the difference of two evaluations, validate_state enforces the
    conditions' composition (CanonicalCalculator.cc:318-360), and one run() conserves the
    composition and leaves the occupation cmx_canonical_sweep gives for the same seed and the
    default swap table."""
    so = PLUGIN / "_build" / "libB200CanonicalCalculator.so"
    r = subprocess.run([str(PLUGIN / "_build" / "test_plugin_canonical"), str(so), str(built)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "canonical plugin ok: 6 passes" in r.stdout
