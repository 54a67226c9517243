"""system.json -> tables / coefficients / axes / events (casmcode_clexmonte_b200/system.py)."""
import json

import numpy as np
import pytest

from conftest import GOLDEN, REFERENCE

from casmcode_clexmonte_b200.system import SystemError_, composition_axes, load_system


def test_composition_axes_forms():
    """Vector and column-matrix end members (FCC and ZrO fixtures of the reference)."""
    fcc = composition_axes({"a": [0.0, 1.0, 0.0], "b": [0.0, 0.0, 1.0], "components": ["A", "B", "Va"],
                            "independent_compositions": 2, "origin": [1.0, 0.0, 0.0]})
    assert np.allclose(fcc["Rt"], [[-1 / 3, 2 / 3, -1 / 3], [-1 / 3, -1 / 3, 2 / 3]])
    zro = composition_axes({"a": [[2.0], [0.0], [2.0]], "components": ["Zr", "Va", "O"],
                            "independent_compositions": 1, "origin": [[2.0], [2.0], [0.0]]})
    assert np.allclose(zro["Rt"], [[0.0, -0.25, 0.25]])
    assert zro["origin"] == [2.0, 2.0, 0.0] and zro["end_members"] == [[2.0, 0.0, 2.0]]
    with pytest.raises(SystemError_):
        composition_axes({"components": ["A"], "independent_compositions": 1, "origin": [1.0]})


def test_loader_errors(tmp_path):
    with pytest.raises(SystemError_, match="prim"):
        load_system({"composition_axes": {}})
    base = {"prim": {"basis": [{"coordinate": [0, 0, 0], "occupants": ["A", "X"]}]},
            "composition_axes": {"a": [0.0, 1.0], "components": ["A", "B"], "independent_compositions": 1,
                                 "origin": [1.0, 0.0]}}
    with pytest.raises(SystemError_, match="occupant 'X'"):
        load_system(base)
    base["prim"]["basis"][0]["occupants"] = ["A", "B"]
    base["basis_sets"] = {"default": {"source": "nowhere/clexulator.cc"}}
    with pytest.raises(SystemError_, match="basis_sets/default/source"):
        load_system(base, search_path=[tmp_path])
    del base["basis_sets"]
    base["clex"] = {"formation_energy": {"basis_set": "default", "coefficients": "eci.json"}}
    with pytest.raises(SystemError_, match="basis_set 'default'"):
        load_system(base)
    del base["clex"]
    s = load_system(base)
    assert s.occ_to_species == [[0, 1]] and s.sublat_to_asym == [0] and s.mutable_sublats == [0]


@pytest.mark.skipif(not REFERENCE.exists(), reason="the reference's fixtures are not on this machine")
def test_reference_systems_load_to_the_committed_facts(load_tables):
    """The reference's own system files give what tests/golden/systems.json was assembled
    from by hand in round 1, and the same tables the committed .npz exports hold."""
    g = json.loads((GOLDEN / "systems.json").read_text())
    s = load_system(REFERENCE / "python/tests/data/FCC_binary_vacancy/system.json")
    f = g["fcc"]
    assert s.components == f["species"] and s.occ_to_species == f["occ_to_species"]
    assert s.sublat_to_asym == f["sublat_to_asym"] and s.mutable_sublats == f["mutable_sublats"]
    assert np.allclose(s.axes["Rt"], f["axes"]["Rt"]) and s.axes["origin"] == f["axes"]["origin"]
    t, ref = s.basis_sets["default"], load_tables("fcc_default")
    for name in ("nbr", "phi", "term_coef", "factor_n", "factor_f", "global_gbeg", "delta_gbeg"):
        assert np.array_equal(getattr(t, name), getattr(ref, name)), name
    assert s.clex["formation_energy"]["basis_set"] == "default"
    assert "_kmc_events" in s.extra and not s.event_types          # "_" keys are comments
    k = load_system(REFERENCE / "python/tests/data/FCC_binary_vacancy/kmc_system.json")
    assert [e["name"] for e in k.event_types] == [e["name"] for e in g["fcc"]["kmc"]["event_types"]]
    for e, ge in zip(k.event_types, g["fcc"]["kmc"]["event_types"]):
        assert [[list(x) for x in ev["sites"]] for ev in e["events"]] == [ev["sites"] for ev in ge["events"]]
        assert e["kra"][1].tolist() == ge["kra"]["value"] and e["freq"][1].tolist() == ge["freq"]["value"]
        assert len(k.local_basis_sets[e["local_basis_set"]]["tables"]) == 6


@pytest.mark.skipif(not REFERENCE.exists(), reason="the reference's fixtures are not on this machine")
def test_exported_tables_match_the_projects_basis_json(load_tables):
    """the project's basis.json (the reference's own description of the basis set)
    agrees with the tables exported from the generated source -- number of functions, occupants,
    site basis functions, orbit, cluster size and multiplicity of every function -- and the
    checker names what differs when a table is tampered with."""
    import copy

    from casmcode_clexmonte_b200.clexulator_tables import check_tables_against_basis, read_basis_json
    path = REFERENCE / "tests/unit/clexmonte/data/FCC_binary_vacancy/basis_sets/bset.default/basis.json"
    basis = read_basis_json(path)
    assert basis["occupants"] == [["A", "B", "Va"]]
    assert [(o["index"], o["mult"], len(o["sites"]), o["functions"]) for o in basis["orbits"]] == \
        [(0, 1, 0, [0]), (1, 1, 1, [1, 2]), (2, 6, 2, [3, 4, 5]), (3, 3, 2, [6, 7, 8])]
    t = load_tables("fcc_default")
    check_tables_against_basis(t, basis)
    check_tables_against_basis(t, path)
    bad = copy.copy(t)
    bad.phi = t.phi.copy()
    bad.phi[0, 1, 2] = 0.5
    bad.group_div = t.group_div.copy()
    bad.group_div[int(t.global_gbeg[3])] = 4.0
    with pytest.raises(ValueError) as e:
        check_tables_against_basis(bad, basis)
    assert "phi_{0,1}" in str(e.value) and "multiplicity 6" in str(e.value)
    # clust.json (the orbits without the functions) checks the same tables through the source's
    # own function -> orbit assignment and the prototype clusters
    check_tables_against_basis(t, path.with_name("clust.json"))
    bad2 = copy.copy(t)
    bad2.orbit_nbhd = t.orbit_nbhd.copy()
    bad2.orbit_nbhd[int(t.orbit_nbhd_beg[6]):int(t.orbit_nbhd_beg[7]), 1:] += 5
    with pytest.raises(ValueError) as e:
        check_tables_against_basis(bad2, path.with_name("clust.json"))
    assert "site neighborhood" in str(e.value)
    # the synthetic basis of configs[0] has no basis.json: a wrong one is refused
    with pytest.raises(ValueError):
        check_tables_against_basis(load_tables("fcc_synthetic"), basis)
    # load_system runs the same check when a basis set names its basis.json
    # (System_json_io.cc: basis_sets/<name>/{source, basis})
    root = REFERENCE / "python/tests/data/FCC_binary_vacancy"
    data = json.loads((root / "system.json").read_text())
    data["basis_sets"]["default"]["basis"] = str(path)
    s = load_system(data, search_path=[root])
    assert s.basis_sets["default"].corr_size == 9
    data["basis_sets"]["default"]["basis"] = "basis_sets/bset.A_Va_1NN/basis.json"   # another basis set's description
    with pytest.raises(SystemError_) as e:
        load_system(data, search_path=[root])
    assert "basis_sets/default" in str(e.value) and "basis.json" in str(e.value)
