"""Host logic of the run series (casmcode_clexmonte_b200/run_series.py)."""
import json

import numpy as np
import pytest

from conftest import REFERENCE

from casmcode_clexmonte_b200.run_series import conditions_path, make_incremented_values


def test_incremental_conditions_path():
    """IncrementalConditionsStateGenerator: conditions_k = initial + k * increment; keys of the
    increment must exist in the initial conditions
    (include/casm/clexmonte/run/IncrementalConditionsStateGenerator.hh:90-97,111-113)."""
    init = {"temperature": 300.0, "param_chem_pot": [-1.0, 0.0]}
    inc = {"temperature": 100.0, "param_chem_pot": [0.5, 0.0]}
    path = conditions_path(init, inc, 4)
    assert [float(c["temperature"]) for c in path] == [300.0, 400.0, 500.0, 600.0]
    assert [c["param_chem_pot"].tolist() for c in path] == [[-1.0, 0.0], [-0.5, 0.0], [0.0, 0.0], [0.5, 0.0]]
    only_t = conditions_path(init, {"temperature": -50.0}, 3)
    assert [float(c["temperature"]) for c in only_t] == [300.0, 250.0, 200.0]
    assert all(c["param_chem_pot"].tolist() == [-1.0, 0.0] for c in only_t)
    with pytest.raises(ValueError):
        make_incremented_values(init, {"pressure": 1.0}, 1)
    with pytest.raises(ValueError):
        make_incremented_values(init, {"param_chem_pot": [1.0]}, 1)


# ---------------------------------------------------------------------------
# results files (casmcode_clexmonte_b200/results_io.py)
# ---------------------------------------------------------------------------
def validate_summary_data(subdata, expected_keys, expected_size):
    """The checks of python/tests/conftest.py:141-154 of the reference, restated."""
    for x in expected_keys:
        assert x in subdata
        if "component_names" in subdata[x]:
            for y in subdata[x]["component_names"]:
                assert len(subdata[x][y]) == expected_size
        elif "value" in subdata[x]:
            assert subdata[x]["shape"] == []
            assert len(subdata[x]["value"]) == expected_size
        else:
            assert len(subdata[x]) == expected_size


def validate_statistics_data(subdata, expected_keys, expected_size):
    """python/tests/conftest.py:157-176."""
    for x in expected_keys:
        assert x in subdata
        if "component_names" in subdata[x]:
            for y in subdata[x]["component_names"]:
                for z in ("mean", "calculated_precision"):
                    assert len(subdata[x][y][z]) == expected_size
        else:
            assert subdata[x]["shape"] == []
            for z in ("mean", "calculated_precision"):
                assert len(subdata[x]["value"][z]) == expected_size


def validate_summary_file(path, expected_size):
    """python/tests/conftest.py:180-240, semi-grand canonical, no convergence requested."""
    import json
    data = json.loads(path.read_text())
    validate_summary_data(data["analysis"], ["heat_capacity", "mol_susc", "param_susc", "mol_thermochem_susc",
                                             "param_thermochem_susc"], expected_size)
    validate_summary_data(data["completion_check_results"],
                          ["N_samples", "N_samples_for_statistics", "acceptance_rate", "count", "elapsed_clocktime"],
                          expected_size)
    validate_summary_data(data["conditions"], ["temperature", "param_chem_pot"], expected_size)
    validate_statistics_data(data["statistics"], ["potential_energy", "clex.formation_energy", "mol_composition",
                                                  "param_composition"], expected_size)
    return data


def test_summary_file_layout_and_append(tmp_path):
    from casmcode_clexmonte_b200.results_io import SummaryWriter, calculated_precision
    rng = np.random.default_rng(0)
    for run in range(3):
        w = SummaryWriter(tmp_path, ["Zr", "Va", "O"], ["a"])        # a new writer continues the file
        assert w.n_runs() == run
        series = {"potential_energy": rng.normal(size=50), "clex.formation_energy": rng.normal(size=50),
                  "mol_composition": rng.normal(size=(50, 3)), "param_composition": rng.normal(size=(50, 1))}
        analysis = {"heat_capacity": 1.5, "mol_susc": np.arange(9.0).reshape(3, 3), "param_susc": [[2.0]],
                    "mol_thermochem_susc": [1.0, 2.0, 3.0], "param_thermochem_susc": [4.0]}
        w.append({"temperature": 300.0 + run, "param_chem_pot": [-1.0]}, analysis, series, 50, 0.25, 5000, 0.1)
    data = validate_summary_file(tmp_path / "summary.json", 3)
    assert data["conditions"]["temperature"]["value"] == [300.0, 301.0, 302.0]
    assert data["conditions"]["param_chem_pot"]["component_names"] == ["a"]
    assert data["analysis"]["mol_susc"]["component_names"][:4] == ["Zr,Zr", "Va,Zr", "O,Zr", "Zr,Va"]
    assert data["analysis"]["mol_susc"]["Va,Zr"] == [3.0, 3.0, 3.0]      # column-major: element (1, 0)
    assert data["statistics"]["mol_composition"]["component_names"] == ["Zr", "Va", "O"]
    # precision: white noise ~ 1.96 / sqrt(n); a constant series has none
    x = rng.normal(size=20000)
    assert calculated_precision(x) == pytest.approx(1.96 / np.sqrt(20000), rel=0.1)
    assert calculated_precision(np.ones(10)) == 0.0


def test_completed_runs_save_rules_and_restart(tmp_path):
    """run/IncrementalConditionsStateGenerator.hh:134-196 and RunData_json_io.hh:13-25."""
    import json
    from casmcode_clexmonte_b200.results_io import CompletedRuns, RunDataOutputParams, state_to_json
    T = np.diag([2, 2, 2])

    def run(k):
        c = {"temperature": 300.0 + k, "param_chem_pot": [0.5 * k]}
        return {"initial_state": state_to_json(np.zeros(8), T, c), "final_state": state_to_json(np.full(8, k), T, c),
                "conditions": c, "transformation_matrix_to_supercell": T.tolist(), "n_unitcells": 8}

    # defaults: only the LAST final state is kept in memory, nothing but the run records is written
    cr = CompletedRuns(RunDataOutputParams(output_dir=str(tmp_path / "a")))
    assert cr.read() == 0
    for k in range(3):
        cr.append(run(k))
        cr.write()
    assert ["final_state" in r for r in cr.runs] == [False, False, True]
    assert all("initial_state" not in r for r in cr.runs)
    on_disk = json.loads((tmp_path / "a" / "completed_runs.json").read_text())
    assert len(on_disk) == 3 and set(on_disk[0]) == {"conditions", "transformation_matrix_to_supercell", "n_unitcells"}
    assert on_disk[2]["conditions"]["temperature"] == 302.0 and on_disk[2]["n_unitcells"] == 8
    again = CompletedRuns(RunDataOutputParams(output_dir=str(tmp_path / "a")))
    assert again.read() == 3 and again.last_final_occupation() is None     # not written: a dependent series cannot resume
    # everything saved and written
    p = RunDataOutputParams(True, True, True, True, True, str(tmp_path / "b"))
    cr = CompletedRuns(p)
    for k in range(2):
        cr.append(run(k))
        cr.write()
    again = CompletedRuns(p)
    assert again.read() == 2
    assert again.runs[0]["initial_state"]["configuration"]["dof"]["occ"] == [0] * 8
    assert (again.last_final_occupation() == 1).all()
    assert again.runs[1]["final_state"]["configuration"]["transformation_matrix_to_supercell"] == T.tolist()
    assert RunDataOutputParams.from_json({"save_all_final_states": True, "output_dir": "x"}).do_save_all_final_states
    (tmp_path / "b" / "completed_runs.json").write_text('[{"conditions": {}}]')
    with pytest.raises(ValueError):
        CompletedRuns(p).read()


_AXES = dict(components=["Zr", "Va", "O"], Rt=[[0.0, -0.5, 0.5]], origin=[2.0, 2.0, 0.0], end_members=[[2.0, 0.0, 2.0]])


def _run_params(**over):
    fixture = {"sampling": {"sample_by": "pass", "spacing": "linear", "begin": 0, "period": 5,
                            "quantities": ["potential_energy", "param_composition"], "sample_trajectory": False},
               "completion_check": {"cutoff": {"count": {"min": None, "max": 100}},
                                    "convergence": [{"quantity": "potential_energy", "precision": 0.001}]},
               "results_io": {"method": "json", "kwargs": {"output_dir": "out/thermo"}}}
    p = {"state_generation": {"method": "incremental", "kwargs": {
            "initial_configuration": {"method": "fixed", "kwargs": {
                "transformation_matrix_to_supercell": [[8, 0, 0], [0, 8, 0], [0, 0, 4]], "_dof": None}},
            "initial_conditions": {"temperature": 1000.0, "param_chem_pot": {"a": -4.0}},
            "conditions_increment": {"temperature": 0.0, "param_chem_pot": [0.4]},
            "n_states": 11, "dependent_runs": False, "modifiers": []}},
         "sampling_fixtures": {"thermo": fixture}}
    p.update(over)
    return p


def test_run_params_reader():
    """run_params.json of the reference's command-line programs -> the arguments of a run
    series (RunParams_json_io_impl.hh:38-130, StateGenerator_json_io.cc, parse_conditions.cc:45-66);
    what the device path does not do is refused by key, never dropped."""
    import copy

    from casmcode_clexmonte_b200.run_params import RunParamsError, read_run_params
    p = read_run_params(_run_params(), _AXES)
    assert p["N"] == (8, 8, 4) and p["occupation"] is None and p["n_states"] == 11 and not p["dependent_runs"]
    assert p["initial_conditions"] == {"temperature": 1000.0, "param_chem_pot": [-4.0]}
    assert p["conditions_increment"] == {"temperature": 0.0, "param_chem_pot": [0.4]}
    fx = p["fixtures"]["thermo"]
    assert (fx["sample_period"], fx["n_samples"], fx["max_count"], fx["output_dir"]) == (5, 20, 100, "out/thermo")
    assert fx["not_applied"] == ["convergence of 'potential_energy' to 0.001"] and not fx["with_corr"]
    # a skewed supercell is read (and left to the caller: run series on the device need a diagonal one)
    q = _run_params()
    q["state_generation"]["kwargs"]["initial_configuration"]["kwargs"]["transformation_matrix_to_supercell"] = \
        [[-4, 4, 4], [4, -4, 4], [4, 4, -4]]
    assert read_run_params(q, _AXES)["N"] is None
    # an explicit occupation
    q = _run_params()
    q["state_generation"]["kwargs"]["initial_configuration"]["kwargs"]["dof"] = {"occ": [0, 1] * 8}
    assert read_run_params(q, _AXES)["occupation"].tolist() == [0, 1] * 8

    def refused(mutate, key):
        q = copy.deepcopy(_run_params())
        mutate(q)
        with pytest.raises(RunParamsError) as e:
            read_run_params(q, _AXES)
        assert key in str(e.value), str(e.value)

    refused(lambda q: q["state_generation"].update(method="enumeration"), "state_generation/method")
    refused(lambda q: q["state_generation"]["kwargs"]["initial_configuration"].update(method="random"), "initial_configuration/method")
    refused(lambda q: q["state_generation"]["kwargs"].update(modifiers=["match.mol_composition"]), "modifiers")
    refused(lambda q: q["state_generation"]["kwargs"]["initial_conditions"].update(param_chem_pot={"b": 1.0}), "param_chem_pot")
    refused(lambda q: q["state_generation"]["kwargs"]["initial_conditions"].update(order_parameter_pot=[1.0]), "order_parameter_pot")
    refused(lambda q: q["state_generation"]["kwargs"]["conditions_increment"].update(mol_composition=[0, 0, 0]), "conditions_increment/mol_composition")
    refused(lambda q: q["sampling_fixtures"]["thermo"]["sampling"].update(sample_by="time"), "sample_by")
    refused(lambda q: q["sampling_fixtures"]["thermo"]["sampling"].update(spacing="log"), "spacing")
    refused(lambda q: q["sampling_fixtures"]["thermo"]["completion_check"]["cutoff"]["count"].update(max=None), "count/max")
    refused(lambda q: q["sampling_fixtures"]["thermo"]["completion_check"]["cutoff"].update(clocktime={"max": 60}), "clocktime")
    refused(lambda q: q["sampling_fixtures"]["thermo"]["results_io"].update(method="hdf5"), "results_io/method")
    refused(lambda q: q.update(before_first_run={"x": {}}), "before_first_run")
    refused(lambda q: q.update(sampling_fixtures={"thermo": "missing_file.json"}), "missing_file.json")


@pytest.mark.skipif(not REFERENCE.exists(), reason="the reference's fixtures are not on this machine")
def test_run_params_reader_on_the_reference_input_decks():
    """The ZrO input decks of the reference's own tests (tests/unit/clexmonte/data/Clex_ZrO_Occ)."""
    from casmcode_clexmonte_b200.run_params import read_run_params
    from casmcode_clexmonte_b200.system import composition_axes
    root = REFERENCE / "tests/unit/clexmonte/data/Clex_ZrO_Occ"
    axes = composition_axes(json.loads((root / "system.json").read_text())["composition_axes"])
    p = read_run_params(root / "run_params_sgc_complete.json", axes)
    assert p["N"] == (6, 6, 6) and p["n_states"] == 11 and not p["dependent_runs"]
    assert p["initial_conditions"]["param_chem_pot"] == [-4.0] and p["conditions_increment"]["param_chem_pot"] == [0.4]
    fx = p["fixtures"]["thermo"]
    assert fx["n_samples"] == 100 and fx["sample_period"] == 1 and fx["with_corr"]
    # fixtures named by file (the tests of the reference fill the "TODO" placeholders in)
    d = json.loads((root / "run_params_sgc_by_file.json").read_text())
    d["sampling_fixtures"] = {"thermo_period1": "thermo_sampling.period1.json", "thermo_period10": "thermo_sampling.period10.json"}
    q = read_run_params(d, axes, search_path=[root])
    assert {k: (v["sample_period"], v["n_samples"]) for k, v in q["fixtures"].items()} == \
        {"thermo_period1": (1, 1000), "thermo_period10": (10, 100)}
    c = read_run_params(root / "run_params_complete.json", axes)     # the canonical deck: mol_composition path
    assert c["initial_conditions"]["mol_composition"] == [2.0, 1.9, 0.1] and c["conditions_increment"]["temperature"] == 10.0


def test_run_series_from_params_maps_the_deck_onto_run_series(monkeypatch):
    """The arguments run_series receives from an input deck (the device run itself is covered
    by test_run_series_batches_the_conditions_path)."""
    from types import SimpleNamespace

    import casmcode_clexmonte_b200.run_series as RS
    from casmcode_clexmonte_b200.run_params import RunParamsError, read_run_params, run_series_from_params
    seen = {}

    def fake(tables, N, system, eci_index, eci_value, initial, increment, n_states, occupation, **kw):
        seen.update(N=N, eci=(eci_index, eci_value), initial=initial, increment=increment, n_states=n_states,
                    n_sites=occupation.size, **kw)
        return []

    monkeypatch.setattr(RS, "run_series", fake)
    system = SimpleNamespace(occ_to_species=[[0], [0], [1, 2], [1, 2]], clex={"formation_energy": {"index": [0, 1], "value": [0.5, 1.0]}},
                             as_dict=lambda: {"n_species": 3})
    params = read_run_params(_run_params(), _AXES)
    run_series_from_params(None, system, params, seed=7, n_equilibration_passes=3)
    assert seen["N"] == (8, 8, 4) and seen["n_sites"] == 8 * 8 * 4 * 4 and seen["n_states"] == 11
    assert seen["eci"] == ([0, 1], [0.5, 1.0]) and seen["initial"]["param_chem_pot"] == [-4.0]
    assert (seen["n_equilibration_passes"], seen["n_samples"], seen["sample_period"], seen["seed"]) == (3, 20, 5, 7)
    assert seen["dependent_runs"] is False and seen["with_corr"] is False
    assert seen["output_params"].output_dir == "out/thermo"
    two = _run_params()
    two["sampling_fixtures"]["second"] = two["sampling_fixtures"]["thermo"]
    with pytest.raises(RunParamsError):
        run_series_from_params(None, system, read_run_params(two, _AXES))
    run_series_from_params(None, system, read_run_params(two, _AXES), fixture="second", output_dir="elsewhere")
    assert seen["output_params"].output_dir == "elsewhere"


class _FakeState:
    """Records the calls run_series makes (the device calls themselves: tests/test_gpu_sampler.py,
    tests/test_gpu_canonical.py)."""
    log = []

    def __init__(self, tables, N, n_replicas=1):
        self.N, self.n_replicas = N, n_replicas
        self.occ = {}
        _FakeState.log.append(("state", tuple(N), n_replicas))

    def set_eci(self, i, v):
        pass

    def set_occupants(self, *a):
        pass

    def set_conditions(self, T, exch, r=0):
        _FakeState.log.append(("conditions", r, float(T), exch is None))

    def upload_occ(self, occ, r=0):
        self.occ[r] = np.asarray(occ).copy()

    def download_occ(self, r=0, dtype=np.int32):
        return self.occ[r].astype(dtype)

    def canonical_default_swaps(self):
        return [(0, 0, (1, 0, 0))]

    def canonical_set_swaps(self, swaps):
        _FakeState.log.append(("swaps", len(swaps)))

    def canonical_sweep(self, n, seed, first_sweep=0):
        _FakeState.log.append(("canonical_sweep", n, seed))

    def sgc_sweep(self, n, seed, counters=True):
        _FakeState.log.append(("sgc_sweep", n, seed))

    def synchronize(self):
        pass

    def close(self):
        pass


class _FakeSampler:
    def __init__(self, st, n_samples, origin, Rt, with_corr=False):
        self.st, self.n = st, n_samples

    def set_param_chem_pot(self, mu, r):
        _FakeState.log.append(("mu", r))

    def run(self, n_samples, period, seed, first_sweep=0, ensemble="semigrand_canonical"):
        from types import SimpleNamespace
        _FakeState.log.append(("run", n_samples, period, seed, first_sweep, ensemble))
        return [SimpleNamespace(n_attempt=100, n_accept=40) for _ in range(self.st.n_replicas)]

    def series(self, r):
        return {"potential_energy": np.linspace(-1.0, -1.1, self.n)}

    def analysis(self, r):
        return {"heat_capacity": 0.5}

    def close(self):
        pass


def test_run_series_canonical_branch(monkeypatch, tmp_path):
    """The canonical run series (BASELINE configs[0]: a temperature path at fixed composition):
    conditions without an exchange term, the library's default swap table, pair-exchange
    sweeps, the sampled run in the canonical ensemble; the composition conditions are checked
    against the configuration and may not be incremented."""
    import casmcode_clexmonte_b200.run_series as RS
    monkeypatch.setattr(RS._capi, "State", _FakeState)
    monkeypatch.setattr(RS._capi, "Sampler", _FakeSampler)
    from types import SimpleNamespace
    tables = SimpleNamespace(host=SimpleNamespace(max_occ=2))
    system = dict(occ_to_species=[[0, 1]], sublat_to_asym=[0], n_species=2, species=["A", "B"],
                  axes=dict(origin=[1.0, 0.0], Rt=[[0.0, 1.0]]))
    N = (4, 4, 2)
    occ = np.array([0, 1] * 16, dtype=np.int32)
    init = {"temperature": 900.0, "mol_composition": [0.5, 0.5]}
    inc = {"temperature": -300.0, "mol_composition": [0.0, 0.0]}
    _FakeState.log = []
    res = RS.run_series(tables, N, system, [0], [1.0], init, inc, 3, occ, n_equilibration_passes=5, n_samples=4,
                        sample_period=2, seed=11, dependent_runs=True, ensemble="canonical",
                        output_params=RS.RunDataOutputParams(output_dir=str(tmp_path)))
    assert [r["conditions"]["temperature"] for r in res] == [900.0, 600.0, 300.0]
    log = _FakeState.log
    assert [e for e in log if e[0] == "conditions"] == [("conditions", 0, T, True) for T in (900.0, 600.0, 300.0)]
    assert not [e for e in log if e[0] in ("mu", "sgc_sweep")]
    assert [e for e in log if e[0] == "canonical_sweep"] == [("canonical_sweep", 5, 11 + k) for k in range(3)]
    assert [e for e in log if e[0] == "run"] == [("run", 4, 2, 11 + k, 5, "canonical") for k in range(3)]
    assert log.count(("swaps", 1)) == 3
    assert (tmp_path / "summary.json").exists() and (tmp_path / "completed_runs.json").exists()
    # independent runs: one state with a replica per temperature
    _FakeState.log = []
    RS.run_series(tables, N, system, [0], [1.0], init, inc, 3, occ, 5, 4, ensemble="canonical")
    assert ("state", N, 3) in _FakeState.log and _FakeState.log.count(("swaps", 1)) == 1
    # composition conditions
    with pytest.raises(ValueError):
        RS.run_series(tables, N, system, [0], [1.0], {"temperature": 900.0, "mol_composition": [0.75, 0.25]}, {}, 1, occ, 1, 1,
                      ensemble="canonical")
    with pytest.raises(ValueError):
        RS.run_series(tables, N, system, [0], [1.0], init, {"mol_composition": [0.1, -0.1]}, 2, occ, 1, 1, ensemble="canonical")
    with pytest.raises(ValueError):
        RS.run_series(tables, N, system, [0], [1.0], {"temperature": 900.0, "param_chem_pot": [0.0]}, {}, 1, occ, 1, 1,
                      ensemble="canonical")
    RS.run_series(tables, N, system, [0], [1.0], {"temperature": 900.0, "param_composition": [0.5]}, {}, 1, occ, 1, 1,
                  ensemble="canonical")
    # the semi-grand branch is what it was
    _FakeState.log = []
    RS.run_series(tables, N, system, [0], [1.0], {"temperature": 900.0, "param_chem_pot": [0.1]}, {}, 1, occ, 5, 4)
    assert ("sgc_sweep", 5, 0) in _FakeState.log and ("mu", 0) in _FakeState.log
    assert ("run", 4, 1, 0, 5, "semigrand_canonical") in _FakeState.log
