"""Host logic of the run series (casmcode_clexmonte_b200/run_series.py)."""
import numpy as np
import pytest

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
