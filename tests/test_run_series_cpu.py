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
