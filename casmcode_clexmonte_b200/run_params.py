"""Run parameters of the reference's command-line programs (ccasm_clexmonte_canonical /
semigrand_canonical: `run_params.json`) read into what `run_series.run_series` takes, so an
existing input deck drives the device path unchanged.

Follows the reference's parsers, key by key:

  RunParams                 include/casm/clexmonte/run/io/json/RunParams_json_io_impl.hh:38-130
      "state_generation", "sampling_fixtures" (each value an object or the name of a file that
      holds one), "before_first_run" / "before_each_run" (not supported here), "global_cutoff"
  state generation          src/casm/clexmonte/run/io/json/StateGenerator_json_io.cc
      method "incremental" (IncrementalConditionsStateGenerator.hh:60-135): "initial_configuration"
      (method "fixed": "transformation_matrix_to_supercell", optional "dof"), "initial_conditions",
      "conditions_increment", "n_states", "dependent_runs", "modifiers"
  conditions                src/casm/clexmonte/state/io/json/parse_conditions.cc:45-66
      "temperature"; "param_chem_pot" / "param_composition" as {"a": .., "b": ..} or an array over
      the composition axes; "mol_composition" as {"A": ..} or an array over the components
  sampling fixture          [EXT] libcasm-monte SamplingFixtureParams_json_io
      "sampling": {"sample_by", "spacing", "begin", "period", "quantities", "sample_trajectory"},
      "completion_check": {"cutoff": {"count": {"min", "max"}, ...}, "convergence": [...]},
      "results_io": {"method": "json", "kwargs": {"output_dir", ...}}

What the device path runs is a fixed number of samples per state (`Sampler.run`): the count
cutoff.  Convergence / equilibration checks of the completion check are reported in
`fixture["not_applied"]` and not evaluated; anything else this reader does not implement is an
error naming the key, never silently dropped.  Host code only.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from .results_io import RunDataOutputParams

_PARAM_NAMES = "abcdefghijklmnopqrstuvwxyz"


class RunParamsError(ValueError):
    """Run parameters this reader cannot use (the message names the key)."""


def _live(d: Dict) -> Dict:
    """keys starting with "_" are comments in the reference's input files"""
    return {k: v for k, v in d.items() if not str(k).startswith("_")}


def _vector(value, names: Sequence[str], what: str) -> List[float]:
    if isinstance(value, dict):
        v = _live(value)
        unknown = sorted(set(v) - set(names))
        if unknown:
            raise RunParamsError(f"{what}: {unknown} not among {list(names)}")
        missing = [n for n in names if n not in v]
        if missing:
            raise RunParamsError(f"{what}: missing {missing}")
        return [float(v[n]) for n in names]
    arr = np.asarray(value, dtype=float).reshape(-1)
    if arr.size != len(names):
        raise RunParamsError(f"{what}: {arr.size} values for {list(names)}")
    return arr.tolist()


def parse_conditions(data: Dict, axes: Dict, what: str = "conditions") -> Dict:
    """parse_conditions.cc:45-66 for the conditions the Metropolis calculators read.  `axes`:
    system.composition_axes(...) (components, Rt)."""
    params = list(_PARAM_NAMES[:len(axes["Rt"])])
    out: Dict = {}
    for key, value in _live(data).items():
        if key == "temperature":
            out[key] = float(value)
        elif key in ("param_chem_pot", "param_composition"):
            out[key] = _vector(value, params, f"{what}/{key}")
        elif key == "mol_composition":
            out[key] = _vector(value, axes["components"], f"{what}/{key}")
        else:
            raise RunParamsError(f"{what}/{key}: this condition is outside the hot path (SURVEY 8g)")
    return out


def parse_sampling_fixture(name: str, data, roots: Sequence[Path]) -> Dict:
    if isinstance(data, str):
        path = next((r / data for r in roots if (r / data).exists()), None)
        if path is None:
            raise RunParamsError(f"sampling_fixtures/{name}: file {data!r} not found (searched {[str(r) for r in roots]})")
        data = json.loads(path.read_text())
    data = _live(data)
    s = _live(data.get("sampling") or {})
    if s.get("sample_by", "pass") != "pass":
        raise RunParamsError(f"sampling_fixtures/{name}/sampling/sample_by: only \"pass\" (count-based sampling by passes)")
    if s.get("spacing", "linear") != "linear":
        raise RunParamsError(f"sampling_fixtures/{name}/sampling/spacing: only \"linear\"")
    if s.get("sample_trajectory", False):
        raise RunParamsError(f"sampling_fixtures/{name}/sampling/sample_trajectory: trajectories are not kept on the device")
    period = int(s.get("period", 1))
    begin = int(s.get("begin", 0))
    if period < 1 or begin < 0:
        raise RunParamsError(f"sampling_fixtures/{name}/sampling: period >= 1 and begin >= 0 required")
    quantities = [str(q) for q in s.get("quantities", [])]
    cc = _live(data.get("completion_check") or {})
    count = _live(_live(cc.get("cutoff") or {}).get("count") or {})
    if count.get("max") is None:
        raise RunParamsError(f"sampling_fixtures/{name}/completion_check/cutoff/count/max: required "
                             "(the device path runs a fixed number of passes per state)")
    extra_cutoffs = sorted(set(_live(cc.get("cutoff") or {})) - {"count"})
    if extra_cutoffs:
        raise RunParamsError(f"sampling_fixtures/{name}/completion_check/cutoff: {extra_cutoffs} not supported (count only)")
    not_applied = [f"convergence of {c.get('quantity')!r} to {c.get('precision', c.get('abs_precision'))}"
                   for c in cc.get("convergence") or []]
    io = _live(data.get("results_io") or {})
    if io and io.get("method", "json") != "json":
        raise RunParamsError(f"sampling_fixtures/{name}/results_io/method: only \"json\"")
    kw = _live(io.get("kwargs") or {})
    max_count = int(count["max"])
    return dict(name=name, sample_period=period, begin=begin, quantities=quantities,
                max_count=max_count, min_count=None if count.get("min") is None else int(count["min"]),
                # samples taken while `begin + k period <= max_count` passes have been done, k >= 1
                n_samples=max(0, (max_count - begin) // period),
                with_corr=any(q.endswith("_corr") or q.startswith("corr") for q in quantities),
                output_dir=kw.get("output_dir"), write_observations=bool(kw.get("write_observations", False)),
                not_applied=not_applied)


def read_run_params(source, axes: Dict, search_path: Sequence = ()) -> Dict:
    """source: path of a run_params.json or the parsed dict.  Returns the arguments of a run
    series: transformation matrix, initial occupation (None: the default configuration, every
    site its first occupant), conditions path, fixtures."""
    if isinstance(source, (str, Path)):
        path = Path(source)
        data = json.loads(path.read_text())
        roots = [path.resolve().parent] + [Path(p) for p in search_path]
    else:
        data = dict(source)
        roots = [Path(p) for p in search_path] or [Path.cwd()]
    data = _live(data)
    for key in ("before_first_run", "before_each_run", "random_number_generator"):
        if data.get(key):
            raise RunParamsError(f"{key}: not supported by the device run series")
    unknown = sorted(set(data) - {"state_generation", "sampling_fixtures", "global_cutoff", "before_first_run",
                                   "before_each_run", "random_number_generator"})
    if unknown:
        raise RunParamsError(f"unknown run parameters {unknown}")
    sg = _live(data.get("state_generation") or {})
    if sg.get("method") != "incremental":
        raise RunParamsError("state_generation/method: only \"incremental\" (IncrementalConditionsStateGenerator)")
    kw = _live(sg.get("kwargs") or {})
    ic = _live(kw.get("initial_configuration") or {})
    if ic.get("method") != "fixed":
        raise RunParamsError("state_generation/kwargs/initial_configuration/method: only \"fixed\" (FixedConfigGenerator)")
    ikw = _live(ic.get("kwargs") or {})
    if "transformation_matrix_to_supercell" not in ikw:
        raise RunParamsError("initial_configuration/kwargs/transformation_matrix_to_supercell: required")
    T = np.asarray(ikw["transformation_matrix_to_supercell"], dtype=np.int64)
    if T.shape != (3, 3) or round(float(np.linalg.det(T))) <= 0:
        raise RunParamsError("transformation_matrix_to_supercell: a 3 x 3 integer matrix of positive determinant")
    occupation = None
    dof = ikw.get("dof")
    if dof:
        if set(_live(dof)) - {"occ"}:
            raise RunParamsError(f"initial_configuration/kwargs/dof: only \"occ\" (found {sorted(_live(dof))})")
        occupation = np.asarray(dof["occ"], dtype=np.int32)
    if kw.get("modifiers"):
        raise RunParamsError("state_generation/kwargs/modifiers: state-modifying functions are not supported")
    for key in ("initial_conditions", "conditions_increment", "n_states"):
        if key not in kw:
            raise RunParamsError(f"state_generation/kwargs/{key}: required")
    initial = parse_conditions(kw["initial_conditions"], axes, "initial_conditions")
    increment = parse_conditions(kw["conditions_increment"], axes, "conditions_increment")
    if "temperature" not in initial:
        raise RunParamsError("initial_conditions/temperature: required")
    for key in increment:
        if key not in initial:
            raise RunParamsError(f"conditions_increment/{key}: not an initial condition")
    fixtures_in = data.get("sampling_fixtures")
    if not isinstance(fixtures_in, dict) or not _live(fixtures_in):
        raise RunParamsError("sampling_fixtures: required (at least one)")
    fixtures = {name: parse_sampling_fixture(name, f, roots) for name, f in _live(fixtures_in).items()}
    diagonal = bool((T == np.diag(np.diag(T))).all())
    return dict(transformation_matrix_to_supercell=T.tolist(), N=tuple(int(x) for x in np.diag(T)) if diagonal else None,
                occupation=occupation, initial_conditions=initial, conditions_increment=increment,
                n_states=int(kw["n_states"]), dependent_runs=bool(kw.get("dependent_runs", True)),
                fixtures=fixtures, global_cutoff=bool(data.get("global_cutoff", True)))


def run_series_from_params(tables, system, params: Dict, fixture: Optional[str] = None, clex: str = "formation_energy",
                           seed: int = 0, n_equilibration_passes: int = 0, output_dir=None,
                           ensemble: Optional[str] = None) -> List[Dict]:
    """A run series from read_run_params(...) output.  `system`: a system.System (load_system);
    `tables`: _capi.Tables of the clex's basis set.  The ensemble follows the conditions, as the
    choice of program does for the reference: "param_chem_pot" -> semi-grand canonical
    (ccasm_clexmonte_semigrand_canonical), a composition -> canonical (ccasm_clexmonte_canonical).
    One sampling fixture drives a device run: `fixture` names it when the parameters hold several."""
    from .run_series import run_series
    if params["N"] is None:
        raise RunParamsError("run series on the device need a diagonal transformation_matrix_to_supercell")
    if ensemble is None:
        ensemble = "semigrand_canonical" if "param_chem_pot" in params["initial_conditions"] else "canonical"
    if ensemble == "semigrand_canonical" and "param_chem_pot" not in params["initial_conditions"]:
        raise RunParamsError("initial_conditions/param_chem_pot: required by the semi-grand canonical run series")
    names = sorted(params["fixtures"])
    if fixture is None:
        if len(names) != 1:
            raise RunParamsError(f"several sampling fixtures {names}: name the one that drives the run")
        fixture = names[0]
    fx = params["fixtures"][fixture]
    n_sites = int(np.prod(params["N"])) * len(system.occ_to_species)
    occ = params["occupation"] if params["occupation"] is not None else np.zeros(n_sites, dtype=np.int32)
    if occ.size != n_sites:
        raise RunParamsError(f"initial occupation has {occ.size} sites, the supercell {n_sites}")
    c = system.clex[clex]
    out_dir = output_dir if output_dir is not None else fx["output_dir"]
    return run_series(tables, params["N"], system.as_dict(), c["index"], c["value"], params["initial_conditions"],
                      params["conditions_increment"], params["n_states"], occ,
                      n_equilibration_passes=n_equilibration_passes + fx["begin"], n_samples=fx["n_samples"],
                      sample_period=fx["sample_period"], seed=seed, dependent_runs=params["dependent_runs"],
                      with_corr=fx["with_corr"],
                      output_params=RunDataOutputParams(output_dir=out_dir) if out_dir else None, ensemble=ensemble)
