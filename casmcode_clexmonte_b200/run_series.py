"""Run series over an incremental conditions path, batched.

Host-side mirror of the reference's run loop for the one case that maps onto the device
in a single launch stream:

  IncrementalConditionsStateGenerator  include/casm/clexmonte/run/IncrementalConditionsStateGenerator.hh:60-135
      conditions_k = initial_conditions + k * conditions_increment,  k = 0 .. n_states-1
      dependent_runs == false: every state starts from the same configuration
      (FixedConfigGenerator, run/FixedConfigGenerator.hh:25)
  run_series                           include/casm/clexmonte/run/functions.hh:83-166

With independent runs the states of the path do not depend on each other, so they are the
replicas of ONE device state: one equilibration call and one sampled run (Sampler.run)
advance the whole path.  dependent_runs == true is inherently sequential (state k starts
from the final configuration of state k-1) and is run one state at a time on replica 0.

Results are returned per state with the reference's sampler / analysis function names and,
when an output directory is given, written as the reference writes them: one column per run
appended to summary.json, RunData appended to completed_runs.json (results_io.py); a series
started again with the same output directory reads completed_runs.json and continues after
the last completed state (run/IncrementalConditionsStateGenerator.hh:100-196).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

import time

from . import _capi
from .potential import semigrand_exchange_table
from .results_io import CompletedRuns, RunDataOutputParams, SummaryWriter, state_to_json


def make_incremented_values(initial: Dict, increment: Dict, k: int) -> Dict:
    """monte::make_incremented_values: initial + k * increment, key by key (keys of
    `increment` must be keys of `initial`, IncrementalConditionsStateGenerator.hh:90-97)."""
    for key in increment:
        if key not in initial:
            raise ValueError(f"conditions_increment key {key!r} is not an initial condition")
    out = {}
    for key, v in initial.items():
        v = np.asarray(v, dtype=np.float64)
        inc = np.asarray(increment.get(key, np.zeros_like(v)), dtype=np.float64)
        if inc.shape != v.shape:
            raise ValueError(f"Mismatch between initial conditions and conditions increment for {key!r}")
        out[key] = v + k * inc
    return out


def conditions_path(initial: Dict, increment: Dict, n_states: int) -> List[Dict]:
    return [make_incremented_values(initial, increment, k) for k in range(int(n_states))]


def mol_composition(system: Dict, occupation, n_cells: int) -> np.ndarray:
    """species per unit cell of an occupation in the reference's site order l = b n_cells + cell"""
    occ = np.asarray(occupation)
    out = np.zeros(int(system["n_species"]))
    for b, species in enumerate(system["occ_to_species"]):
        counts = np.bincount(occ[b * n_cells:(b + 1) * n_cells], minlength=len(species))
        if len(counts) > len(species):
            raise ValueError(f"occupation: occupant index {len(counts) - 1} on sublattice {b} with {len(species)} occupants")
        for o, sp in enumerate(species):
            out[sp] += counts[o] / n_cells
    return out


def check_canonical_conditions(system: Dict, initial: Dict, increment: Dict, occupation, N, tol: float = 1e-4) -> None:
    """validate_state of the canonical calculator (CanonicalCalculator.cc:318-360): the
    configuration's composition is the conditions'; and the path keeps it."""
    for key in ("mol_composition", "param_composition"):
        if key in increment and np.any(np.asarray(increment[key], dtype=float) != 0.0):
            raise ValueError(f"run_series (canonical): conditions_increment/{key} must be zero "
                             "(the path cannot change the composition of a configuration)")
    if "param_chem_pot" in initial:
        raise ValueError("run_series (canonical): param_chem_pot is a semi-grand canonical condition")
    n_cells = int(np.prod(N if not np.isscalar(N) else (N, N, N)))
    mol = mol_composition(system, occupation, n_cells)
    axes = system["axes"]
    if "mol_composition" in initial and np.abs(mol - np.asarray(initial["mol_composition"], dtype=float)).max() > tol:
        raise ValueError(f"run_series (canonical): the initial occupation has mol_composition {mol.tolist()}, "
                         f"the conditions {list(initial['mol_composition'])}")
    if "param_composition" in initial:
        param = np.asarray(axes["Rt"], dtype=float) @ (mol - np.asarray(axes["origin"], dtype=float))
        if np.abs(param - np.asarray(initial["param_composition"], dtype=float)).max() > tol:
            raise ValueError(f"run_series (canonical): the initial occupation has param_composition {param.tolist()}, "
                             f"the conditions {list(initial['param_composition'])}")


def run_series(tables: "_capi.Tables", N: Sequence[int], system: Dict, eci_index, eci_value,
               initial_conditions: Dict, conditions_increment: Dict, n_states: int, occupation: np.ndarray,
               n_equilibration_passes: int, n_samples: int, sample_period: int = 1, seed: int = 0,
               dependent_runs: bool = False, with_corr: bool = False,
               output_params: Optional[RunDataOutputParams] = None, summary_dir=None,
               ensemble: str = "semigrand_canonical") -> List[Dict]:
    """Run series in the semi-grand canonical (default) or the canonical ensemble.  `system`:
    occ_to_species, sublat_to_asym, n_species, axes {origin, Rt}.  Conditions: {"temperature": T,
    "param_chem_pot": [...]}; canonical: {"temperature": T} plus optionally "mol_composition" /
    "param_composition", which the initial occupation must have (CanonicalCalculator.cc:318-360)
    and which the path may not change (parallel pair exchanges over the library's default swap
    table conserve the composition).
    Returns one dict per state RUN BY THIS CALL: conditions, means of the sampled quantities,
    the analysis functions, acceptance rate, final occupation.  output_params.output_dir:
    completed_runs.json (read first: completed states are not run again); summary_dir
    (default: the same directory): summary.json."""
    if ensemble not in ("semigrand_canonical", "canonical"):
        raise ValueError(f"run_series: unknown ensemble {ensemble!r}")
    canonical = ensemble == "canonical"
    path = conditions_path(initial_conditions, conditions_increment, n_states)
    axes = system["axes"]
    if canonical:
        check_canonical_conditions(system, initial_conditions, conditions_increment, occupation, N)
    completed = CompletedRuns(output_params or RunDataOutputParams())
    n_done = min(completed.read(), len(path))
    if summary_dir is None:
        summary_dir = completed.params.output_dir
    species = system.get("species") or [str(i) for i in range(system["n_species"])]
    param_names = ["abcdefghijklmnopqrstuvwxyz"[q] for q in range(len(axes["Rt"]))]
    summary = SummaryWriter(summary_dir, species, param_names) if summary_dir else None
    T_super = np.diag([int(x) for x in (N if not np.isscalar(N) else (N, N, N))])

    def record(res: Dict, cond: Dict, occ_initial, series, n_attempt, seconds) -> None:
        jc = {k: np.asarray(v).tolist() for k, v in cond.items()}
        completed.append({"initial_state": state_to_json(occ_initial, T_super, jc),
                          "final_state": state_to_json(res["final_occupation"], T_super, jc),
                          "conditions": jc, "transformation_matrix_to_supercell": T_super.tolist(),
                          "n_unitcells": int(np.prod(np.diag(T_super)))})
        completed.write()
        if summary is not None:
            analysis = {k: res[k] for k in ("heat_capacity", "mol_susc", "param_susc", "mol_thermochem_susc",
                                            "param_thermochem_susc") if k in res}
            summary.append(jc, analysis, series, int(n_samples), res["acceptance_rate"], int(n_attempt), seconds)

    max_occ = tables.host.max_occ
    o2s = np.full((len(system["occ_to_species"]), max_occ), -1, dtype=np.int32)
    for b, row in enumerate(system["occ_to_species"]):
        o2s[b, :len(row)] = row

    def make_state(conds: List[Dict]):
        st = _capi.State(tables, N, len(conds))
        st.set_eci(eci_index, eci_value)
        st.set_occupants(system["sublat_to_asym"], o2s, system["n_species"])
        sm = _capi.Sampler(st, n_samples, axes["origin"], axes["Rt"], with_corr=with_corr)
        for r, c in enumerate(conds):
            if canonical:       # the potential is the formation energy: no exchange term, no mu . x
                st.set_conditions(float(c["temperature"]), None, r)
                continue
            mu = np.atleast_1d(c["param_chem_pot"])
            st.set_conditions(float(c["temperature"]),
                              semigrand_exchange_table(system["occ_to_species"], axes["Rt"], mu, system["n_species"]), r)
            sm.set_param_chem_pot(mu, r)
        return st, sm

    def equilibrate(st, n_passes, run_seed) -> None:
        if canonical:
            st.canonical_set_swaps(st.canonical_default_swaps())
            if n_passes > 0:
                st.canonical_sweep(int(n_passes), seed=run_seed)
        else:
            st.sgc_sweep(int(n_passes), seed=run_seed, counters=False)

    def collect(st, sm, r, cond, counters, occ_initial=None, seconds=0.0) -> Dict:
        ser = sm.series(r)
        out = _collect(st, sm, r, cond, counters, ser)
        record(out, cond, occupation if occ_initial is None else occ_initial, ser, counters[r].n_attempt, seconds)
        return out

    def _collect(st, sm, r, cond, counters, ser) -> Dict:
        out = {"conditions": {k: np.asarray(v).tolist() for k, v in cond.items()},
               "acceptance_rate": counters[r].n_accept / max(1, counters[r].n_attempt),
               "final_occupation": st.download_occ(r, dtype=np.int8)}
        for name, v in ser.items():
            out[name] = {"mean": np.mean(v, axis=0).tolist(), "n_samples": int(v.shape[0])}
        out.update({k: np.asarray(v).tolist() for k, v in sm.analysis(r).items()})
        return out

    results = []
    todo = path[n_done:]
    if not todo:
        return results
    if not dependent_runs:
        t0 = time.perf_counter()
        st, sm = make_state(todo)
        for r in range(len(todo)):
            st.upload_occ(occupation, r)
        equilibrate(st, n_equilibration_passes, seed)
        cnt = sm.run(int(n_samples), int(sample_period), seed=seed, first_sweep=int(n_equilibration_passes),
                     ensemble=ensemble)
        st.synchronize()
        dt = (time.perf_counter() - t0) / len(todo)
        results = [collect(st, sm, r, todo[r], cnt, seconds=dt) for r in range(len(todo))]
        sm.close()
        st.close()
        return results
    occ = np.asarray(occupation)
    if n_done:
        occ = completed.last_final_occupation()
        if occ is None:
            raise ValueError("run_series: when dependent_runs == true the final state of the last completed run "
                             "must have been saved (save_last_final_state / write_final_states)")
    for k, cond in enumerate(todo, start=n_done):
        t0 = time.perf_counter()
        st, sm = make_state([cond])
        st.upload_occ(occ)
        equilibrate(st, n_equilibration_passes, seed + k)
        cnt = sm.run(int(n_samples), int(sample_period), seed=seed + k, first_sweep=int(n_equilibration_passes),
                     ensemble=ensemble)
        st.synchronize()
        res = collect(st, sm, 0, cond, cnt, occ_initial=occ, seconds=time.perf_counter() - t0)
        occ = res["final_occupation"]
        results.append(res)
        sm.close()
        st.close()
    return results
