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

Results are returned per state with the reference's sampler / analysis function names;
writing summary.json / completed_runs.json stays with the reference's results I/O
([EXT] libcasm-monte jsonResultsIO), which is not present in this repository's toolchain.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _capi
from .potential import semigrand_exchange_table


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


def run_series(tables: "_capi.Tables", N: Sequence[int], system: Dict, eci_index, eci_value,
               initial_conditions: Dict, conditions_increment: Dict, n_states: int, occupation: np.ndarray,
               n_equilibration_passes: int, n_samples: int, sample_period: int = 1, seed: int = 0,
               dependent_runs: bool = False, with_corr: bool = False) -> List[Dict]:
    """Semi-grand canonical run series.  `system`: occ_to_species, sublat_to_asym, n_species,
    axes {origin, Rt}.  Conditions: {"temperature": T, "param_chem_pot": [...]}.
    Returns one dict per state: conditions, means of the sampled quantities, the analysis
    functions, acceptance rate, final occupation."""
    path = conditions_path(initial_conditions, conditions_increment, n_states)
    axes = system["axes"]
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
            mu = np.atleast_1d(c["param_chem_pot"])
            st.set_conditions(float(c["temperature"]),
                              semigrand_exchange_table(system["occ_to_species"], axes["Rt"], mu, system["n_species"]), r)
            sm.set_param_chem_pot(mu, r)
        return st, sm

    def collect(st, sm, r, cond, counters) -> Dict:
        ser = sm.series(r)
        out = {"conditions": {k: np.asarray(v).tolist() for k, v in cond.items()},
               "acceptance_rate": counters[r].n_accept / max(1, counters[r].n_attempt),
               "final_occupation": st.download_occ(r, dtype=np.int8)}
        for name, v in ser.items():
            out[name] = {"mean": np.mean(v, axis=0).tolist(), "n_samples": int(v.shape[0])}
        out.update({k: np.asarray(v).tolist() for k, v in sm.analysis(r).items()})
        return out

    results = []
    if not dependent_runs:
        st, sm = make_state(path)
        for r in range(len(path)):
            st.upload_occ(occupation, r)
        st.sgc_sweep(int(n_equilibration_passes), seed=seed, counters=False)
        cnt = sm.run(int(n_samples), int(sample_period), seed=seed, first_sweep=int(n_equilibration_passes))
        results = [collect(st, sm, r, path[r], cnt) for r in range(len(path))]
        sm.close()
        st.close()
        return results
    occ = np.asarray(occupation)
    for k, cond in enumerate(path):
        st, sm = make_state([cond])
        st.upload_occ(occ)
        st.sgc_sweep(int(n_equilibration_passes), seed=seed + k, counters=False)
        cnt = sm.run(int(n_samples), int(sample_period), seed=seed + k, first_sweep=int(n_equilibration_passes))
        res = collect(st, sm, 0, cond, cnt)
        occ = res["final_occupation"]
        results.append(res)
        sm.close()
        st.close()
    return results
