"""Run-series outputs in the reference's file layout.

  <output_dir>/summary.json          one column per completed run, appended run by run, the
                                     layout the reference's results writer produces
                                     ([EXT] libcasm-monte jsonResultsIO) and its tests check
                                     (python/tests/conftest.py:141-259 validate_summary_file /
                                     validate_statistics_data): "conditions", "analysis",
                                     "statistics", "completion_check_results".
  <output_dir>/completed_runs.json   list of RunData
                                     (include/casm/clexmonte/run/io/json/RunData_json_io.hh:13-25):
                                     "conditions", "transformation_matrix_to_supercell",
                                     "n_unitcells" and, when asked, "initial_state" /
                                     "final_state" = {"configuration", "conditions", "properties"}
                                     (state/io/json/State_json_io.cc:13-20); read back on restart
                                     (run/IncrementalConditionsStateGenerator.hh:154-196).

Scalars are {"shape": [], "value": [...]}, vectors {"shape": [n], "component_names": [...],
"<name>": [...]}, matrices {"shape": [n, n], "component_names": ["a,a", "b,a", ...], ...}
(column-major unrolling, as monte::default_component_names does).  Host code only.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np


@dataclass
class RunDataOutputParams:
    """include/casm/clexmonte/run/RunData.hh:21-40 (same names, same defaults as
    run/io/json/RunData_json_io.hh:75-95)."""
    do_save_all_initial_states: bool = False
    do_save_all_final_states: bool = False
    do_save_last_final_state: bool = True
    write_initial_states: bool = False
    write_final_states: bool = False
    output_dir: Optional[str] = None

    @classmethod
    def from_json(cls, d: Dict) -> "RunDataOutputParams":
        return cls(bool(d.get("save_all_initial_states", False)), bool(d.get("save_all_final_states", False)),
                   bool(d.get("save_last_final_state", True)), bool(d.get("write_initial_states", False)),
                   bool(d.get("write_final_states", False)), d.get("output_dir") or None)


def _jsonable(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.floating, np.integer)):
        return v.item()
    if isinstance(v, dict):
        return {k: _jsonable(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    return v


def _write_atomic(path: Path, data) -> None:
    path.parent.mkdir(parents=True, exist_ok=True)
    tmp = path.with_suffix(path.suffix + ".tmp")
    tmp.write_text(json.dumps(_jsonable(data)))
    os.replace(tmp, path)            # (the reference writes through SafeOfstream: tmp file, then rename)


def state_to_json(occupation, transformation_matrix, conditions: Dict, properties: Optional[Dict] = None) -> Dict:
    return {"configuration": {"dof": {"occ": np.asarray(occupation).astype(int).tolist()},
                              "transformation_matrix_to_supercell": np.asarray(transformation_matrix).astype(int).tolist()},
            "conditions": _jsonable(conditions), "properties": _jsonable(properties or {})}


class CompletedRuns:
    """The m_completed_runs bookkeeping of IncrementalConditionsStateGenerator
    (run/IncrementalConditionsStateGenerator.hh:134-196)."""

    def __init__(self, params: RunDataOutputParams):
        self.params = params
        self.runs: List[Dict] = []

    @property
    def path(self) -> Optional[Path]:
        return Path(self.params.output_dir) / "completed_runs.json" if self.params.output_dir else None

    def read(self) -> int:
        self.runs = []
        if self.path is not None and self.path.exists():
            data = json.loads(self.path.read_text())
            if not isinstance(data, list):
                raise ValueError(f"{self.path}: expected a list of runs")
            for k, r in enumerate(data):
                for key in ("conditions", "transformation_matrix_to_supercell", "n_unitcells"):
                    if key not in r:
                        raise ValueError(f"{self.path}: run {k} misses {key!r}")
            self.runs = data
        return len(self.runs)

    def append(self, run: Dict) -> None:
        """run: {"initial_state", "final_state", "conditions", "transformation_matrix_to_supercell",
        "n_unitcells"}; states are dropped per the save flags exactly as the reference does."""
        p = self.params
        if self.runs and not p.do_save_all_final_states:
            self.runs[-1].pop("final_state", None)
        run = dict(run)
        if not p.do_save_all_initial_states:
            run.pop("initial_state", None)
        if not p.do_save_last_final_state and not p.do_save_all_final_states:
            run.pop("final_state", None)
        self.runs.append(run)

    def write(self) -> None:
        if self.path is None:
            return
        p = self.params
        out = []
        for r in self.runs:
            o = {k: r[k] for k in ("conditions", "transformation_matrix_to_supercell", "n_unitcells")}
            if p.write_initial_states and "initial_state" in r:
                o["initial_state"] = r["initial_state"]
            if p.write_final_states and "final_state" in r:
                o["final_state"] = r["final_state"]
            out.append(o)
        _write_atomic(self.path, out)

    def last_final_occupation(self) -> Optional[np.ndarray]:
        if self.runs and "final_state" in self.runs[-1]:
            return np.array(self.runs[-1]["final_state"]["configuration"]["dof"]["occ"], dtype=np.int32)
        return None


# ---------------------------------------------------------------------------
# summary.json
# ---------------------------------------------------------------------------
def matrix_component_names(names: Sequence[str]) -> List[str]:
    return [f"{names[i]},{names[j]}" for j in range(len(names)) for i in range(len(names))]  # column major


def calculated_precision(x: np.ndarray, confidence: float = 0.95) -> float:
    """Precision of the mean of a correlated series: z sqrt(var (1 + rho) / (1 - rho) / N) with
    the lag-1 autocorrelation rho (the estimator family of monte::BasicStatistics [EXT];
    constant series -> 0)."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    if n < 2:
        return 0.0
    d = x - x.mean()
    var = float(d @ d) / n
    if var <= 0.0:
        return 0.0
    rho = float(d[:-1] @ d[1:]) / n / var
    rho = min(max(rho, 0.0), 0.999)
    z = math.sqrt(2.0) * _erfinv(confidence)
    return z * math.sqrt(var * (1.0 + rho) / (1.0 - rho) / n)


def _erfinv(y: float) -> float:
    lo, hi = 0.0, 6.0
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if math.erf(mid) < y:
            lo = mid
        else:
            hi = mid
    return 0.5 * (lo + hi)


class SummaryWriter:
    """Appends one run to <output_dir>/summary.json (reading what is there first, as the
    reference's writer does, so a restarted series continues the same file)."""

    def __init__(self, output_dir, species: Sequence[str], param_names: Sequence[str]):
        self.path = Path(output_dir) / "summary.json"
        self.species = list(species)
        self.params = list(param_names)
        self.data = json.loads(self.path.read_text()) if self.path.exists() else {}

    def _names(self, key: str, n: int) -> List[str]:
        if key.startswith("mol_") and "susc" not in key and n == len(self.species):
            return self.species
        if key.startswith("param_") and "susc" not in key and n == len(self.params):
            return self.params
        if key in ("mol_thermochem_susc",) and n == len(self.species):
            return [f"S,{s}" for s in self.species]
        if key in ("param_thermochem_susc",) and n == len(self.params):
            return [f"S,{s}" for s in self.params]
        if key == "mol_susc" and n == len(self.species) ** 2:
            return matrix_component_names(self.species)
        if key == "param_susc" and n == len(self.params) ** 2:
            return matrix_component_names(self.params)
        return [str(i) for i in range(n)]

    def _push(self, section: str, key: str, value, stat: bool = False) -> None:
        sec = self.data.setdefault(section, {})
        v = np.asarray(value, dtype=np.float64) if not stat else None
        if stat:
            series = np.asarray(value, dtype=np.float64)
            cols = series.reshape(series.shape[0], -1)
            shape = list(series.shape[1:])
            entry = sec.setdefault(key, {"shape": shape})
            if not shape:
                tgt = entry.setdefault("value", {"mean": [], "calculated_precision": []})
                tgt["mean"].append(float(cols[:, 0].mean()))
                tgt["calculated_precision"].append(calculated_precision(cols[:, 0]))
            else:
                names = self._names(key, cols.shape[1])
                entry["component_names"] = names
                for q, nm in enumerate(names):
                    tgt = entry.setdefault(nm, {"mean": [], "calculated_precision": []})
                    tgt["mean"].append(float(cols[:, q].mean()))
                    tgt["calculated_precision"].append(calculated_precision(cols[:, q]))
            return
        shape = list(v.shape)
        entry = sec.setdefault(key, {"shape": shape})
        if not shape:
            entry.setdefault("value", []).append(float(v))
        else:
            flat = v.reshape(-1, order="F")
            names = self._names(key, flat.size)
            entry["component_names"] = names
            for q, nm in enumerate(names):
                entry.setdefault(nm, []).append(float(flat[q]))

    def append(self, conditions: Dict, analysis: Dict, series: Dict[str, np.ndarray], n_samples: int,
               acceptance_rate: float, count: int, elapsed_clocktime: float) -> None:
        for k, v in conditions.items():
            self._push("conditions", k, v)
        for k, v in analysis.items():
            self._push("analysis", k, v)
        for k, v in series.items():
            self._push("statistics", k, v, stat=True)
        ccr = self.data.setdefault("completion_check_results", {})
        for k, v in (("N_samples", int(n_samples)), ("N_samples_for_statistics", int(n_samples)),
                     ("acceptance_rate", float(acceptance_rate)), ("count", int(count)),
                     ("elapsed_clocktime", float(elapsed_clocktime))):
            ccr.setdefault(k, []).append(v)
        _write_atomic(self.path, self.data)

    def n_runs(self) -> int:
        ccr = self.data.get("completion_check_results", {})
        return len(ccr.get("count", []))
