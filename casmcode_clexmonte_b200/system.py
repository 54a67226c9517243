"""The reference's `system` input (system.json) -> everything the device path needs.

Mirror of the part of ``parse(InputParser<System>&, ...)`` the hot path consumes
(src/casm/clexmonte/system/io/json/System_json_io.cc:302-452 documents the format,
:462-760 parses it): "prim", "composition_axes", "basis_sets", "clex", "multiclex",
"local_basis_sets", "kmc_events".  Paths inside the file are resolved against the
directory of system.json first and the given search path after it, as the reference's
``resolve_path`` does.  Keys the hot path does not use ("dof_spaces", "dof_subspaces",
"event_system", "local_clex", "n_dimensions", anything prefixed "_") are carried through
untouched under ``extra``.

What comes out (``System``):

  components / n_species   species names in the order of composition_axes["components"]
  occ_to_species           per sublattice: species index of every allowed occupant
  sublat_to_asym           sublattices with the same occupant list share an asymmetric-unit
                           index.  (The reference asks the prim's factor group [EXT
                           libcasm-xtal]; "sublat_to_asym" in the file overrides this rule.)
  axes                     {components, origin, end_members, Rt}: Rt = d(param)/d(mol), the
                           left pseudo-inverse of (end_members - origin)
                           (composition::CompositionConverter [EXT])
  basis_sets[name]         ClexulatorTables exported from the generated Clexulator source
  clex[name]               {"basis_set", "index", "value"}   (coefficients: both ECI formats)
  multiclex[name]          {"basis_set", "coefficients": {key: {"index", "value"}}}
  local_basis_sets[name]   {"tables": [ClexulatorTables per equivalent], "equivalents_info"}
                           equivalents are found next to "source" as <dir>/<i>/<stem>_<i>.cc
  event_types              list for kmc.make_prim_event_list / _capi.Kmc: per "kmc_events"
                           entry the events of every equivalent and the kra / freq coefficients

No CUDA in this module.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import kmc as K
from .clexulator_tables import ClexulatorTables, check_tables_against_basis, parse_clexulator_source, read_eci


class SystemError_(ValueError):
    """A system.json the loader cannot use (message names the key)."""


@dataclass
class System:
    components: List[str]
    occ_to_species: List[List[int]]
    sublat_to_asym: List[int]
    mutable_sublats: List[int]
    axes: Dict
    basis_sets: Dict[str, ClexulatorTables] = field(default_factory=dict)
    clex: Dict[str, Dict] = field(default_factory=dict)
    multiclex: Dict[str, Dict] = field(default_factory=dict)
    local_basis_sets: Dict[str, Dict] = field(default_factory=dict)
    event_types: List[Dict] = field(default_factory=list)
    prim: Dict = field(default_factory=dict)
    extra: Dict = field(default_factory=dict)

    @property
    def n_species(self) -> int:
        return len(self.components)

    def as_dict(self, clex_name: Optional[str] = None) -> Dict:
        """The plain dict the runners of this package take (replicas.ReplicaRunner,
        run_series.run_series, ...)."""
        return dict(species=self.components, n_species=self.n_species, occ_to_species=self.occ_to_species,
                    sublat_to_asym=self.sublat_to_asym, mutable_sublats=self.mutable_sublats, axes=self.axes)


def composition_axes(a: Dict) -> Dict:
    """"composition_axes": components, origin, the end members "a", "b", ... (vectors or
    column matrices), independent_compositions."""
    try:
        comps = [str(x) for x in a["components"]]
        k = int(a["independent_compositions"])
        flat = lambda v: [float(x) for x in np.array(v, dtype=float).reshape(-1)]  # noqa: E731
        origin = flat(a["origin"])
        ends = [flat(a["abcdefghijklmnopqrstuvwxyz"[q]]) for q in range(k)]
    except KeyError as e:
        raise SystemError_(f"composition_axes: missing {e}") from None
    if any(len(v) != len(comps) for v in [origin] + ends):
        raise SystemError_("composition_axes: origin / end members must have one entry per component")
    if k:
        Q = (np.array(ends) - np.array(origin)).T
        Rt = (np.linalg.inv(Q.T @ Q) @ Q.T).tolist()
    else:
        Rt = np.zeros((0, len(comps))).tolist()
    return dict(components=comps, origin=origin, end_members=ends, Rt=Rt)


def _resolve(p, roots: Sequence[Path], what: str) -> Path:
    p = Path(str(p))
    if p.is_absolute() and p.exists():
        return p
    for r in roots:
        if (r / p).exists():
            return r / p
    raise SystemError_(f"{what}: file {str(p)!r} not found (searched {[str(r) for r in roots]})")


def _coefficients(x, roots, corr_size, what) -> Dict:
    data = x if isinstance(x, (list, dict)) and not isinstance(x, str) and ("orbits" in x or isinstance(x, list)) \
        else json.loads(_resolve(x, roots, what).read_text())
    idx, val = read_eci(data, corr_size)
    return dict(index=idx.tolist(), value=val.tolist())


def _local_sources(src: Path) -> List[Path]:
    out = []
    while True:
        p = src.parent / str(len(out)) / f"{src.stem}_{len(out)}{src.suffix}"
        if not p.exists():
            break
        out.append(p)
    return out


def load_system(source, search_path: Sequence = ()) -> System:
    """source: path of a system.json, or the parsed dict (then relative paths resolve against
    `search_path` only)."""
    if isinstance(source, (str, Path)):
        path = Path(source)
        data = json.loads(path.read_text())
        roots = [path.resolve().parent] + [Path(p) for p in search_path]
    else:
        data = dict(source)
        roots = [Path(p) for p in search_path] or [Path.cwd()]
    if "kwargs" in data and "system" in data["kwargs"]:     # a whole run input: {"method", "kwargs": {"system": ...}}
        data = data["kwargs"]["system"]
    for key in ("prim", "composition_axes"):
        if key not in data:
            raise SystemError_(f"system: missing required {key!r}")
    axes = composition_axes(data["composition_axes"])
    comps = axes["components"]
    occ_to_species = []
    for b, site in enumerate(data["prim"]["basis"]):
        row = []
        for name in site["occupants"]:
            if name not in comps:
                raise SystemError_(f"prim basis site {b}: occupant {name!r} is not a composition-axes component")
            row.append(comps.index(name))
        occ_to_species.append(row)
    if "sublat_to_asym" in data:
        s2a = [int(x) for x in data["sublat_to_asym"]]
    else:
        seen: Dict[tuple, int] = {}
        s2a = [seen.setdefault(tuple(r), len(seen)) for r in occ_to_species]
    sysd = System(components=comps, occ_to_species=occ_to_species, sublat_to_asym=s2a,
                  mutable_sublats=[b for b, r in enumerate(occ_to_species) if len(r) > 1], axes=axes,
                  prim=data["prim"])

    for name, bs in (data.get("basis_sets") or {}).items():
        src = _resolve(bs["source"], roots, f"basis_sets/{name}/source")
        t = parse_clexulator_source(src, name=name)
        if t.n_sublat != len(occ_to_species):
            raise SystemError_(f"basis_sets/{name}: clexulator has {t.n_sublat} sublattices, prim has {len(occ_to_species)}")
        for b, r in enumerate(occ_to_species):
            if b in set(int(x) for x in t.nlist_sublat) and int(t.n_occ[b]) != len(r):
                raise SystemError_(f"basis_sets/{name}: sublattice {b} has {int(t.n_occ[b])} occupants in the "
                                   f"clexulator and {len(r)} in the prim")
        # "basis": the project's basis.json (System_json_io.cc: basis_sets/<name>/{source, basis});
        # when it is there the exported tables must agree with it
        if bs.get("basis"):
            try:
                check_tables_against_basis(t, _resolve(bs["basis"], roots, f"basis_sets/{name}/basis"))
            except ValueError as e:
                raise SystemError_(f"basis_sets/{name}: {e}") from None
        sysd.basis_sets[name] = t

    def basis(name, entry, what):
        bs = entry.get("basis_set")
        if bs not in sysd.basis_sets:
            raise SystemError_(f"{what}/{name}: basis_set {bs!r} is not one of {sorted(sysd.basis_sets)}")
        return bs

    for name, c in (data.get("clex") or {}).items():
        bs = basis(name, c, "clex")
        sysd.clex[name] = dict(basis_set=bs, **_coefficients(c["coefficients"], roots, sysd.basis_sets[bs].corr_size,
                                                             f"clex/{name}/coefficients"))
    for name, c in (data.get("multiclex") or {}).items():
        bs = basis(name, c, "multiclex")
        sysd.multiclex[name] = dict(basis_set=bs, coefficients={
            k: _coefficients(v, roots, sysd.basis_sets[bs].corr_size, f"multiclex/{name}/coefficients/{k}")
            for k, v in c["coefficients"].items()})

    for name, lb in (data.get("local_basis_sets") or {}).items():
        src = _resolve(lb["source"], roots, f"local_basis_sets/{name}/source")
        eq_sources = _local_sources(src)
        if not eq_sources:
            raise SystemError_(f"local_basis_sets/{name}: no equivalents found next to {src} "
                               f"(expected {src.parent}/0/{src.stem}_0{src.suffix}, ...)")
        info = json.loads(_resolve(lb["equivalents_info"], roots, f"local_basis_sets/{name}/equivalents_info").read_text())
        if len(info.get("equivalents", [])) != len(eq_sources):
            raise SystemError_(f"local_basis_sets/{name}: {len(eq_sources)} equivalent sources, "
                               f"{len(info.get('equivalents', []))} entries of equivalents_info")
        sysd.local_basis_sets[name] = dict(
            tables=[parse_clexulator_source(p, name=f"{name}_{k}") for k, p in enumerate(eq_sources)],
            equivalents_info=info)

    for name, ev in (data.get("kmc_events") or {}).items():
        lbs = ev.get("local_basis_set")
        if lbs not in sysd.local_basis_sets:
            raise SystemError_(f"kmc_events/{name}: local_basis_set {lbs!r} is not one of {sorted(sysd.local_basis_sets)}")
        coef = ev.get("coefficients") or {}
        for k in ("kra", "freq"):
            if k not in coef:
                raise SystemError_(f"kmc_events/{name}/coefficients: missing {k!r}")
        et = K.read_event_type(_resolve(ev["event"], roots, f"kmc_events/{name}/event"),
                               sysd.local_basis_sets[lbs]["equivalents_info"],
                               json.loads(_resolve(coef["kra"], roots, f"kmc_events/{name}/coefficients/kra").read_text()),
                               json.loads(_resolve(coef["freq"], roots, f"kmc_events/{name}/coefficients/freq").read_text()),
                               name=name)
        et["local_basis_set"] = lbs
        sysd.event_types.append(et)

    used = {"prim", "composition_axes", "basis_sets", "clex", "multiclex", "local_basis_sets", "kmc_events",
            "sublat_to_asym"}
    sysd.extra = {k: v for k, v in data.items() if k not in used}
    return sysd
