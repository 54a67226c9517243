"""Host-side mirror of the reference's KMC event bookkeeping that feeds the device
event-state kernel (csrc/cmx_kmc.cu).  Names follow the reference:

  PrimEventData / make_prim_event_list   src/casm/clexmonte/events/event_methods.cc:80-133
  EventState                             include/casm/clexmonte/events/event_data.hh:19-30
  EventStateCalculator                   src/casm/clexmonte/monte_calculator/BaseMonteEventData.cc:27-156

Only what the rate recompute needs is here; event lists, impact tables and the
selector stay with the reference's host code.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from .clexulator_tables import read_eci


def make_prim_event_list(event_type_data: Sequence[dict]) -> List[dict]:
    """Linear list of the events associated with the origin unit cell.

    `event_type_data[y]["events"]` lists the equivalent events of type y, each a
    dict {sites: [(b,i,j,k)...], occ_init: [...], occ_final: [...]}, in the order
    of the local basis set's equivalents.  As in append_to_prim_event_list
    (event_methods.cc:80-120) every equivalent contributes its forward event and,
    when it differs, the reverse event right after it; both share the
    equivalent_index (the same local clexulator describes the hop in both
    directions)."""
    out = []
    for y, et in enumerate(event_type_data):
        for eq, ev in enumerate(et["events"]):
            fwd = dict(event_type=y, event_type_name=et.get("name", str(y)), equivalent_index=eq,
                       is_forward=True, prim_event_index=len(out), sites=[tuple(s) for s in ev["sites"]],
                       occ_init=list(ev["occ_init"]), occ_final=list(ev["occ_final"]))
            out.append(fwd)
            if list(ev["occ_init"]) != list(ev["occ_final"]):
                rev = dict(fwd, is_forward=False, prim_event_index=len(out), occ_init=list(ev["occ_final"]),
                           occ_final=list(ev["occ_init"]))
                out.append(rev)
    return out


def read_event_type(event_json, equivalents_info_json, kra_eci, freq_eci, name: str = "") -> dict:
    """Event type data from the reference's input files (system.json "kmc_events"
    + "local_basis_sets", parsed at system/io/json/System_json_io.cc:88-250,546-602):
    the prototype event's occupation (event.json "occupation"/"index") placed on
    the phenomenal cluster of every equivalent (equivalents_info.json
    "equivalents"[k]/"phenomenal"/"sites", listed in the site order of the
    transformed event), plus the kra / freq coefficients (dense or sparse form)."""
    def load(x):
        return json.loads(Path(x).read_text()) if isinstance(x, (str, Path)) else x
    ev = load(event_json)
    info = load(equivalents_info_json)
    occ_init = [int(x) for x in ev["occupation"]["index"]["initial"]]
    occ_final = [int(x) for x in ev["occupation"]["index"]["final"]]
    events = []
    for eq in info["equivalents"]:
        sites = [tuple(int(c) for c in s) for s in eq["phenomenal"]["sites"]]
        if len(sites) != len(occ_init):
            raise ValueError("event and local basis set have different phenomenal clusters")
        events.append(dict(sites=sites, occ_init=occ_init, occ_final=occ_final))
    ki, kv = read_eci(load(kra_eci))
    fi, fv = read_eci(load(freq_eci))
    return dict(name=name, events=events, kra=(ki, kv), freq=(fi, fv))


def event_linear_site_index(N: Sequence[int], unitcell_index: int, sites: Sequence[Sequence[int]]) -> List[int]:
    """set_event_linear_site_index (event_methods.cc:141-155) without a
    SuperNeighborList: l = b*n_cells + cell of (unit cell + site offset), periodic."""
    N0, N1, N2 = (int(x) for x in N)
    n_cells = N0 * N1 * N2
    i, j, k = unitcell_index % N0, (unitcell_index // N0) % N1, unitcell_index // (N0 * N1)
    return [int(b * n_cells + ((i + di) % N0) + N0 * (((j + dj) % N1) + N1 * ((k + dk) % N2)))
            for b, di, dj, dk in sites]


def complete_event_list(n_cells: int, n_prim_events: int):
    """(unitcell_index, prim_event_index) of every event, unit cell major -- the
    order of the CompleteEventList (events/CompleteEventList.cc:10-91)."""
    uc = np.repeat(np.arange(n_cells, dtype=np.int64), n_prim_events)
    pe = np.tile(np.arange(n_prim_events, dtype=np.int32), n_cells)
    return uc, pe


# ---------------------------------------------------------------------------
# impact table (which event rates must be recomputed after an event happened)
# ---------------------------------------------------------------------------
def _function_sites(t, gbeg, c) -> set:
    """Neighbor-list indices read by function `c` of a group table (global_gbeg or
    delta_gbeg row) of exported Clexulator tables."""
    used = set()
    for g in range(int(gbeg[c]), int(gbeg[c + 1])):
        for el in range(int(t.group_ebeg[g]), int(t.group_ebeg[g + 1])):
            for tm in range(int(t.elem_tbeg[el]), int(t.elem_tbeg[el + 1])):
                for f in range(int(t.term_fbeg[tm]), int(t.term_fbeg[tm + 1])):
                    used.add(int(t.factor_n[f]))
    return used


def required_update_neighborhood(formation_tables, eci_index, local_tables, coef_index, sites) -> set:
    """Sites (b, i, j, k), relative to the event's unit cell, whose occupation the rate
    of a prim event depends on -- EventImpactInfo::required_update_neighborhood
    (src/casm/clexmonte/events/event_methods.cc:157-246 builds it from the
    formation-energy cluster-expansion neighborhood of the event's sites and the
    local-cluster-expansion neighborhoods of kra and freq).  Here it is read off the
    exported tables: the event's own sites (event_is_allowed), for each of them the
    neighbors its selected delta functions read (dE_final), and the neighbors the
    selected functions of the equivalent's local clexulator read (E_kra, freq)."""
    ft = formation_tables
    nbr = np.asarray(ft.nbr).reshape(-1, 4)
    nlist_sublat = [int(x) for x in np.asarray(ft.nlist_sublat)]
    out = set()
    for b, di, dj, dk in sites:
        out.add((int(b), int(di), int(dj), int(dk)))
        if int(b) not in nlist_sublat:
            continue
        p = nlist_sublat.index(int(b))
        for c in eci_index:
            for n in _function_sites(ft, ft.delta_gbeg, p * int(ft.corr_size) + int(c)):
                o = nbr[n]
                out.add((int(o[3]), int(di + o[0]), int(dj + o[1]), int(dk + o[2])))
    lt = local_tables
    lnbr = np.asarray(lt.nbr).reshape(-1, 4)
    for c in coef_index:
        for n in _function_sites(lt, lt.global_gbeg, int(c)):
            o = lnbr[n]
            out.add((int(o[3]), int(o[0]), int(o[1]), int(o[2])))
    return out


def make_relative_impact_table(prim_events: Sequence[dict], neighborhoods: Sequence[set]):
    """make_relative_impact_table (src/casm/clexmonte/events/ImpactTable.cc:150-185):
    table[j] = events (i, translation) whose rate must be recomputed after prim event j
    happened in the origin unit cell: a site event j changes (its phenomenal sites) lies
    in the required update neighborhood of event i translated by `translation`.
    Returns (beg[n_prim + 1], entries[n][4] = (i, dx, dy, dz))."""
    beg, entries = [0], []
    for ev_j in prim_events:
        rows = set()
        for i, nb_i in enumerate(neighborhoods):
            for (b, x, y, z) in nb_i:
                for (pb, px, py, pz) in ev_j["sites"]:
                    if int(pb) == b:
                        # event i at cell (phenomenal site - neighborhood offset) reads that site
                        rows.add((i, int(px) - x, int(py) - y, int(pz) - z))
        # cell-major order (dz, dy, dx, prim event): the event ids of a list then ascend except
        # across periodic wraps, which lets the device skip repeated tree parents cheaply
        entries.extend(sorted(rows, key=lambda e: (e[3], e[2], e[1], e[0])))
        beg.append(len(entries))
    return np.asarray(beg, dtype=np.int32), np.asarray(entries, dtype=np.int32).reshape(-1, 4)


# ---------------------------------------------------------------------------
# calculator parameters (accepted as the reference accepts them)
# ---------------------------------------------------------------------------
KINETIC_PARAMS = {
    # key: (default, allowed values or None)
    "verbosity": ("standard", None),
    "print_event_data_summary": (False, (True, False)),
    "mol_composition_tol": (1e-10, None),
    "event_data_type": ("default", ("high_memory", "default", "low_memory")),
    "event_selector_type": ("vector_sum_tree", ("vector_sum_tree", "sum_tree", "direct_sum")),
    "abnormal_event_handling": (None, None),
    "impact_table_type": ("neighborlist", ("neighborlist", "relative")),
    "assign_allowed_events_only": (True, (True, False)),
    "selected_event_data": (None, None),
}


def parse_kinetic_params(params: Optional[dict]) -> dict:
    """The "params" of the kinetic calculator as KineticCalculator::_reset reads them
    (src/casm/clexmonte/monte_calculator/KineticCalculator.cc:78-100 lists the keys,
    :707-830 the values): unknown keys and invalid values are errors with the reference's
    wording; the accepted values are returned with defaults filled in.

    On the device these choose nothing: the selector is always the block-local sum tree over
    the COMPLETE event list with the relative impact table (cmx_kmc.cu), whose selections equal
    the reference's for every event_data_type / event_selector_type / impact_table_type --
    those only trade memory for speed on the host and do not change which event a random
    number picks.  ``device_equivalent`` in the result names what runs."""
    params = dict(params or {})
    unknown = sorted(set(params) - set(KINETIC_PARAMS))
    if unknown:
        raise ValueError(f"Error in KineticCalculator: unrecognized params: {unknown}")
    out = {}
    for key, (default, allowed) in KINETIC_PARAMS.items():
        v = params.get(key, default)
        if allowed is not None and v not in allowed:
            raise ValueError(f"Invalid {key}: {v!r} (allowed: {list(allowed)})")
        out[key] = v
    out["device_equivalent"] = dict(event_data_type="high_memory", event_selector_type="vector_sum_tree",
                                    impact_table_type="relative")
    return out


def allowed_event_list(is_allowed: np.ndarray, n_prim_events: int):
    """The allowed-event list of AllowedEventList (src/casm/clexmonte/events/AllowedEventList.cc:70-98):
    the (unitcell_index, prim_event_index) of every event whose is_allowed flag is set, in
    complete-list order.  `is_allowed`: flags of the complete event list (unit cell major),
    e.g. ``Kmc.event_states(*complete_event_list(...))["is_allowed"]``."""
    idx = np.flatnonzero(np.asarray(is_allowed).reshape(-1))
    return idx // int(n_prim_events), (idx % int(n_prim_events)).astype(np.int32)
