"""ORACLE / TEST INFRASTRUCTURE ONLY.

Plain-Python evaluation of the flat clexulator tables (the wire format produced
by ``casmcode_clexmonte_b200.clexulator_tables``), in the reference's operation
order.  Purpose: validate the *exporter* against ``oracle/_ref`` (the
reference's own generated kernels) on CPU, so that a table bug cannot hide
behind a matching CUDA evaluator.  Pure-Python loops: small cases only.

Follows the generated source's evaluation (e.g.
FCC_binary_vacancy_Clexulator_default.cc:446-498 global, :500-553 point,
:555-610 delta) and [EXT] Correlations::occ_delta / per_supercell
(SURVEY.md Appendix B).
"""
from __future__ import annotations

import numpy as np


def _wrap_cell(N, i, j, k):
    return (i % N[0]) + N[0] * ((j % N[1]) + N[1] * (k % N[2]))


def neighbor_sites(tables, N, cell: int) -> np.ndarray:
    """Linear site indices l = b*n_cells + cell' of the neighbor list of ``cell``."""
    n_cells = N[0] * N[1] * N[2]
    i = cell % N[0]
    j = (cell // N[0]) % N[1]
    k = cell // (N[0] * N[1])
    out = np.zeros(tables.nlist_len, dtype=np.int64)
    for n, (di, dj, dk, b) in enumerate(tables.nbr):
        out[n] = int(b) * n_cells + _wrap_cell(N, i + int(di), j + int(dj), k + int(dk))
    return out


def _eval_function(t, gbeg, gend, sites, occ, occ_i=None, occ_f=None, pb=None):
    total = None
    for g in range(gbeg, gend):
        s = None
        if t.group_has_sum[g]:
            for e in range(t.group_ebeg[g], t.group_ebeg[g + 1]):
                ev = None
                for tm in range(t.elem_tbeg[e], t.elem_tbeg[e + 1]):
                    tv = float(t.term_coef[tm])
                    for q in range(t.term_fbeg[tm], t.term_fbeg[tm + 1]):
                        n = int(t.factor_n[q])
                        b = int(t.nbr[n, 3])
                        tv = tv * float(t.phi[b, int(t.factor_f[q]), int(occ[sites[n]])])
                    ev = tv if ev is None else ev + tv
                s = ev if s is None else s + ev
        if t.group_dphi[g] >= 0:
            f = int(t.group_dphi[g])
            d = float(t.phi[pb, f, occ_f]) - float(t.phi[pb, f, occ_i])
            v = d * s if t.group_has_sum[g] else d
        else:
            v = s
        if t.group_div[g] != 0.0:
            v = v / float(t.group_div[g])
        total = v if total is None else total + v
    return 0.0 if total is None else total


def cell_corr(t, N, occ, cell: int) -> np.ndarray:
    sites = neighbor_sites(t, N, cell)
    return np.array([_eval_function(t, t.global_gbeg[c], t.global_gbeg[c + 1], sites, occ)
                     for c in range(t.corr_size)])


def global_corr(t, N, occ) -> np.ndarray:
    n_cells = N[0] * N[1] * N[2]
    out = np.zeros(t.corr_size)
    for v in range(n_cells):
        out += cell_corr(t, N, occ, v)
    return out


def _point_index(t, n_cells, l):
    b = l // n_cells
    pos = int(np.where(t.nlist_sublat == b)[0][0])
    return pos, b


def point_corr(t, N, occ, l: int) -> np.ndarray:
    n_cells = N[0] * N[1] * N[2]
    p, _ = _point_index(t, n_cells, l)
    sites = neighbor_sites(t, N, l % n_cells)
    return np.array([_eval_function(t, t.point_gbeg[p * t.corr_size + c],
                                    t.point_gbeg[p * t.corr_size + c + 1], sites, occ)
                     for c in range(t.corr_size)])


def delta_corr(t, N, occ, l: int, new_occ: int) -> np.ndarray:
    n_cells = N[0] * N[1] * N[2]
    p, b = _point_index(t, n_cells, l)
    sites = neighbor_sites(t, N, l % n_cells)
    return np.array([_eval_function(t, t.delta_gbeg[p * t.corr_size + c],
                                    t.delta_gbeg[p * t.corr_size + c + 1], sites, occ,
                                    int(occ[l]), int(new_occ), b)
                     for c in range(t.corr_size)])
