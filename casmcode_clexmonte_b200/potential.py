"""Host-side pieces of the potentials (rows a7/a8 of SURVEY.md section 8a).

The device evaluates ``dE_potential = dE_clex - exch[b][occ_i][occ_f]``; this
module tabulates ``exch`` exactly the way the reference evaluates the
semi-grand term for a single-site change
(src/casm/clexmonte/monte_calculator/SemiGrandCanonicalCalculator.cc:201-212):

    delta_N[curr_species] += -1;  delta_N[new_species] += +1
    dE_pot = dE_clex - param_chem_pot . (R^T * delta_N)

with ``R^T = CompositionConverter::dparam_dmol()`` (:144).  Scalars only; the
O(N) work stays on the device.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

KB = 8.6173303e-05  # eV/K, CASM::KB (pinned by _MonteCalculator.py:186-199)


def dparam_dmol(origin: Sequence[float], end_members: Sequence[Sequence[float]]) -> np.ndarray:
    """R^T with x = R^T (n - origin): left pseudo-inverse of (end_members - origin).

    [EXT] composition::CompositionConverter (libcasm-composition v2.3.0).
    """
    Q = (np.asarray(end_members, dtype=np.float64) - np.asarray(origin, dtype=np.float64)).T
    return np.linalg.inv(Q.T @ Q) @ Q.T


def semigrand_exchange_table(occ_to_species, Rt, param_chem_pot, n_species: int) -> np.ndarray:
    """exch[b][occ_i][occ_f] = param_chem_pot . (R^T dN) for the change occ_i -> occ_f on sublattice b.

    ``occ_to_species[b][occ]`` is -1 padded.  Operation order follows the
    reference (dense mat-vec over species, then dot over parameters) so that
    the value is bit-identical to the host computation it replaces.
    """
    o2s = np.asarray(occ_to_species, dtype=np.int64)
    Rt = np.asarray(Rt, dtype=np.float64).reshape(-1, n_species)
    mu = np.asarray(param_chem_pot, dtype=np.float64).reshape(-1)
    if Rt.shape[0] != mu.shape[0]:
        raise ValueError("Error in SemiGrandCanonicalPotential: param_chem_pot size error")
    n_sublat, max_occ = o2s.shape
    out = np.zeros((n_sublat, max_occ, max_occ), dtype=np.float64)
    for b in range(n_sublat):
        for oi in range(max_occ):
            for of in range(max_occ):
                si, sf = o2s[b, oi], o2s[b, of]
                if si < 0 or sf < 0:
                    continue
                dN = np.zeros(n_species)
                dN[si] += -1.0
                dN[sf] += 1.0
                dot = 0.0
                for p in range(len(mu)):
                    xp = 0.0
                    for s in range(n_species):
                        xp += Rt[p, s] * dN[s]
                    dot += mu[p] * xp
                out[b, oi, of] = dot
    return out


def mol_composition(counts: np.ndarray, occ_to_species, n_species: int, n_cells: int) -> np.ndarray:
    """CompositionCalculator::mean_num_each_component from device occupant counts[b][occ]
    (sampling_functions.cc:37-54): number of each component per unit cell."""
    o2s = np.asarray(occ_to_species, dtype=np.int64)
    n = np.zeros(n_species, dtype=np.float64)
    for b in range(o2s.shape[0]):
        for o in range(o2s.shape[1]):
            if o2s[b, o] >= 0:
                n[o2s[b, o]] += float(counts[b, o])
    return n / float(n_cells)


def param_composition(mol_comp: np.ndarray, origin, Rt) -> np.ndarray:
    """CompositionConverter::param_composition: x = R^T (n - origin)."""
    return np.asarray(Rt, dtype=np.float64) @ (np.asarray(mol_comp, dtype=np.float64)
                                               - np.asarray(origin, dtype=np.float64))


def semigrand_potential_per_supercell(e_clex: float, mol_comp, origin, Rt, param_chem_pot,
                                      n_cells: int) -> float:
    """SemiGrandCanonicalPotential::per_supercell (SemiGrandCanonicalCalculator.cc:171-179)."""
    Rt = np.asarray(Rt, dtype=np.float64)
    mu = np.asarray(param_chem_pot, dtype=np.float64)
    dot = 0.0
    for p in range(len(mu)):
        xp = 0.0
        for s in range(Rt.shape[1]):
            xp += Rt[p, s] * (mol_comp[s] - origin[s])
        dot += mu[p] * xp
    return e_clex - float(n_cells) * dot


def canonical_swap_types(tables, sublat_to_asym, occ_to_species, N, n_shell: int = 12, long_range: bool = True):
    """Swap types for `State.canonical_set_swaps`, the parallel counterpart of the
    reference's canonical swap table (make_canonical_swaps [EXT], built at
    src/casm/clexmonte/system/System.cc:55-58: every pair of candidates with the
    same species allowed on both asymmetric units).

    For every ordered pair of mutable sublattices on the same asymmetric unit:
    the `n_shell` shortest translations of the prim neighbor list (one of +t/-t
    when the sublattices are equal -- both give the same set of pairs), and, with
    `long_range`, one long translation per pair that moves species across the
    whole box (t_i = 2 mod 4, so a stride-4 colouring along i is conflict free
    for any short-ranged basis)."""
    nbr = np.asarray(tables.nbr).reshape(-1, 4)
    n_occ = np.asarray(tables.n_occ)
    mutable = [b for b in range(len(n_occ)) if n_occ[b] > 1]
    swaps = []
    for ba in mutable:
        for bb in mutable:
            if bb < ba or sublat_to_asym[ba] != sublat_to_asym[bb]:
                continue
            if not set(s for s in occ_to_species[ba] if s >= 0) & set(s for s in occ_to_species[bb] if s >= 0):
                continue
            seen = set()
            count = 0
            for di, dj, dk, b in nbr:
                t = (int(di), int(dj), int(dk))
                if b != bb or (t == (0, 0, 0) and ba == bb):
                    continue
                if ba == bb and (tuple(-x for x in t) in seen):
                    continue
                seen.add(t)
                swaps.append((ba, bb, t))
                count += 1
                if count >= (n_shell // 2 if ba == bb else n_shell):
                    break
            if long_range and N[0] >= 8 and N[0] % 4 == 0:
                t0 = (N[0] // 2) - ((N[0] // 2) % 4) + 2
                swaps.append((ba, bb, (t0 % N[0], N[1] // 3, N[2] // 2 + 1 if N[2] > 2 else 0)))
    return swaps
