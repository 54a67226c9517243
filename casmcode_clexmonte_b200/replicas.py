"""Independent replicas / trajectories across the GPUs of one box (SURVEY.md section 8e row 1,
BASELINE configs[1] and [4]).

One process per GPU.  The reference runs a (mu, T) grid or a batch of KMC trajectories as
independent runs of ``run_series`` (include/casm/clexmonte/run/functions.hh:83-166), one
after the other on one core.  Here they are the replicas of device states: replica `i` of
the global list goes to rank ``i % world`` (round robin), every rank sweeps and samples its
share without talking to anybody, and the ONLY collective is one NCCL all-reduce of the
sampled statistics at the end -- a few hundred bytes per replica:

  * Metropolis grids (``ReplicaRunner``): per replica the additive moments
    ``{n, sum q, sum q q^T}`` of q = (formation energy, potential energy, mol composition,
    param composition) per unit cell, accumulated on the device (``cmx_sampler_moments``)
    straight into the tensor that is reduced; heat capacity and the susceptibilities
    (monte_calculator/analysis_functions.cc:43-173; covariances per
    run/io/convariance_functions.cc:29-143) follow from the reduced sums on every rank.
  * KMC trajectories (``KmcEnsembleRunner``): per trajectory (steps, time), reduced the same
    way.

The dealing, the packing and the analysis are plain functions (no CUDA): they are tested with
the gloo backend, world size 2, on CPU (tests/test_replicas_cpu.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _capi
from .potential import KB, semigrand_exchange_table


# ---------------------------------------------------------------------------
# host logic (no CUDA)
# ---------------------------------------------------------------------------
def deal(n_items: int, world: int, rank: int) -> List[int]:
    """Global indices of the items of `rank`: round robin."""
    return list(range(rank, n_items, world))


def n_moments(n_quantities: int) -> int:
    return 1 + n_quantities + n_quantities * n_quantities


def moments_from_series(q: np.ndarray) -> np.ndarray:
    """{n, sum q, sum q q^T} of a series q[n_samples][Q] -- what cmx_sampler_moments
    computes on the device, restated for the CPU tests."""
    q = np.asarray(q, dtype=np.float64)
    return np.concatenate([[float(q.shape[0])], q.sum(axis=0), (q.T @ q).reshape(-1)])


def scatter_local(local: np.ndarray, ids: Sequence[int], n_global: int) -> np.ndarray:
    """The rank's rows at their global positions, zeros elsewhere: summing these over the
    ranks (the all-reduce) assembles the table of all replicas on every rank."""
    out = np.zeros((n_global, local.shape[1]), dtype=np.float64)
    out[list(ids)] = local
    return out


def allreduce_rows(local, ids: Sequence[int], n_global: int, dist=None, device=None):
    """One all-reduce (sum) of the rows every rank holds.  `local`: numpy [len(ids)][M] or a
    torch tensor already placed with `scatter_local` semantics ([n_global][M]).  Returns
    numpy [n_global][M].  dist None (single process): no collective."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        if isinstance(local, np.ndarray) and local.shape[0] != n_global:
            return scatter_local(local, ids, n_global)
        return np.asarray(local.cpu() if hasattr(local, "cpu") else local)
    import torch
    if isinstance(local, np.ndarray):
        t = torch.from_numpy(scatter_local(local, ids, n_global))
        if device is not None:
            t = t.to(device)
    else:
        t = local
    dist.all_reduce(t)
    return t.cpu().numpy()


def analysis_from_moments(m: np.ndarray, temperature: float, n_unitcells: int, n_species: int, n_param: int) -> Dict:
    """The reference's analysis functions from the additive moments of one replica.
    q = (formation_energy, potential_energy, mol_composition[S], param_composition[P]) per
    unit cell; cov(a, b) = <ab> - <a><b> (population covariance, as Sampler.analysis)."""
    Q = 2 + n_species + n_param
    n = m[0]
    mean = m[1:1 + Q] / n
    cov = m[1 + Q:].reshape(Q, Q) / n - np.outer(mean, mean)
    s0, p0 = 2, 2 + n_species
    c_heat = KB * temperature * temperature / n_unitcells
    c_susc = KB * temperature / n_unitcells
    return {"n_samples": int(round(n)),
            "clex.formation_energy": float(mean[0]), "potential_energy": float(mean[1]),
            "mol_composition": mean[s0:p0].copy(), "param_composition": mean[p0:].copy(),
            "heat_capacity": float(cov[1, 1] / c_heat),
            "mol_susc": cov[s0:p0, s0:p0] / c_susc, "param_susc": cov[p0:, p0:] / c_susc,
            "mol_thermochem_susc": cov[1, s0:p0] / c_susc, "param_thermochem_susc": cov[1, p0:] / c_susc}


# ---------------------------------------------------------------------------
# semi-grand canonical / canonical replica grids
# ---------------------------------------------------------------------------
class ReplicaRunner:
    """A grid of (temperature, param_chem_pot) replicas of one supercell dealt over the ranks.

    conditions: list of {"temperature": T, "param_chem_pot": [...]} for ALL replicas; the
    runner keeps ``deal(len(conditions), world, rank)`` of them as the replicas of one device
    state (its own stream, its own sampler)."""

    def __init__(self, tables: _capi.Tables, N, system: Dict, eci: Dict, conditions: Sequence[Dict], rank: int = 0,
                 world: int = 1, n_samples: int = 64, seed_init: int = 1, init_occ: Optional[np.ndarray] = None):
        self.rank, self.world = int(rank), int(world)
        self.conditions = list(conditions)
        self.ids = deal(len(self.conditions), world, rank)
        self.system = system
        if not self.ids:
            raise _capi.CmxError(_capi.CMX_ERR_INVALID, "more ranks than replicas")
        axes = system["axes"]
        max_occ = tables.host.max_occ
        o2s = np.full((len(system["occ_to_species"]), max_occ), -1, dtype=np.int32)
        for b, row in enumerate(system["occ_to_species"]):
            o2s[b, :len(row)] = row
        self.state = _capi.State(tables, N, len(self.ids))
        self.state.set_eci(eci["index"], eci["value"])
        self.state.set_occupants(system["sublat_to_asym"], o2s, system["n_species"])
        self.sampler = _capi.Sampler(self.state, n_samples, axes["origin"], axes["Rt"])
        for r, i in enumerate(self.ids):
            c = self.conditions[i]
            mu = np.atleast_1d(np.asarray(c["param_chem_pot"], dtype=np.float64))
            self.state.set_conditions(float(c["temperature"]),
                                      semigrand_exchange_table(system["occ_to_species"], axes["Rt"], mu,
                                                               system["n_species"]), r)
            self.sampler.set_param_chem_pot(mu, r)
        if init_occ is None:
            # replica i starts from the same i.i.d. configuration whatever the decomposition:
            # the state's generator is keyed by the LOCAL replica index, so draw on the host
            rng_n = self.state.n_sites
            for r, i in enumerate(self.ids):
                occ = np.random.default_rng([seed_init, i]).integers(0, int(tables.host.n_occ[0]), rng_n).astype(np.int8)
                self.state.upload_occ(occ, r)
        else:
            for r in range(len(self.ids)):
                self.state.upload_occ(init_occ, r)
        self.n_species, self.n_param = self.sampler.n_species, self.sampler.n_param

    def close(self):
        self.sampler.close()
        self.state.close()

    def run(self, n_equilibration: int, n_samples: int, sample_period: int, seed_of=lambda i: 1000 + i):
        """Equilibrate, then sample (n_samples samples, sample_period passes apart).  The
        counter-based generator is keyed by (seed, LOCAL replica index, sweep, site); every
        rank uses its own seed, so the replicas of the grid draw from independent streams.
        The sampled statistics therefore agree between decompositions within their
        statistical error, not bit for bit."""
        seed = seed_of(self.rank)
        self.state.sgc_sweep(int(n_equilibration), seed=seed, counters=False)
        self.sampler.reset()
        return self.sampler.run(int(n_samples), int(sample_period), seed=seed, first_sweep=int(n_equilibration))

    def reduce(self, dist=None, device=None) -> List[Dict]:
        """ONE all-reduce of the moments of all replicas; returns the analysis of every
        replica of the global list (the same on every rank)."""
        Q = self.sampler.n_scalar_quantities
        M = n_moments(Q)
        n_global = len(self.conditions)
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1 and device is not None:
            import torch
            # the device kernel writes the local rows into a scratch tensor on the state's
            # stream; they are scattered to their global rows of the tensor NCCL reduces
            stream = torch.cuda.ExternalStream(self.state.stream())
            with torch.cuda.stream(stream):
                loc = torch.zeros((len(self.ids), M), dtype=torch.float64, device=device)
                self.sampler.moments(device_ptr=loc.data_ptr())
                glob = torch.zeros((n_global, M), dtype=torch.float64, device=device)
                glob[torch.as_tensor(self.ids, device=device)] = loc
                dist.all_reduce(glob)
            stream.synchronize()
            table = glob.cpu().numpy()
        else:
            table = allreduce_rows(self.sampler.moments(), self.ids, n_global, dist)
        n_cells = self.state.n_cells
        return [analysis_from_moments(table[i], float(self.conditions[i]["temperature"]), n_cells, self.n_species,
                                      self.n_param) for i in range(n_global)]


# ---------------------------------------------------------------------------
# KMC trajectories
# ---------------------------------------------------------------------------
class KmcEnsembleRunner:
    """n_trajectories independent rejection-free KMC trajectories dealt over the ranks
    (BASELINE configs[4]); trajectory i is seeded with seed0 + i whatever the decomposition,
    so the ensemble is IDENTICAL for every number of ranks."""

    def __init__(self, tables: _capi.Tables, N, eci: Dict, kmc_factory, temperature: float,
                 n_trajectories: int, rank: int = 0, world: int = 1, seed0: int = 1, occ_of=None):
        """kmc_factory(state) -> _capi.Kmc (event types, local clexulators, prim events);
        occ_of(i) -> initial occupation of trajectory i of the global list."""
        self.rank, self.world = int(rank), int(world)
        self.n_global = int(n_trajectories)
        self.ids = deal(self.n_global, world, rank)
        self.state = _capi.State(tables, N, len(self.ids))
        self.state.set_eci(eci["index"], eci["value"])
        for r, i in enumerate(self.ids):
            self.state.upload_occ(occ_of(i), r)
            self.state.set_conditions(float(temperature), None, r)
        self.kmc = kmc_factory(self.state)
        self.seeds = np.array([seed0 + i for i in self.ids], dtype=np.uint64)
        self.kmc.run_begin(self.seeds)

    def close(self):
        self.kmc.close()
        self.state.close()

    def run(self, n_steps: int) -> Dict:
        return self.kmc.run(int(n_steps))

    def reduce(self, res: Dict, dist=None, device=None) -> np.ndarray:
        """[n_trajectories][3] = (steps, time, total rate) of every trajectory on every rank."""
        local = np.stack([res["n_steps"].astype(np.float64), res["time"], res["total_rate"]], axis=1)
        return allreduce_rows(local, self.ids, self.n_global, dist, device)
