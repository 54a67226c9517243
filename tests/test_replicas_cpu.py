"""Replica grids / KMC trajectories over ranks: the host logic --
dealing, packing, ONE all-reduce of additive moments, analysis from the reduced sums --
with the gloo backend, world size 2, on CPU.  (The device side, cmx_sampler_moments, is
checked against moments_from_series in tests/test_gpu_sampler.py.)"""
import os

import numpy as np
import pytest

from casmcode_clexmonte_b200 import replicas as R
from casmcode_clexmonte_b200.potential import KB


def _series(i, n=200, S=3, P=2):
    rng = np.random.default_rng(100 + i)
    e = rng.normal(-0.1 * i, 0.01, n)
    comp = rng.dirichlet(np.ones(S), n)
    x = comp[:, 1:1 + P] + 0.1 * e[:, None]
    return np.column_stack([e, e - 0.3 * x[:, 0], comp, x])


def test_deal_is_a_partition():
    for n, w in ((64, 8), (5, 2), (4096, 8), (3, 4)):
        seen = sorted(i for r in range(w) for i in R.deal(n, w, r))
        assert seen == list(range(n))
        assert max(len(R.deal(n, w, r)) for r in range(w)) - min(len(R.deal(n, w, r)) for r in range(w)) <= 1


def test_moments_are_additive_and_give_the_reference_analysis():
    S, P, T, ncell = 3, 2, 900.0, 4096
    q = _series(0)
    m = R.moments_from_series(q)
    ma, mb = R.moments_from_series(q[:70]), R.moments_from_series(q[70:])
    np.testing.assert_allclose(ma + mb, m, rtol=1e-13)
    a = R.analysis_from_moments(m, T, ncell, S, P)
    e, n, x = q[:, 1], q[:, 2:5], q[:, 5:7]
    assert a["heat_capacity"] == pytest.approx(e.var() * ncell / (KB * T * T), rel=1e-9)
    cov = lambda u, v: ((u - u.mean(0)).T @ (v - v.mean(0))) / len(u)   # noqa: E731
    np.testing.assert_allclose(a["mol_susc"], cov(n, n) * ncell / (KB * T), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(a["param_susc"], cov(x, x) * ncell / (KB * T), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(a["param_thermochem_susc"], cov(e[:, None], x)[0] * ncell / (KB * T), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(a["mol_composition"], n.mean(0), rtol=1e-13)


def _worker(rank, world, port, n_global, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = R.deal(n_global, world, rank)
    local = np.stack([R.moments_from_series(_series(i)) for i in ids])
    table = R.allreduce_rows(local, ids, n_global, dist)
    # KMC style rows (steps, time, rate) reduce the same way
    kmc = R.allreduce_rows(np.array([[10.0, 0.5 * i, 2.0] for i in ids]), ids, n_global, dist)
    out[rank] = (table, kmc)
    dist.destroy_process_group()


def test_two_ranks_reduce_the_grid_with_one_allreduce():
    import torch.multiprocessing as mp
    n_global, world = 7, 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29800 + os.getpid() % 1000
    mp.spawn(_worker, args=(world, port, n_global, out), nprocs=world, join=True)
    want = np.stack([R.moments_from_series(_series(i)) for i in range(n_global)])
    for rank in range(world):
        table, kmc = out[rank]
        np.testing.assert_allclose(table, want, rtol=1e-14)
        np.testing.assert_allclose(kmc[:, 1], 0.5 * np.arange(n_global))
    # every rank can now evaluate every replica
    a = R.analysis_from_moments(out[1][0][3], 800.0, 512, 3, 2)
    assert a["n_samples"] == 200 and np.isfinite(a["heat_capacity"])
