"""Multi-GPU parity (-m gpu, needs >= 2 devices): the slab-decomposed sweep with
NCCL halo exchange reproduces the single-GPU trajectory bit for bit (the RNG
counters use global coordinates, so the decomposition must be invisible)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, N, n_sweeps, p2p, out, calls=1, flags=0):
    import torch
    import torch.distributed as dist
    from casmcode_clexmonte_b200 import _capi
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    from casmcode_clexmonte_b200.potential import semigrand_exchange_table
    from casmcode_clexmonte_b200.slab import SlabRunner
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sysd = json.loads((GOLDEN / "systems.json").read_text())["fcc"]
    tables = _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / "fcc_default.npz"), device=rank)
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.1, -0.2], 3)
    Nt = (N, N, N) if np.isscalar(N) else tuple(N)
    init = np.random.default_rng(4).integers(0, 3, int(np.prod(Nt))).astype(np.int8)
    run = SlabRunner(tables, Nt, sysd["eci_sparse"], 900.0, ex, rank, world, rank, init_occ=init, p2p=p2p)
    assert run.p2p == p2p
    run.state.set_sweep_flags(_capi.CMX_SWEEP_DE_SUM | flags)
    run.state.counters_reset()
    for c in range(calls):   # several calls: the layer counters carry over between launches
        run.sweep(n_sweeps, seed=17, first_sweep=c * n_sweeps)
    run.synchronize()
    cnt = run.counters()
    g = run.gather_global()
    if rank == 0:
        out["occ"] = g
        out["cnt"] = cnt
    dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [True, False])
def test_two_slabs_equal_one_gpu(p2p):
    import torch
    import torch.multiprocessing as mp
    from casmcode_clexmonte_b200 import _capi
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    from casmcode_clexmonte_b200.potential import semigrand_exchange_table
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    N, n_sweeps, world = 32, 6, 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + os.getpid() % 1000 + (7 if p2p else 0)
    mp.spawn(_worker, args=(world, port, N, n_sweeps, p2p, out), nprocs=world, join=True)
    sysd = json.loads((GOLDEN / "systems.json").read_text())["fcc"]
    tables = _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / "fcc_default.npz"))
    st = _capi.State(tables, (N, N, N))
    st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
    st.set_conditions(900.0, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.1, -0.2], 3))
    st.upload_occ(np.random.default_rng(4).integers(0, 3, N ** 3).astype(np.int8))
    st.set_sweep_flags(_capi.CMX_SWEEP_DE_SUM)
    cnt = st.sgc_sweep(n_sweeps, seed=17)
    assert (st.download_occ(dtype=np.int8) == out["occ"]).all()
    assert cnt[0].n_accept == int(out["cnt"][1]) and cnt[0].n_attempt == int(out["cnt"][0])
    assert cnt[0].dE_sum == pytest.approx(float(out["cnt"][2]), rel=1e-9)
    st.close()


@pytest.mark.parametrize("kernel", ["pass", "stream"])
@pytest.mark.parametrize("world,N,n_sweeps,calls", [(2, (128, 64, 16), 3, 3), (4, (128, 64, 32), 3, 2),
                                                    (8, (128, 64, 64), 3, 2), (2, (512, 512, 128), 2, 2),
                                                    (4, (512, 512, 256), 2, 2), (8, (512, 512, 512), 2, 2)])
def test_slabs_over_peer_memory_equal_one_gpu(world, N, n_sweeps, calls, kernel):
    """2 / 4 / 8 slabs over NVLink peer memory (boundary rows are stored into the ring
    neighbours' ghost layers by the sweep kernel itself; ordered by ring epochs in the
    colour-pass kernel, by the neighbours' layer counters in the streaming kernel) leave the SAME
    occupation and acceptance counts as one GPU: small boxes whose slabs are a few layers
    thick (every unit touches a ghost layer or waits for one), and 64-layer slabs of
    512 x 512 layers -- the decomposition of BASELINE configs[2] (the 8-GPU case IS the
    512^3 box).  Several calls in a row: the counters carry over between launches."""
    import torch
    import torch.multiprocessing as mp
    from casmcode_clexmonte_b200 import _capi
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    from casmcode_clexmonte_b200.potential import semigrand_exchange_table
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29700 + os.getpid() % 1000 + world + (50 if kernel == "stream" else 0)
    from casmcode_clexmonte_b200 import _capi as _c
    flags = _c.CMX_SWEEP_STREAM if kernel == "stream" else 0
    mp.spawn(_worker, args=(world, port, N, n_sweeps, True, out, calls, flags), nprocs=world, join=True)
    sysd = json.loads((GOLDEN / "systems.json").read_text())["fcc"]
    tables = _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / "fcc_default.npz"))
    st = _capi.State(tables, N)
    st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
    st.set_conditions(900.0, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.1, -0.2], 3))
    st.upload_occ(np.random.default_rng(4).integers(0, 3, int(np.prod(N))).astype(np.int8))
    cnt = st.sgc_sweep(n_sweeps * calls, seed=17)
    occ = st.download_occ(dtype=np.int8)
    assert (occ == out["occ"]).all(), f"{(occ != out['occ']).sum()} sites differ"
    assert cnt[0].n_accept == int(out["cnt"][1]) and cnt[0].n_attempt == int(out["cnt"][0])
    st.close()


def test_one_slab_with_fused_halo_equals_periodic_box():
    """A single slab with ghost layers whose 'neighbours' are itself (the fused
    peer-memory halo push writing its own ghost layers) reproduces the plain
    periodic box -- runs on one GPU."""
    import torch
    from casmcode_clexmonte_b200 import _capi
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    from casmcode_clexmonte_b200.potential import semigrand_exchange_table
    from casmcode_clexmonte_b200.slab import SlabRunner
    N, n_sweeps = 32, 5
    sysd = json.loads((GOLDEN / "systems.json").read_text())["fcc"]
    tables = _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / "fcc_default.npz"))
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.1, -0.2], 3)
    init = np.random.default_rng(4).integers(0, 3, N ** 3).astype(np.int8)
    res = {}
    for p2p, flags in ((True, 0), (False, 0), ("stream", _capi.CMX_SWEEP_STREAM), ("stream_host", _capi.CMX_SWEEP_STREAM)):
        run = SlabRunner(tables, N, sysd["eci_sparse"], 900.0, ex, 0, 1, 0, init_occ=init,
                         p2p=(p2p is True or p2p == "stream"))
        run.state.set_sweep_flags(flags)
        run.state.counters_reset()
        run.sweep(n_sweeps, seed=17)
        run.synchronize()
        res[p2p] = (run.download_local(), run.state.counters_read()[0].n_accept)
        run.state.close()
    st = _capi.State(tables, (N, N, N))
    st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
    st.set_conditions(900.0, ex)
    st.upload_occ(init)
    cnt = st.sgc_sweep(n_sweeps, seed=17)
    ref = st.download_occ(dtype=np.int8)
    for p2p in res:
        assert (res[p2p][0] == ref).all(), f"p2p={p2p}"
        assert res[p2p][1] == cnt[0].n_accept
    st.close()


def test_one_slab_of_the_dense_eci_equals_periodic_box():
    """The reference's dense FCC ECI (1NN + 2NN pairs): a box runs the two-class count-table
    kernel, a slab of it (ghost layers, halo exchanged by the host path) the pair-sum kernel --
    with the same random bits, so the decomposition does not change the trajectory.  One GPU."""
    from casmcode_clexmonte_b200 import _capi
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    from casmcode_clexmonte_b200.potential import semigrand_exchange_table
    from casmcode_clexmonte_b200.slab import SlabRunner
    N, n_sweeps = 32, 4
    sysd = json.loads((GOLDEN / "systems.json").read_text())["fcc"]
    tables = _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / "fcc_default.npz"))
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.1, -0.2], 3)
    init = np.random.default_rng(4).integers(0, 3, N ** 3).astype(np.int8)
    run = SlabRunner(tables, N, sysd["eci_full"], 900.0, ex, 0, 1, 0, init_occ=init, p2p=False)
    assert run.state.sweep_info()["evaluator"] == "pair_sum"
    run.state.counters_reset()
    run.sweep(n_sweeps, seed=17)
    run.synchronize()
    got, acc = run.download_local(), run.state.counters_read()[0].n_accept
    run.state.close()
    st = _capi.State(tables, (N, N, N))
    st.set_eci(sysd["eci_full"]["index"], sysd["eci_full"]["value"])
    st.set_conditions(900.0, ex)
    st.upload_occ(init)
    assert st.sweep_info()["evaluator"] == "pair_lut2"
    cnt = st.sgc_sweep(n_sweeps, seed=17)
    ref = st.download_occ(dtype=np.int8)
    assert (got == ref).all(), f"{(got != ref).sum()} sites differ"
    assert acc == cnt[0].n_accept
    st.close()
