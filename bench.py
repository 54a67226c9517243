#!/usr/bin/env python
"""bench.py -- attempted MC steps/s of the semi-grand canonical checkerboard sweep.

Workload (BASELINE.json configs[2], the config the metric is quoted on; it fits
one B200): FCC ternary (A-B-Va) semi-grand canonical, 512^3 primitive supercell
(134 217 728 sites), the reference's shipped ECI (points + nearest-neighbour
pairs), T = 800 K, param_chem_pot = (0, 0), i.i.d. random initial occupation.
A "step" of this benchmark = `sweeps_per_step` full lattice sweeps (each sweep =
one attempted Metropolis step at every site).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun, one rank per GPU): the 512^3 box is split into N slabs along k
with one ghost layer each side; boundary rows travel over NVLink peer memory inside
the sweep kernel (strong scaling).  --impl reference times the reference's own CPU path (the
generated Clexulator kernels of oracle/_ref inside the restated sequential
loop), one chain per host core.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"

METRIC = "attempted MC steps/sec (FCC ternary SGC, 512^3 = 134M sites)"
UNIT = "steps/s"
N_BOX = 512
TEMPERATURE = 800.0
MU = (0.0, 0.0)
SWEEPS_PER_STEP = 10


def load_system():
    sysd = json.loads((GOLDEN / "systems.json").read_text())["fcc"]
    return sysd


from casmcode_clexmonte_b200.clocks import ClockSampler  # noqa: E402


# ---------------------------------------------------------------------------
# reference CPU path (oracle/_ref), one chain per core
# ---------------------------------------------------------------------------
def cpu_reference_rate(seconds_budget: float = 12.0, box: int = 64, threads: int | None = None,
                       eci_key: str = "eci_sparse") -> dict:
    """Sequential semi-grand Metropolis with the reference's generated kernels
    (methods/occupation_metropolis.hh:92-120 restated in oracle/harness.cpp), one
    independent chain per host core, each on a `box`^3 periodic sample of the
    workload (a 512^3 SuperNeighborList would need 20 GB per chain)."""
    from oracle import oracle as O
    kind = "reference"
    if not O.available("fcc_default"):
        raise RuntimeError("oracle/_ref not built")
    sysd = load_system()
    eci = sysd[eci_key]
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=3, Rt=np.array(sysd["axes"]["Rt"]))
    cores = threads or len(os.sched_getaffinity(0))
    clex = O.RefClexulator("fcc_default")
    scs = [clex.supercell(box) for _ in range(cores)]
    occs = [np.random.default_rng(c).integers(0, 3, box ** 3).astype(np.int32) for c in range(cores)]
    # calibrate on one core
    r = scs[0].metropolis_run(0, occs[0], prim, eci["index"], eci["value"], TEMPERATURE, 1, 200_000,
                              param_chem_pot=np.array(MU))
    rate1 = 200_000 / max(r["seconds"], 1e-9)
    n_steps = int(max(200_000, rate1 * seconds_budget))
    out = [None] * cores

    def work(c):
        out[c] = scs[c].metropolis_run(0, occs[c], prim, eci["index"], eci["value"], TEMPERATURE, 100 + c,
                                       n_steps, param_chem_pot=np.array(MU))

    ths = [threading.Thread(target=work, args=(c,)) for c in range(cores)]
    t0 = time.time()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.time() - t0
    total = n_steps * cores
    return dict(value=total / wall, unit=UNIT, cores=cores, kind=kind,
                sample=f"{cores} independent chains x {n_steps} sequential steps on a {box}^3-site periodic "
                       f"box each (same basis/ECI/T/mu; wall {wall:.1f}s)",
                steps_per_s_per_core=total / wall / cores)


# ---------------------------------------------------------------------------
# end-to-end legs: host buffers through the C ABI, copies inside the timed region
# ---------------------------------------------------------------------------
def e2e_single(torch, _capi, st, tables, eci, ex, N, K, S, sweep_flags) -> dict:
    """The headline e2e is the leg a reference-side plugin would run: ONE state, every step
    uploads the int32 occupation the previous step downloaded (the reference hands the
    calculator an Eigen::VectorXi, int32), sweeps, downloads int32 occupation + counters;
    nothing can overlap.  Reported beside it: the same dependent leg with int8 host buffers
    (4x fewer PCIe bytes) and a pipeline of independent jobs (three states in flight)."""
    n_sites = N ** 3

    def dependent(dtype, n_steps):
        host = torch.empty(n_sites, dtype={np.int8: torch.int8, np.int32: torch.int32}[dtype]).pin_memory()
        harr = host.numpy()
        st.download_occ(dtype=dtype, out=harr)
        for _ in range(1):
            st.upload_occ(harr)
            st.sgc_sweep(S, seed=3, first_sweep=0, counters=True)
            st.download_occ(dtype=dtype, out=harr)
        torch.cuda.synchronize()
        te0 = time.perf_counter()
        for k in range(n_steps):
            st.upload_occ(harr)
            st.sgc_sweep(S, seed=3, first_sweep=(k + 1) * S, counters=True)
            st.download_occ(dtype=dtype, out=harr)
        torch.cuda.synchronize()
        te1 = time.perf_counter()
        b = n_sites * np.dtype(dtype).itemsize
        return {"value": n_steps * S * n_sites / (te1 - te0), "unit": UNIT, "h2d_bytes_per_step": b,
                "d2h_bytes_per_step": b + 32, "ms_per_step": (te1 - te0) * 1e3 / n_steps}

    e2e = dependent(np.int32, K)
    e2e["note"] = ("one state, every step uploads the int32 occupation the previous step downloaded "
                   "(the reference's Eigen::VectorXi), S sweeps, downloads int32 occupation + counters")
    e2e["int8_dependent"] = dependent(np.int8, K)
    # pipelined: every step is an independent job (its own pinned input and output buffers),
    # three states in flight: the upload of job k+1, the sweeps of job k and the download of
    # job k-1 overlap (asynchronous C ABI, one stream per state)
    n_buf = 3
    states = [st]
    for _ in range(n_buf - 1):
        s2 = _capi.State(tables, (N, N, N))
        s2.set_eci(eci["index"], eci["value"])
        s2.set_conditions(TEMPERATURE, ex)
        s2.set_sweep_flags(sweep_flags)
        states.append(s2)
    h_in = [torch.empty(n_sites, dtype=torch.int8).pin_memory().numpy() for _ in range(n_buf)]
    h_out = [torch.empty(n_sites, dtype=torch.int8).pin_memory().numpy() for _ in range(n_buf)]
    st.download_occ(dtype=np.int8, out=h_in[0])
    for b in range(1, n_buf):
        h_in[b][:] = h_in[0]

    def pipeline(n_jobs):
        acc = 0
        for k in range(n_jobs + n_buf):
            b = k % n_buf
            if k >= n_buf:                      # job k - n_buf: its result is read here
                acc += states[b].counters_read()[0].n_accept
            if k < n_jobs:
                states[b].upload_occ_async(h_in[b])
                states[b].sgc_sweep_async(S, seed=3, first_sweep=(k + 1) * S)
                states[b].download_occ_async(h_out[b])
        return acc

    pipeline(n_buf)
    torch.cuda.synchronize()
    tp0 = time.perf_counter()
    acc = pipeline(K)
    torch.cuda.synchronize()
    tp1 = time.perf_counter()
    assert acc > 0 and (h_out[0] != h_in[0]).any()
    e2e["int8_pipelined"] = {"value": K * S * n_sites / (tp1 - tp0), "unit": UNIT,
                             "h2d_bytes_per_step": n_sites, "d2h_bytes_per_step": n_sites + 32,
                             "ms_per_step": (tp1 - tp0) * 1e3 / K,
                             "note": "independent jobs, 3 states in flight (upload / sweeps / download overlap)"}
    for s2 in states[1:]:
        s2.close()
    return e2e


def bench_slabs(torch, dist, runner, K: int, W: int, S: int, no_e2e: bool = False) -> dict:
    """N > 1: the box cut into k-slabs, one rank per GPU (casmcode_clexmonte_b200.slab)."""
    for w in range(W):
        runner.sweep(S, seed=1, first_sweep=w * S)
    runner.synchronize()
    runner.state.counters_reset()
    dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(torch.cuda.current_device())
    clocks.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    ev0.record(runner.stream)
    for k in range(K):
        runner.sweep(S, seed=1, first_sweep=(W + k) * S)
    ev1.record(runner.stream)
    runner.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=runner.mem.device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    cnt = runner.counters()
    clk = clocks.stop(t0, t1)
    n_sites = runner.N[0] * runner.N[1] * runner.N[2] * runner.n_sublat
    stream = bool(runner.p2p and runner.info()["one_launch_per_call"])
    lps = runner.info()["launches_per_sweep"]
    launches = K if stream else K * S * max(lps, 1)
    e2e = None
    if not no_e2e:
        # end to end: every rank's int32 host slab in, S sweeps, int32 slab out, every step
        host = torch.empty(runner.layer * runner.n2 * runner.n_sublat, dtype=torch.int32).pin_memory()
        harr = host.numpy()
        runner.state.download_occ(dtype=np.int32, out=harr)
        dist.barrier()
        torch.cuda.synchronize()
        te0 = time.perf_counter()
        for k in range(K):
            runner.state.upload_occ(harr)
            runner.exchange(None)
            runner.sweep(S, seed=3, first_sweep=k * S)
            runner.state.download_occ(dtype=np.int32, out=harr)
        runner.synchronize()
        dist.barrier()
        te = torch.tensor([time.perf_counter() - te0], dtype=torch.float64, device=runner.mem.device)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": K * S * n_sites / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": 4 * n_sites, "d2h_bytes_per_step": 4 * n_sites,
               "ms_per_step": float(te.item()) * 1e3 / K,
               "note": "every rank: int32 host slab uploaded, ghost layers exchanged (NCCL), S sweeps, "
                       "int32 slab downloaded, every step"}
    return dict(ms=float(ms.item()), clocks=clk, accept_rate=float(cnt[1] / max(cnt[0], 1.0)),
                launches=launches, kernel_ms=float(ms.item()) / launches, stream=stream,
                sweeps_per_launch=(float(S) if stream else 1.0 / max(lps, 1)), e2e=e2e)


# ---------------------------------------------------------------------------
# replica grids / KMC trajectories dealt over the GPUs (BASELINE configs[1], [4])
# ---------------------------------------------------------------------------
def bench_replicas(args, torch, _capi, rank, world, local_rank):
    """Independent replicas: every rank runs its share, ONE NCCL all-reduce of statistics at
    the end (inside the timed region).  Whole-job rate = all replicas / max-over-ranks time;
    the total work is fixed (strong scaling)."""
    from casmcode_clexmonte_b200 import kmc as K
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    from casmcode_clexmonte_b200.replicas import KmcEnsembleRunner, ReplicaRunner
    dist = None
    device = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        dist.all_reduce(torch.zeros(1, dtype=torch.float64, device=device))   # communicator set-up is not the workload
    sysd = load_system()
    K_, W_ = args.steps, max(3, args.warmup)
    tb = lambda n: _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / f"{n}.npz"), device=local_rank)  # noqa: E731

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "c2":
        N, S = 128, 10 * SWEEPS_PER_STEP   # a sample every 100 passes: the (mu, T) grid is a long run per point
        conds = [{"temperature": float(T), "param_chem_pot": [float(mu), 0.0]}
                 for mu in np.linspace(-1, 1, 8) for T in np.arange(400.0, 1801.0, 200.0)]
        run = ReplicaRunner(tb("fcc_default"), (N, N, N), sysd, sysd["eci_sparse"], conds, rank, world,
                            n_samples=K_ + 1, seed_init=7)
        run.state.sgc_sweep(W_ * SWEEPS_PER_STEP, seed=1, counters=False)
        run.sampler.run(1, 1, seed=1 + rank, first_sweep=W_ * SWEEPS_PER_STEP)
        run.reduce(dist, device)            # warm-up of the collective at its real size
        run.sampler.reset()
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        time.sleep(0.3)
        stream = torch.cuda.ExternalStream(run.state.stream())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        ev0.record(stream)
        cnt = run.sampler.run(K_, S, seed=1 + rank, first_sweep=W_ * S)   # K samples, S passes apart
        res = run.reduce(dist, device)                                     # ONE all-reduce of the moments
        ev1.record(stream)
        barrier()
        t1 = time.time()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        clk = clocks.stop(t0, t1)
        n_sites = len(conds) * N ** 3
        value = K_ * S * n_sites / (ms * 1e-3)
        info = run.state.sweep_info()
        metric = "attempted MC steps/sec (FCC ternary SGC, 64 replicas x 128^3)"
        workload = ("FCC A-B-Va semi-grand canonical, 128^3 primitive supercell x 64 replicas (8 param_chem_pot x 8 T), "
                    "sampled every 100 passes, statistics all-reduced")
        extra = {"heat_capacity_first_last": [res[0]["heat_capacity"], res[-1]["heat_capacity"]],
                 "accept_rate_local_min_max": [min(c.n_accept / c.n_attempt for c in cnt),
                                               max(c.n_accept / c.n_attempt for c in cnt)],
                 "samples_per_replica": res[0]["n_samples"]}
        unit, alg_bytes_per_unit, kernel = UNIT, 2.0, "k_sweep_pass16"
        launches = K_ * 3
        run.close()
    else:
        names = ["fcc_default"] + [f"fcc_{ev}_{k}" for ev in ("A_Va_1NN", "B_Va_1NN") for k in range(6)]
        tbs = {n: tb(n) for n in names}
        types = [dict(et, kra=(et["kra"]["index"], et["kra"]["value"]), freq=(et["freq"]["index"], et["freq"]["value"]))
                 for et in sysd["kmc"]["event_types"]]
        prim = K.make_prim_event_list(types)
        R_, N = 4096, (16, 16, 16)
        n = int(np.prod(N))
        base = np.random.default_rng(5).choice(3, size=(64, n), p=[0.899, 0.1, 0.001]).astype(np.int32)

        def factory(st):
            dev_types = [dict(local_tables=[tbs[x] for x in et["local_tables"]], kra=et["kra"], freq=et["freq"])
                         for et in types]
            return _capi.Kmc(st, dev_types, prim)

        run = KmcEnsembleRunner(tbs["fcc_default"], N, sysd["eci_dense"], factory, 1200.0, R_, rank, world, seed0=1,
                                occ_of=lambda i: base[i % 64])
        S = 100
        run.reduce(run.run(W_ * S), dist, device)   # warm-up, the collective at its real size included
        barrier()
        clocks = ClockSampler(local_rank)
        clocks.start()
        time.sleep(0.3)
        stream = torch.cuda.ExternalStream(run.state.stream())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        ev0.record(stream)
        out = run.run(K_ * S)
        table = run.reduce(out, dist, device)                               # ONE all-reduce of (steps, time, rate)
        ev1.record(stream)
        barrier()
        t1 = time.time()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        clk = clocks.stop(t0, t1)
        hops = float(table[:, 0].sum()) - W_ * S * R_
        value = hops / (ms * 1e-3)
        metric = "KMC events/sec (FCC A-B-Va, 4096 trajectories x 4096 cells)"
        workload = ("FCC A-B-Va rejection-free KMC, 16^3 primitive cells per trajectory, x=(0.899, 0.1, 0.001), "
                    "T=1200 K, 4096 trajectories dealt over the GPUs, (steps, time) all-reduced")
        extra = {"mean_time_per_trajectory": float(table[:, 1].mean()), "trajectories": R_}
        unit, alg_bytes_per_unit, kernel = "events/s", 708 * 90.0, "k_kmc_run"
        launches = 1
        info = {"evaluator": "kmc"}
        run.close()
    if rank != 0:
        return
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks_file.read_text())["hbm_gbs"] if peaks_file.exists() else 6650.0
    achieved = alg_bytes_per_unit * value / world / 1e9
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": K_, "warmup": W_,
            "ms_per_step": ms / K_, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "parallelism": f"replicas over {world} rank(s), round robin",
                       "collective": "one NCCL all-reduce of statistics (inside the timed region)"},
            "clocks": clk, "e2e": None, "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": kernel,
                         "note": ("per GPU; the KMC step is latency bound (one block per trajectory, a serial event "
                                  "chain): the HBM figure is the contract's, not the binding resource")
                         if args.workload == "c5" else "per GPU, 2 B per attempted step"},
            "evaluator": info["evaluator"]}
    line.update(extra)
    print(json.dumps(line))


# ---------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c2", "c5"],
                    help="c3 (default, the headline): 512^3 box, slabs over the GPUs; c2: 64-replica (mu, T) grid of "
                         "128^3 boxes dealt over the GPUs; c5: 4096 KMC trajectories dealt over the GPUs")
    ap.add_argument("--eci", default="sparse", choices=["sparse", "full"],
                    help="c3 only: the reference's sparse FCC ECI (points + 1NN pairs: the headline) or all nine functions "
                         "(points + 1NN + 2NN pairs, SURVEY 8d's 19-site neighbourhood: the two-class count table; one GPU)")
    ap.add_argument("--box", type=int, default=N_BOX)
    ap.add_argument("--layers", type=int, default=0, help="tuning runs: N2 of a (box, box, layers) supercell on one GPU")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-flags", type=int, default=0, help="CMX_SWEEP_* bits (1 = dE sum, 2 = generic evaluator, 8 = streaming kernel)")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs: skip the end-to-end leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    eci_key = "eci_full" if args.eci == "full" else "eci_sparse"
    eci_text = ("all nine functions of the FCC basis (points + 1NN + 2NN pairs, 19-site neighbourhood)" if args.eci == "full"
                else "shipped sparse ECI (points+1NN pairs)")
    workload = (f"FCC A-B-Va semi-grand canonical, {args.box}^3 primitive supercell, {eci_text}, "
                f"T={TEMPERATURE:g} K, param_chem_pot={list(MU)}")
    config = {"workload": workload, "sites": args.box ** 3, "sweeps_per_step": SWEEPS_PER_STEP,
              "l2_policy": "inputs larger than L2 (134 MB lattice, streamed once per sweep and direction)",
              "parallelism": f"slab{world}" if world > 1 else "single"}

    if args.impl == "reference":
        if rank != 0:
            return
        res = cpu_reference_rate(seconds_budget=max(5.0, min(60.0, 4.0 * args.steps)), eci_key=eci_key)
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": res,
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    from casmcode_clexmonte_b200 import _capi
    from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables
    from casmcode_clexmonte_b200.potential import semigrand_exchange_table

    if not torch.cuda.is_available() or _capi.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path")
    torch.cuda.set_device(local_rank)
    if args.workload != "c3":
        return bench_replicas(args, torch, _capi, rank, world, local_rank)
    if world > 1:
        if args.eci != "sparse":
            raise SystemExit("bench.py: --eci full runs on one GPU (slab states take one neighbor class)")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from casmcode_clexmonte_b200.slab import SlabRunner
    sysd = load_system()
    eci = sysd[eci_key]
    tables = _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / "fcc_default.npz"), device=local_rank)
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], MU, 3)
    N = args.box
    K, W, S = args.steps, max(3, args.warmup), SWEEPS_PER_STEP

    n_sites = N ** 3
    if world == 1:
        if args.layers:
            n_sites = N * N * args.layers
        st = _capi.State(tables, (N, N, args.layers or N))
        st.set_eci(eci["index"], eci["value"])
        st.set_conditions(TEMPERATURE, ex)
        st.randomize(2026)
        st.set_sweep_flags(args.sweep_flags)
        info = st.sweep_info()
        stream = torch.cuda.ExternalStream(st.stream())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # ---- device-resident throughput: K steps, each one call of S sweeps (the streaming
        # kernel runs a call as ONE cooperative launch; the other evaluators launch per colour)
        for w in range(W):   # the same calls as the timed steps (the unit list of a call is cached)
            st.sgc_sweep(S, seed=1, first_sweep=w * S, counters=False)
        torch.cuda.synchronize()
        clocks = ClockSampler(local_rank)
        clocks.start()
        time.sleep(0.3)
        st.counters_reset()
        t0 = time.time()
        ev0.record(stream)
        for k in range(K):
            st.sgc_sweep_enqueue(S, seed=1, first_sweep=(W + k) * S)
        ev1.record(stream)
        torch.cuda.synchronize()
        t1 = time.time()
        cnt = st.counters_read()
        ms = ev0.elapsed_time(ev1)
        clk = clocks.stop(t0, t1)
        attempts = K * S * n_sites
        assert cnt[0].n_attempt == attempts
        value = attempts / (ms * 1e-3)
        lps = info["launches_per_sweep"]
        sweep_launches = K if lps == 0 else K * S * lps
        sweeps_per_launch = float(S) if lps == 0 else 1.0 / lps
        launches = sweep_launches + 1          # + the counter reduction
        kernel_ms = ms / sweep_launches
        kernel_name = ("k_sweep_stream16" if info["stream"] else "k_sweep_pass16" if info["one_launch_per_call"]
                       else "k_sweep_pair16" if info["evaluator"] == "pair_lut" else "k_sweep_generic")
        accept_rate = cnt[0].n_accept / cnt[0].n_attempt
        e2e = None
        if not args.no_e2e and not args.layers:
            e2e = e2e_single(torch, _capi, st, tables, eci, ex, N, K, S, args.sweep_flags)
        st.close()
    else:
        import torch.distributed as dist
        runner = SlabRunner(tables, N, eci, TEMPERATURE, ex, rank, world, local_rank, seed_init=2026)
        info = runner.info()
        res = bench_slabs(torch, dist, runner, K, W, S, no_e2e=args.no_e2e)
        ms, clk, e2e, accept_rate = res["ms"], res["clocks"], res["e2e"], res["accept_rate"]
        value = K * S * n_sites / (ms * 1e-3)
        launches, kernel_ms, sweeps_per_launch = res["launches"], res["kernel_ms"], res["sweeps_per_launch"]
        kernel_name = ("k_sweep_stream16" if info["stream"] else "k_sweep_pass16") if res["stream"] else "k_sweep_pair16"
        if rank != 0:
            return

    # ---- roofline of the dominant kernel: algorithmic bytes of ONE launch / its duration
    peaks_file = ROOT / "MEASURED_PEAKS.json"
    if peaks_file.exists():
        peak, peak_src = json.loads(peaks_file.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    sites_per_launch = n_sites / world * sweeps_per_launch   # attempted steps of one launch (one rank)
    # DRAM traffic of one launch and the binding resource from the committed ncu --set full
    # capture of this kernel at this workload (512^3, one GPU, S sweeps per launch); null otherwise
    traffic, binding = None, None
    prof = ROOT / "profiles" / (f"r02_ncu_full_{kernel_name}_two_class.csv" if info["evaluator"] == "pair_lut2"
                                else f"r02_ncu_full_{kernel_name}.csv")
    if world == 1 and args.box == N_BOX and not args.layers and prof.exists():
        import csv
        m = {}
        for row in csv.reader(prof.open()):   # (the last launch of the capture wins)
            if len(row) == 5 and row[0].isdigit():
                m[row[2]] = (row[3], row[4])
        try:
            unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            traffic = sum(float(m[k][1]) * unit[m[k][0]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            binding = {"resource": "instruction issue slots (ALU pipe)", "unit": "% of peak sustained",
                       "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]),
                       "alu_pipe_pct": float(m["sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"][1]),
                       "dram_pct": float(m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][1]),
                       "source": f"profiles/{prof.name} (ncu --set full, one launch = {S} sweeps)"}
        except (KeyError, ValueError):
            pass
    alg_bytes = 2.0 * sites_per_launch           # 2 B per step at the HBM level
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": f"bytes per launch, dram read + write, profiles/{prof.name}",
                "algorithmic_bytes_per_launch": alg_bytes, "binding_resource": binding, "peak_source": peak_src,
                "kernel": kernel_name, "kernel_ms": kernel_ms, "sweeps_per_launch": sweeps_per_launch,
                "algorithmic_bytes_per_step_hbm": 2.0,
                "algorithmic_bytes_per_step_l2": info["bytes_per_step"]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": clk, "e2e": e2e,
            "gpu_launches": launches, "roofline": roofline, "evaluator": info["evaluator"],
            "accept_rate": accept_rate, "dE_evals_per_s": value,
            "schedule": {k: info.get(k) for k in ("stream", "stream_blocks", "stream_group_rowsteps",
                                                  "stream_gap_units", "launches_per_sweep")}}
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_reference_rate(seconds_budget=args.cpu_seconds, eci_key=eci_key)
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                                    "sample": f"unavailable: {e}"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
