#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs[1], [3], [4]) and the sampling pass,
measured on one B200 through the C ABI with CUDA events on the state's stream.
bench.py stays the headline (configs[2]); this prints one JSON line per workload.

    python tools/bench_workloads.py [c2] [sample] [c4] [c5]      (default: all)

SURVEY.md section 8(d) per-unit figures are used for `achieved`:
  c2     FCC ternary SGC, 128^3 x 64 (mu, T) replicas            2 B / step at HBM
  sample energy + composition sample of 512^3 (one fused pass)    1 B / site at HBM
  c4     ZrO canonical O<->Va exchanges, quadruplet basis         226 B / single-site dE (L2)
  c5     FCC A-B-Va KMC event rates, 4096 trajectories            90 B / allowed event
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"

import torch  # noqa: E402

from casmcode_clexmonte_b200 import _capi, kmc as K  # noqa: E402
from casmcode_clexmonte_b200.clexulator_tables import ClexulatorTables  # noqa: E402
from casmcode_clexmonte_b200.potential import canonical_swap_types, semigrand_exchange_table  # noqa: E402

SYS = json.loads((GOLDEN / "systems.json").read_text())
PEAKS = ROOT / "MEASURED_PEAKS.json"
HBM = json.loads(PEAKS.read_text())["hbm_gbs"] if PEAKS.exists() else 6650.0


def tables(name):
    return _capi.Tables(ClexulatorTables.load(GOLDEN / "tables" / f"{name}.npz"))


def o2s(sysd, max_occ):
    a = np.full((len(sysd["occ_to_species"]), max_occ), -1, dtype=np.int32)
    for b, row in enumerate(sysd["occ_to_species"]):
        a[b, :len(row)] = row
    return a


def timed(st, fn, reps=1):
    stream = torch.cuda.ExternalStream(st.stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        out = fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def emit(**kw):
    print(json.dumps(kw), flush=True)


def c2():
    sysd = SYS["fcc"]
    t = tables("fcc_default")
    R, N = 64, 128
    st = _capi.State(t, (N, N, N), R)
    st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
    st.set_occupants(sysd["sublat_to_asym"], o2s(sysd, 3), 3)
    sm = _capi.Sampler(st, 64, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    r = 0
    for mu in np.linspace(-1, 1, 8):
        for T in np.arange(400.0, 1801.0, 200.0):
            st.set_conditions(T, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], (mu, 0.0), 3), r)
            sm.set_param_chem_pot([mu, 0.0], r)
            r += 1
    st.randomize(7)
    st.sgc_sweep(10, seed=1)
    S = 50
    ms, cnt = timed(st, lambda: st.sgc_sweep(S, seed=1, first_sweep=10))
    steps = S * R * N ** 3
    rate = steps / (ms * 1e-3)
    emit(workload="c2: FCC A-B-Va SGC, 128^3 x 64 replicas (8 mu x 8 T), shipped sparse ECI", metric="attempted MC steps/s",
         value=rate, ms=ms, sweeps=S, launches=1, kernel="k_sweep_pass16",
         roofline={"bound": "hbm", "achieved": 2.0 * rate / 1e9, "peak": HBM, "unit": "GB/s", "frac": 2.0 * rate / 1e9 / HBM},
         accept_min=min(c.n_accept / c.n_attempt for c in cnt), accept_max=max(c.n_accept / c.n_attempt for c in cnt))
    # the same run sampled every pass (device-side samplers, one sync at the end)
    sm.run(2, 1, seed=1, first_sweep=60)
    sm.reset()
    ms2, _ = timed(st, lambda: sm.run(S, 1, seed=1, first_sweep=62))
    emit(workload="c2 sampled every pass: potential_energy, formation_energy, mol/param composition of all 64 replicas",
         metric="attempted MC steps/s", value=steps / (ms2 * 1e-3), ms=ms2, samples=S,
         sampling_overhead=ms2 / ms - 1.0, heat_capacity_replica0=sm.analysis(0)["heat_capacity"])
    sm.close()
    st.close()
    t.close()


def c3full():
    """BASELINE configs[2] with ALL nine functions of the FCC basis selected (constant, points,
    1NN and 2NN pairs -- the 19-site neighbourhood SURVEY section 8d quotes; coefficient values
    of the golden fixture `eci_full`): two neighbour classes.  Default evaluator: the colour-pass
    kernel with the two-class count table ("pair_lut2"); beside it the per-neighbor-table kernel
    it replaced (CMX_SWEEP_PAIR_SUM, one site per thread)."""
    sysd = SYS["fcc"]
    t = tables("fcc_default")
    for N in (256, 512):
        for flags in (0, _capi.CMX_SWEEP_PAIR_SUM):
            st = _capi.State(t, (N, N, N), 1)
            st.set_eci(sysd["eci_full"]["index"], sysd["eci_full"]["value"])
            st.set_conditions(800.0, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], (0.0, 0.0), 3))
            st.set_sweep_flags(flags)
            st.randomize(7)
            st.sgc_sweep(2, seed=1)
            S = 5 if flags else 10
            ms, cnt = timed(st, lambda: st.sgc_sweep(S, seed=1, first_sweep=2))
            rate = S * N ** 3 / (ms * 1e-3)
            info = st.sweep_info()
            two = info.get("evaluator") == "pair_lut2"
            emit(workload=f"c3 all functions: FCC A-B-Va SGC, {N}^3, points + 1NN + 2NN pairs (19-site neighbourhood)",
                 metric="attempted MC steps/s", value=rate, ms=ms, sweeps=S, evaluator=info.get("evaluator"),
                 kernel="k_sweep_pass16<3, fcc 1NN, fcc 2NN>" if two else "k_sweep_pairsum",
                 launches=1 if two else 8 * S,
                 bytes_per_step_l2=info.get("bytes_per_step"), flops_per_step=info.get("flops_per_step"),
                 roofline={"bound": "hbm", "achieved": 2.0 * rate / 1e9, "peak": HBM, "unit": "GB/s",
                           "frac": 2.0 * rate / 1e9 / HBM,
                           "note": ("2 B per step at HBM, 20 B per step through L2 / shared memory; issue bound "
                                    "(about 28 instructions per step: two byte-lane sums, three index-map lookups and one table lookup per site)") if two else
                                   "2 B per step at HBM; the pair-sum evaluator is issue bound (20 B per step through L1/L2, ~550 instructions)"},
                 accept_rate=cnt[0].n_accept / cnt[0].n_attempt)
            st.close()
    t.close()


def c1():
    """FCC A-B canonical pair exchanges (BASELINE configs[0] method) at GPU scale."""
    sysd = SYS["fcc"]
    t = tables("fcc_default")
    for N, R in ((128, 1), (256, 1), (32, 64)):
        n = N ** 3
        st = _capi.State(t, (N, N, N), R)
        st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
        m = o2s(sysd, 3)
        st.set_occupants(sysd["sublat_to_asym"], m, 3)
        rng = np.random.default_rng(2)
        occ = (rng.random(n) < 0.5).astype(np.int32)
        for r in range(R):
            st.upload_occ(occ, r)
            st.set_conditions(600.0 + 20.0 * r, None, r)
        swaps = canonical_swap_types(st.tables.host, sysd["sublat_to_asym"], m.tolist(), (N, N, N))
        st.canonical_set_swaps(swaps)
        st.canonical_sweep(1, seed=1)
        S = 3
        ms, cnt = timed(st, lambda: st.canonical_sweep(S, seed=1, first_sweep=1))
        att = float(sum(c.n_attempt for c in cnt))
        emit(workload=f"c1: FCC A-B canonical pair exchanges, {N}^3 sites x {R} replicas, shipped sparse ECI, {len(swaps)} swap types",
             metric="attempted MC steps/s (each a two-site dE)", value=att / (ms * 1e-3), ms=ms, sweeps=S,
             accept_rate=sum(c.n_accept for c in cnt) / att, kernel="k_canonical_pairs")
        st.close()
    t.close()


def sample():
    sysd = SYS["fcc"]
    t = tables("fcc_default")
    N = 512
    st = _capi.State(t, (N, N, N), 1)
    st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
    st.set_occupants(sysd["sublat_to_asym"], o2s(sysd, 3), 3)
    st.set_conditions(800.0, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], (0.0, 0.0), 3))
    st.randomize(3)
    sm = _capi.Sampler(st, 64, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    sm.set_param_chem_pot([0.0, 0.0])
    for _ in range(3):
        sm.sample()
    ms, _ = timed(st, sm.sample, reps=20)
    gbs = N ** 3 / (ms * 1e-3) / 1e9
    emit(workload="sample: potential_energy + compositions of one 512^3 replica (k_energy_row16: integer bond + occupant counts)",
         metric="sites sampled/s", value=N ** 3 / (ms * 1e-3), ms=ms, kernel="k_energy_row16",
         roofline={"bound": "hbm", "achieved": gbs, "peak": HBM, "unit": "GB/s", "frac": gbs / HBM,
                   "algorithmic_bytes_per_site": 1.0})
    # global correlations (all 9 functions) by integer bond counts of both FCC pair shells in one pass
    st.global_corr()
    ms, _ = timed(st, st.global_corr, reps=20)
    gbs = N ** 3 / (ms * 1e-3) / 1e9
    emit(workload="corr: Correlations::per_supercell of one 512^3 replica, all 9 functions (k_energy_row16 with both pair shells + k_corr_lin_final; includes the 72-byte result read-back)",
         metric="sites/s", value=N ** 3 / (ms * 1e-3), ms=ms, kernel="k_energy_row16<3, fcc 1NN, fcc 2NN>",
         roofline={"bound": "hbm", "achieved": gbs, "peak": HBM, "unit": "GB/s", "frac": gbs / HBM,
                   "algorithmic_bytes_per_site": 1.0})
    sm.close()
    st.close()
    t.close()


def c4():
    sysd = SYS["zro"]
    t = tables("zro")
    for N in (24, 48, 96):
        n_cells = N ** 3
        st = _capi.State(t, (N, N, N), 1)
        st.set_eci(sysd["eci"]["index"], sysd["eci"]["value"])
        m = o2s(sysd, st.tables.host.max_occ)
        st.set_occupants(sysd["sublat_to_asym"], m, sysd["n_species"])
        rng = np.random.default_rng(1)
        occ = np.zeros(n_cells * len(sysd["occ_to_species"]), dtype=np.int32)
        for b in sysd["mutable_sublats"]:
            occ[b * n_cells:(b + 1) * n_cells] = rng.random(n_cells) < 0.25
        st.upload_occ(occ)
        st.set_conditions(600.0, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.0], sysd["n_species"]))
        swaps = canonical_swap_types(st.tables.host, sysd["sublat_to_asym"], m.tolist(), (N, N, N))
        st.canonical_set_swaps(swaps)
        st.canonical_sweep(1, seed=1)
        S = 3
        ms, cnt = timed(st, lambda: st.canonical_sweep(S, seed=1, first_sweep=1))
        rate = cnt[0].n_attempt / (ms * 1e-3)
        emit(workload=f"c4: ZrO canonical O<->Va pair exchanges, {N}^3 cells (4 sublattices, 2 mutable), quadruplet basis (33 ECI), T=600 K",
             metric="attempted MC steps/s (each a two-site dE)", value=rate, single_site_dcorr_per_s=2 * rate, ms=ms,
             sweeps=S, swap_types=len(swaps), accept_rate=cnt[0].n_accept / cnt[0].n_attempt, kernel="k_canonical_pairs_warp (one cooperative launch per swap type)",
             roofline=wide_roofline(st.sweep_info(), rate, evals_per_step=2))
        info = st.sweep_info()
        st.sgc_sweep(1, seed=2)
        ms, cnt = timed(st, lambda: st.sgc_sweep(S, seed=2, first_sweep=1))
        rate = cnt[0].n_attempt / (ms * 1e-3)
        emit(workload=f"c4b: ZrO semi-grand O/Va flips, {N}^3 cells, generic term-list evaluator, one site per warp ({info['n_colours']} colours)",
             metric="attempted MC steps/s", value=rate, ms=ms, sweeps=S, kernel="k_sweep_generic_warp",
             bytes_per_step=info["bytes_per_step"], flops_per_step=info["flops_per_step"],
             roofline=wide_roofline(info, rate, evals_per_step=1))
        st.close()
    t.close()


def wide_roofline(info, rate, evals_per_step):
    """Wide orbit sets (one site per warp, term tables and staged site-function values in shared
    memory): the binding resource is shared-memory bandwidth -- per merged term one packed index
    word, one weight and four staged values (6 x 8 B), per neighbor the staged values written once
    -- and, behind it, the FP64 pipe.  Peaks: 148 SMs x 128 B/clk x 1.965 GHz of shared-memory
    bandwidth (B300_MICROARCH.md: 128 B/clk/SM), 37 TFLOP/s nominal FP64 (datasheet, not measured)."""
    smem_peak = 148 * 128 * 1.965  # GB/s
    smem_bytes = evals_per_step * (info["terms_per_step"] * 48.0 + info["neighbors_per_step"] * 16.0)
    flops = evals_per_step * (info["terms_per_step"] * 5.0) + 25.0
    return {"bound": "shared-memory bandwidth", "achieved": smem_bytes * rate / 1e9, "peak": smem_peak, "unit": "GB/s",
            "frac": smem_bytes * rate / 1e9 / smem_peak, "shared_bytes_per_step": smem_bytes,
            "terms_per_single_site_dE": info["terms_per_step"], "neighbors_per_single_site_dE": info["neighbors_per_step"],
            "fp64": {"achieved_tflops": flops * rate / 1e12, "peak_tflops_nominal": 37.0, "frac": flops * rate / 1e12 / 37.0,
                     "flops_per_step": flops}}


def c4_cpu():
    """Reference kernels (oracle/_ref ZrO clexulator) in the restated sequential loop, one core."""
    try:
        from oracle import oracle as O
        if not O.available("zro"):
            raise RuntimeError("oracle/_ref zro not built")
    except Exception as e:  # noqa: BLE001
        emit(workload="c4 cpu baseline", unavailable=str(e))
        return
    sysd = SYS["zro"]
    N = 12
    n_cells = N ** 3
    rng = np.random.default_rng(1)
    occ = np.zeros(n_cells * len(sysd["occ_to_species"]), dtype=np.int32)
    for b in sysd["mutable_sublats"]:
        occ[b * n_cells:(b + 1) * n_cells] = rng.random(n_cells) < 0.25
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=sysd["n_species"], Rt=np.array(sysd["axes"]["Rt"]), origin=np.array(sysd["axes"]["origin"]))
    sc = O.RefClexulator("zro").supercell(N)
    for mode, name in ((1, "canonical"), (0, "semi-grand")):
        n_steps = 20000
        r = sc.metropolis_run(mode, occ, prim, sysd["eci"]["index"], sysd["eci"]["value"], 600.0, 7, n_steps,
                              param_chem_pot=np.array([0.0]) if mode == 0 else None)
        emit(workload=f"c4 cpu baseline: ZrO {name} sequential Metropolis, reference generated kernels, {N}^3 cells",
             metric="attempted MC steps/s", value=n_steps / r["seconds"], cores=1, kind="reference",
             sample=f"{n_steps} steps on one core")


def c5():
    sysd = SYS["fcc"]
    names = ["fcc_default"] + [f"fcc_{ev}_{k}" for ev in ("A_Va_1NN", "B_Va_1NN") for k in range(6)]
    tb = {n: tables(n) for n in names}
    types = []
    for et in sysd["kmc"]["event_types"]:
        types.append(dict(et, kra=(et["kra"]["index"], et["kra"]["value"]), freq=(et["freq"]["index"], et["freq"]["value"])))
    prim = K.make_prim_event_list(types)
    R, N = 4096, (16, 16, 16)
    n = int(np.prod(N))
    st = _capi.State(tb["fcc_default"], N, R)
    eci = sysd["eci_dense"]
    st.set_eci(eci["index"], eci["value"])
    rng = np.random.default_rng(5)
    base = rng.choice(3, size=(64, n), p=[0.899, 0.1, 0.001]).astype(np.int32)
    for r in range(R):
        st.upload_occ(base[r % 64], r)
        st.set_conditions(1200.0, None, r)
    dev_types = [dict(local_tables=[tb[x] for x in et["local_tables"]], kra=et["kra"], freq=et["freq"]) for et in types]
    kmc = _capi.Kmc(st, dev_types, prim)
    kmc.all_rates(rates=False)
    ms, (_, tot) = timed(st, lambda: kmc.all_rates(rates=False))
    n_events = R * n * len(prim)
    n_vac = int((base == 2).sum()) * (R // 64)
    emit(workload=f"c5: FCC A-B-Va KMC, event-state/rate evaluation of the complete event list, {R} trajectories x {n} cells x {len(prim)} prim events, x=(0.899,0.1,0.001), T=1200 K",
         metric="event-rate evaluations/s", value=n_events / (ms * 1e-3), ms=ms, events=n_events,
         allowed_events_upper_bound=n_vac * 12, allowed_event_rates_per_s=n_vac * 12 / (ms * 1e-3),
         kernel="k_kmc_all_rates", total_rate_mean=float(np.mean(tot)))
    # impact-list style batch: 708 events per hop and trajectory (events_System_impact_table_test.cpp:50-52)
    B = 708 * R
    uc = rng.integers(0, n, B)
    pe = rng.integers(0, len(prim), B).astype(np.int32)
    rp = np.repeat(np.arange(R, dtype=np.int32), 708)
    kmc.event_states(uc[:1000], pe[:1000], rp[:1000])
    t0 = time.perf_counter()
    s = kmc.event_states(uc, pe, rp)
    dt = time.perf_counter() - t0
    emit(workload=f"c5b: impact-list batch, 708 events x {R} trajectories (host lists in, EventState records out: e2e)",
         metric="event-state evaluations/s", value=B / dt, ms=dt * 1e3, allowed=int(s["is_allowed"].sum()),
         kernel="k_kmc_event_states")
    # whole KMC steps on the device: one block per trajectory (select, apply, 708 impacted rates, tree)
    kmc.run_begin(np.arange(R) + 1)
    kmc.run(20)
    S = 200
    ms, out = timed(st, lambda: kmc.run(S))
    hops = float(np.sum(out["n_steps"])) - 20.0 * R
    emit(workload=f"c5c: rejection-free KMC steps on the device (lotto-order sum tree, mt19937_64), {R} trajectories x {n} cells, 708 impacted events per hop",
         metric="KMC events (hops)/s", value=hops / (ms * 1e-3), ms=ms, steps_per_trajectory=S,
         impacted_event_rate_evaluations_per_s=708.0 * hops / (ms * 1e-3), kernel="k_kmc_run",
         mean_time_per_trajectory_s=float(np.mean(out["time"])))
    kmc.close()
    st.close()
    for x in tb.values():
        x.close()
    # CPU baseline: the reference's own selector + generated kernels, no Python in the loop, one core
    try:
        from oracle import oracle as O
        if not (O.available("fcc_default") and O.lotto_available()):
            raise RuntimeError("oracle/_ref not built")
        Nc = (10, 10, 10)
        beg, ent = kmc._impact
        N0, N1, N2 = Nc
        impacted = []
        for c in range(N0 * N1 * N2):
            i, j, k_ = c % N0, (c // N0) % N1, c // (N0 * N1)
            for pe in range(len(prim)):
                rows = ent[beg[pe]:beg[pe + 1]]
                cells = ((i + rows[:, 1]) % N0) + N0 * (((j + rows[:, 2]) % N1) + N1 * ((k_ + rows[:, 3]) % N2))
                impacted.append(np.unique(cells.astype(np.int64) * len(prim) + rows[:, 0]))
        occ_c = rng.choice(3, size=1000, p=[0.899, 0.1, 0.001]).astype(np.int32)
        occ_c[:1] = 2
        ref = O.KmcReference(Nc, occ_c, prim, types, eci["index"], eci["value"], 1200.0, impacted, seed=3)
        ref.run(2000)
        _, _, sec = ref.run(20000)
        emit(workload="c5 cpu baseline: lotto::RejectionFreeEventSelector (reference, unmodified) + reference generated kernels, 10^3 cells, complete event list",
             metric="KMC events (hops)/s", value=20000 / sec, cores=1, kind="reference", sample="20000 hops of one trajectory on one core")
    except Exception as e:  # noqa: BLE001
        emit(workload="c5 cpu baseline", unavailable=str(e))


def main():
    which = sys.argv[1:] or ["c1", "c2", "c3full", "sample", "c4", "c4cpu", "c5"]
    if not torch.cuda.is_available():
        raise SystemExit("bench_workloads.py: no CUDA device")
    for w in which:
        {"c1": c1, "c2": c2, "c3full": c3full, "sample": sample, "c4": c4, "c4cpu": c4_cpu, "c5": c5}[w]()


if __name__ == "__main__":
    main()
