"""Generate the committed golden fixtures (run in the build container, where
/root/reference exists):

    python tests/golden/make_golden.py

1. tests/golden/tables/*.npz   flat tables exported from the reference's generated
   Clexulator sources (tests/unit/clexmonte/data/**/*_Clexulator_*.cc) by
   casmcode_clexmonte_b200.clexulator_tables -- the sources themselves are not
   copied into this repo, and /root/reference does not exist on the GPU box.
2. tests/golden/vectors_<case>.npz   inputs + outputs of the reference's own
   generated kernels (oracle/_ref, compiled unmodified) driven by oracle/harness.cpp:
   delta corr, point corr, per-cell corr, global corr, multi-site delta E,
   sequential Metropolis trajectories (std::mt19937_64).
3. tests/golden/systems.json    the small prim / ECI / composition-axes facts the
   tests need (values read from the reference's JSON fixtures).
4. the same for the SYNTHETIC FCC binary pair + triplet basis ("fcc_synthetic" / "fcc_syn"):
   source emitted by tests/golden/make_synthetic_clexulator.py in the generated-source
   grammar, compiled into oracle/_ref like the reference's own (`... make_golden.py synthetic`
   regenerates only these).
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from casmcode_clexmonte_b200.clexulator_tables import parse_clexulator_source, read_eci  # noqa: E402
from casmcode_clexmonte_b200 import kmc as K  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = Path("/root/reference")
DATA = REF / "tests/unit/clexmonte/data"
PYDATA = REF / "python/tests/data"
OUT = Path(__file__).resolve().parent

SOURCES = {
    "fcc_default": DATA / "FCC_binary_vacancy/basis_sets/bset.default/FCC_binary_vacancy_Clexulator_default.cc",
    "zro": DATA / "Clex_ZrO_Occ/basis_sets/bset.formation_energy/ZrO_Clexulator_formation_energy.cc",
}
for ev in ("A_Va_1NN", "B_Va_1NN"):
    for k in range(6):
        SOURCES[f"fcc_{ev}_{k}"] = (DATA / f"FCC_binary_vacancy/basis_sets/bset.{ev}/{k}/"
                                    f"FCC_binary_vacancy_Clexulator_{ev}_{k}.cc")


def rt_matrix(origin, end_members):
    """CompositionConverter::dparam_dmol [EXT]: left pseudo-inverse of (end_members - origin)."""
    Q = (np.array(end_members, dtype=float) - np.array(origin, dtype=float)).T
    return np.linalg.inv(Q.T @ Q) @ Q.T


def systems():
    fcc_axes = json.loads((PYDATA / "FCC_binary_vacancy/system.json").read_text())["composition_axes"]
    zro_sys = json.loads((DATA / "Clex_ZrO_Occ/system.json").read_text())
    zro_axes = zro_sys["composition_axes"]
    fcc_sparse = json.loads((DATA / "FCC_binary_vacancy/formation_energy_sparse_eci.json").read_text())
    fcc_dense_idx, fcc_dense_val = read_eci(DATA / "FCC_binary_vacancy/formation_energy_eci.json")
    fcc2_idx, fcc2_val = read_eci(PYDATA / "FCC_binary_vacancy/formation_energy_eci.2.json")
    zro_idx, zro_val = read_eci(DATA / "Clex_ZrO_Occ/formation_energy_eci.json")

    def axes(a):
        comps = a["components"]
        flat = lambda v: [float(x) for x in np.array(v, dtype=float).reshape(-1)]
        origin = flat(a["origin"])
        ends = [flat(a[k]) for k in "abcdefgh"[:int(a["independent_compositions"])]]
        return dict(components=comps, origin=origin, end_members=ends,
                    Rt=rt_matrix(origin, ends).tolist())

    return {
        "fcc": dict(
            tables="fcc_default", n_species=3, species=["A", "B", "Va"],
            sublat_to_asym=[0], occ_to_species=[[0, 1, 2]], mutable_sublats=[0],
            axes=axes(fcc_axes),
            eci_sparse=dict(index=[int(i) for i, _ in fcc_sparse], value=[float(v) for _, v in fcc_sparse]),
            eci_dense=dict(index=fcc_dense_idx.tolist(), value=fcc_dense_val.tolist()),
            eci_2=dict(index=fcc2_idx.tolist(), value=fcc2_val.tolist()),
            # test-only coefficients touching every function (incl. the 2NN pairs
            # the shipped ECI leave at zero and the constant), to exercise the
            # generic evaluator on the FCC basis
            eci_full=dict(index=list(range(9)),
                          value=[0.05, -0.1, 0.3, 0.1, 0.1, 0.5, -0.07, 0.03, 0.11]),
        ),
        "zro": dict(
            tables="zro", n_species=3, species=["Zr", "Va", "O"],
            sublat_to_asym=[0, 0, 1, 1], occ_to_species=[[0, -1], [0, -1], [1, 2], [1, 2]],
            mutable_sublats=[2, 3],
            axes=axes(zro_axes),
            eci=dict(index=zro_idx.tolist(), value=zro_val.tolist()),
        ),
    }


def kmc_system():
    """The FCC A-B-Va KMC system of the reference's tests (kmc_system.json /
    KMCTestSystem.cc:17-53): two event types (A-Va and B-Va nearest-neighbour
    hops), six equivalents each, dense kra/freq coefficients."""
    F = DATA / "FCC_binary_vacancy"
    types = []
    for ev in ("A_Va_1NN", "B_Va_1NN"):
        et = K.read_event_type(F / f"kmc_events/event.{ev}/event.json",
                               F / f"basis_sets/bset.{ev}/equivalents_info.json",
                               F / f"kmc_events/event.{ev}/kra_eci.json",
                               F / f"kmc_events/event.{ev}/freq_eci.json", name=ev)
        ski, skv = read_eci(F / f"kmc_events/event.{ev}/kra_sparse_eci.json")
        sfi, sfv = read_eci(F / f"kmc_events/event.{ev}/freq_sparse_eci.json")
        types.append(dict(name=ev, local_tables=[f"fcc_{ev}_{k}" for k in range(6)],
                          events=[dict(sites=[list(x) for x in e["sites"]], occ_init=e["occ_init"],
                                       occ_final=e["occ_final"]) for e in et["events"]],
                          kra=dict(index=et["kra"][0].tolist(), value=et["kra"][1].tolist()),
                          freq=dict(index=et["freq"][0].tolist(), value=et["freq"][1].tolist()),
                          kra_sparse=dict(index=ski.tolist(), value=skv.tolist()),
                          freq_sparse=dict(index=sfi.tolist(), value=sfv.tolist())))
    return dict(event_types=types)


def kmc_vectors(S, seed=13):
    """Event states from the oracle (reference kernels + the restated
    _default_event_state_calculation): (1) the configuration reproducing the
    event state documented in python/libcasm/clexmonte/_MonteCalculator.py:186-210,
    found by enumerating the occupations compatible with the documented
    local_corr; (2) a random A-B-Va configuration, every prim event at random
    unit cells."""
    import itertools
    sysd = S["fcc"]
    types = sysd["kmc"]["event_types"]
    prim = K.make_prim_event_list(types)
    assert len(prim) == 24
    out = {}
    N = 8
    n = N ** 3
    form = O.RefClexulator("fcc_default").supercell(N)
    local = {(y, k): O.RefClexulator(types[y]["local_tables"][k]).supercell(N)
             for y in range(2) for k in range(6)}

    def site(c):
        return int(c[0] % N + N * (c[1] % N + N * (c[2] % N)))

    # ---- (1) documented event state: B_Va_1NN, equivalent 2, reverse, prim event 17
    pe = prim[17]
    assert (pe["event_type_name"], pe["equivalent_index"], pe["is_forward"]) == ("B_Va_1NN", 2, False)
    assert pe["occ_init"] == [2, 1] and pe["occ_final"] == [1, 2]
    X = np.array([3, 3, 3])
    uc = site(X)
    ls = K.event_linear_site_index((N, N, N), uc, pe["sites"])
    eci2 = sysd["eci_2"]
    g1 = [(0, -1, 1), (0, 2, -2)]
    g3 = [(-1, 0, 0), (1, 0, 0), (-1, 1, -1), (1, 1, -1)]
    g5 = [(-1, 1, 0), (0, 0, -1), (0, 1, 0), (1, 0, -1)]
    g7 = [(-1, 0, 1), (0, -1, 0), (0, 0, 1), (1, -1, 0), (-1, 2, -1), (0, 1, -2), (0, 2, -1), (1, 1, -2)]
    doc = dict(Ekra=0.7375, dE_activated=1.6666666666666665, dE_final=1.6666666666666665,
               freq=1e13, rate=1000704.0785393054, is_normal=False,
               local_corr=[1.0, 0.5, 0.0, 0.5, 0.0, 0.25, 0.0, 0.5, 0.0])
    kat = None
    for a, b, c, d in itertools.product(itertools.combinations(g1, 1), itertools.combinations(g3, 2),
                                        itertools.combinations(g5, 1), itertools.combinations(g7, 4)):
        occ = np.zeros(n, dtype=np.int32)
        occ[ls[0]], occ[ls[1]] = 2, 1
        for s_ in a + b + c + d:
            occ[site(X + np.array(s_))] = 1
        st = O.event_state(form, local[(1, 2)], occ, uc, ls, pe["occ_init"], pe["occ_final"],
                           eci2["index"], eci2["value"],
                           (types[1]["kra"]["index"], types[1]["kra"]["value"]),
                           (types[1]["freq"]["index"], types[1]["freq"]["value"]), 1200.0)
        if all(st[k] == doc[k] for k in ("Ekra", "dE_activated", "dE_final", "freq", "rate", "is_normal")) \
                and (st["local_corr"] == np.array(doc["local_corr"])).all():
            kat = occ
            break
    assert kat is not None, "no configuration reproduces the documented event state"
    out["kat_N"] = np.array([N, N, N])
    out["kat_occ"] = kat.astype(np.int8)
    out["kat_unitcell"] = np.array(uc)
    out["kat_prim_event"] = np.array(17)
    out["kat_T"] = np.array(1200.0)
    for k in ("Ekra", "dE_activated", "dE_final", "freq", "rate"):
        out[f"kat_{k}"] = np.array(doc[k])
    # ---- (2) random configurations (A/B/Va = 60/25/15 %): the dense ECI of the
    # unit tests, and the larger ECI of formation_energy_eci.2.json, which make
    # many events "abnormal" (both clamps of BaseMonteEventData.cc:154-155 fire)
    rng = np.random.default_rng(seed)
    for key, eci, T in (("rand", sysd["eci_dense"], 900.0), ("rand2", sysd["eci_2"], 1200.0)):
        occ = rng.choice(3, size=n, p=[0.6, 0.25, 0.15]).astype(np.int32)
        ucs, pes, rows = [], [], []
        for p_ in range(24):
            pe = prim[p_]
            y, k = pe["event_type"], pe["equivalent_index"]
            # unit cells where the event is allowed, plus a few where it is not
            allowed = [c_ for c_ in range(n)
                       if all(occ[l] == o for l, o in zip(K.event_linear_site_index((N, N, N), c_, pe["sites"]),
                                                          pe["occ_init"]))]
            cells = list(rng.choice(allowed, size=min(30, len(allowed)), replace=False)) + \
                list(rng.integers(0, n, 6))
            for c_ in cells:
                ls = K.event_linear_site_index((N, N, N), int(c_), pe["sites"])
                st = O.event_state(form, local[(y, k)], occ, int(c_), ls, pe["occ_init"], pe["occ_final"],
                                   eci["index"], eci["value"],
                                   (types[y]["kra"]["index"], types[y]["kra"]["value"]),
                                   (types[y]["freq"]["index"], types[y]["freq"]["value"]), T)
                ucs.append(int(c_))
                pes.append(p_)
                rows.append([float(st["is_allowed"]), float(st["is_normal"]), st["dE_final"], st["Ekra"],
                             st["dE_activated"], st["freq"], st["rate"]])
        rows = np.array(rows)
        out[f"{key}_N"] = np.array([N, N, N])
        out[f"{key}_occ"] = occ.astype(np.int8)
        out[f"{key}_T"] = np.array(T)
        out[f"{key}_eci_index"] = np.array(eci["index"], dtype=np.uint32)
        out[f"{key}_eci_value"] = np.array(eci["value"], dtype=np.float64)
        out[f"{key}_unitcell"] = np.array(ucs, dtype=np.int64)
        out[f"{key}_prim_event"] = np.array(pes, dtype=np.int32)
        out[f"{key}_states"] = rows
        print("kmc", key, ": allowed", int(rows[:, 0].sum()), "of", len(rows), "normal", int(rows[:, 1].sum()),
              "clamped to dE_final", int(((rows[:, 4] == rows[:, 2]) & (rows[:, 0] > 0)).sum()),
              "clamped to 0", int(((rows[:, 4] == 0) & (rows[:, 0] > 0)).sum()))
    np.savez_compressed(OUT / "vectors_kmc.npz", **out)


def random_occ(rng, n_cells, n_sublat, mutable, nocc):
    occ = np.zeros(n_cells * n_sublat, dtype=np.int32)
    for b in mutable:
        occ[b * n_cells:(b + 1) * n_cells] = rng.integers(0, nocc, n_cells)
    return occ


def vectors(case, sysd, N, eci, n_events=64, traj_steps=20000, seed=7):
    rng = np.random.default_rng(seed)
    clex = O.RefClexulator(sysd["tables"])
    sc = clex.supercell(N)
    mut = sysd["mutable_sublats"]
    nocc = max(sum(1 for s in row if s >= 0) for row in sysd["occ_to_species"])
    occ = random_occ(rng, sc.n_cells, clex.n_sublat, mut, nocc)
    out = dict(N=np.array(sc.N), occ=occ)
    # single-site
    ls = np.array([mut[rng.integers(len(mut))] * sc.n_cells + rng.integers(sc.n_cells)
                   for _ in range(n_events)], dtype=np.int64)
    new = np.array([(occ[l] + 1 + rng.integers(nocc - 1)) % nocc for l in ls], dtype=np.int32)
    out["l"] = ls
    out["new_occ"] = new
    out["delta_corr"] = np.array([sc.delta_corr(occ, l, n) for l, n in zip(ls, new)])
    out["point_corr"] = np.array([sc.point_corr(occ, l) for l in ls])
    cells = rng.integers(0, sc.n_cells, n_events).astype(np.int64)
    out["cells"] = cells
    out["cell_corr"] = np.array([sc.cell_corr(occ, c) for c in cells])
    out["global_corr"] = sc.global_corr(occ)
    eidx = np.array(eci["index"], dtype=np.uint32)
    eval_ = np.array(eci["value"], dtype=np.float64)
    out["eci_index"], out["eci_value"] = eidx, eval_
    out["delta_e_1"] = np.array([sc.occ_delta_value(occ, [l], [n], eidx, eval_) for l, n in zip(ls, new)])
    # two-site events: neighbouring and distant pairs, sequential semantics
    l2 = []
    n2 = []
    for q in range(n_events):
        la = ls[q]
        if q % 2 == 0:  # a near neighbour: next cell along i on a mutable sublattice
            b2 = mut[rng.integers(len(mut))]
            cell = (la % sc.n_cells)
            i = cell % sc.N[0]
            cell2 = cell - i + (i + 1) % sc.N[0]
            lb = b2 * sc.n_cells + cell2
        else:
            lb = mut[rng.integers(len(mut))] * sc.n_cells + rng.integers(sc.n_cells)
        if lb == la:
            lb = mut[0] * sc.n_cells + (la + 1) % sc.n_cells
        l2.append([la, lb])
        n2.append([new[q], (occ[lb] + 1 + rng.integers(nocc - 1)) % nocc])
    l2 = np.array(l2, dtype=np.int64)
    n2 = np.array(n2, dtype=np.int32)
    out["l2"], out["new_occ2"] = l2, n2
    out["delta_e_2"] = np.array([sc.occ_delta_value(occ, a, b, eidx, eval_) for a, b in zip(l2, n2)])
    # potential + composition
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=sysd["n_species"], Rt=np.array(sysd["axes"]["Rt"]),
                origin=np.array(sysd["axes"]["origin"]))
    mu = np.array([0.3, -0.2][:len(sysd["axes"]["end_members"])])
    out["param_chem_pot"] = mu
    e, comp = sc.potential_per_supercell(occ, prim, eidx, eval_, mu)
    out["potential_per_supercell"] = np.array(e)
    out["mol_composition"] = comp
    # sequential trajectories (std::mt19937_64)
    for mode, name in ((0, "sgc"), (1, "canonical")):
        T = 800.0
        res = sc.metropolis_run(mode, occ, prim, eidx, eval_, T, seed=12345 + mode, n_steps=traj_steps,
                                param_chem_pot=mu if mode == 0 else None, log_cap=256)
        out[f"{name}_T"] = np.array(T)
        out[f"{name}_seed"] = np.array(12345 + mode, dtype=np.uint64)
        out[f"{name}_steps"] = np.array(traj_steps)
        out[f"{name}_final_occ"] = res["occ"].astype(np.int8)
        out[f"{name}_n_accept"] = np.array(res["n_accept"])
        out[f"{name}_hash"] = np.array(res["hash"], dtype=np.uint64)
        log = res["log"]
        out[f"{name}_log_l0"] = np.array([s["l0"] for s in log], dtype=np.int64)
        out[f"{name}_log_l1"] = np.array([s["l1"] for s in log], dtype=np.int64)
        out[f"{name}_log_new0"] = np.array([s["new0"] for s in log], dtype=np.int32)
        out[f"{name}_log_new1"] = np.array([s["new1"] for s in log], dtype=np.int32)
        out[f"{name}_log_acc"] = np.array([s["accepted"] for s in log], dtype=np.int32)
        out[f"{name}_log_dE"] = np.array([s["dE"] for s in log], dtype=np.float64)
    np.savez_compressed(OUT / f"vectors_{case}.npz", **out)
    print(case, "N", sc.N, "sgc acc", out["sgc_n_accept"], "can acc", out["canonical_n_accept"])


def local_vectors(seed=11):
    """Per-cell contributions of the 12 local (KMC) clexulators."""
    rng = np.random.default_rng(seed)
    N = 6
    out = {}
    occ = rng.integers(0, 3, N ** 3).astype(np.int32)
    cells = rng.integers(0, N ** 3, 16).astype(np.int64)
    out["N"] = np.array([N, N, N])
    out["occ"] = occ
    out["cells"] = cells
    for ev in ("A_Va_1NN", "B_Va_1NN"):
        for k in range(6):
            name = f"fcc_{ev}_{k}"
            sc = O.RefClexulator(name).supercell(N)
            out[name] = np.array([sc.cell_corr(occ, c) for c in cells])
    np.savez_compressed(OUT / "vectors_local.npz", **out)


def rng_vectors():
    """libstdc++ draws from std::mt19937_64 (the stream the reference consumes)."""
    rng = np.random.default_rng(3)
    n = 4096
    kinds = rng.integers(0, 3, n).astype(np.int32)
    int_max = rng.integers(0, 10 ** 7, n).astype(np.int64)
    int_max[::7] = 0
    int_max[1::97] = 2 ** 62
    real_max = rng.uniform(0.5, 1e7, n)
    real_max[::5] = 1.0
    oi, orl, oraw = O.rng_stream(20261017, kinds, int_max, real_max)
    np.savez_compressed(OUT / "vectors_rng.npz", seed=np.array(20261017, dtype=np.uint64), kinds=kinds,
                        int_max=int_max, real_max=real_max, out_int=oi, out_real=orl, out_raw=oraw)


def synthetic_system():
    """The synthetic FCC binary A-B pair + triplet basis (make_synthetic_clexulator.py, SURVEY 8c
    "Gap" / BASELINE configs[0]): composition axis from pure A to pure B, test coefficients on
    every function."""
    origin, ends = [1.0, 0.0], [[0.0, 1.0]]
    return dict(tables="fcc_synthetic", n_species=2, species=["A", "B"], sublat_to_asym=[0],
                occ_to_species=[[0, 1]], mutable_sublats=[0],
                axes=dict(components=["A", "B"], origin=origin, end_members=ends, Rt=rt_matrix(origin, ends).tolist()),
                eci=dict(index=[0, 1, 2, 3, 4], value=[-0.02, 0.01, 0.035, -0.012, 0.008]))


def synthetic():
    """Only the synthetic-basis fixtures (the reference-derived ones stay untouched):
    python tests/golden/make_golden.py synthetic"""
    O.build()
    from make_synthetic_clexulator import NAME
    t = parse_clexulator_source(ROOT / "oracle/_ref" / f"{NAME}.cc", name="fcc_synthetic")
    t.save(OUT / "tables" / "fcc_synthetic.npz")
    S = json.loads((OUT / "systems.json").read_text())
    S["fcc_syn"] = synthetic_system()
    (OUT / "systems.json").write_text(json.dumps(S, indent=1))
    vectors("fcc_synthetic", S["fcc_syn"], 6, S["fcc_syn"]["eci"], seed=10)


def main():
    if sys.argv[1:] == ["synthetic"]:
        return synthetic()
    O.build()
    (OUT / "tables").mkdir(exist_ok=True)
    for name, src in SOURCES.items():
        t = parse_clexulator_source(src, name=name)
        t.save(OUT / "tables" / f"{name}.npz")
    S = systems()
    S["fcc"]["kmc"] = kmc_system()
    (OUT / "systems.json").write_text(json.dumps(S, indent=1))
    kmc_vectors(S)
    vectors("fcc_sparse", S["fcc"], 6, S["fcc"]["eci_sparse"])
    vectors("fcc_full", S["fcc"], 6, S["fcc"]["eci_full"], seed=8)
    vectors("zro", S["zro"], 8, S["zro"]["eci"], n_events=32, traj_steps=5000, seed=9)
    synthetic()
    local_vectors()
    rng_vectors()


if __name__ == "__main__":
    main()
