"""GPU parity of the KMC event-state kernel (csrc/cmx_kmc.cu) through the C ABI.

Tolerances: is_allowed / is_normal / dE_final / Ekra / dE_activated / freq are
BIT-EXACT against the reference kernels (faithful evaluation order); rate goes
through exp(), whose device and glibc implementations may differ in the last
bit: rtol 4e-16 * |beta dE| + 1e-15, checked as rel 1e-13."""
import numpy as np
import pytest

from conftest import GOLDEN

from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200 import kmc as K

pytestmark = pytest.mark.gpu


def _types(systems, sparse=False):
    out = []
    for et in systems["fcc"]["kmc"]["event_types"]:
        kra = et["kra_sparse" if sparse else "kra"]
        freq = et["freq_sparse" if sparse else "freq"]
        out.append(dict(et, kra=(kra["index"], kra["value"]), freq=(freq["index"], freq["value"])))
    return out


@pytest.fixture(scope="module")
def kmc_tables(load_tables):
    t = {name: _capi.Tables(load_tables(name)) for name in
         ["fcc_default"] + [f"fcc_{ev}_{k}" for ev in ("A_Va_1NN", "B_Va_1NN") for k in range(6)]}
    yield t
    for x in t.values():
        x.close()


def _kmc(kmc_tables, systems, N, occ, eci_index, eci_value, T, n_replicas=1, sparse=False):
    types = _types(systems, sparse)
    prim = K.make_prim_event_list(types)
    st = _capi.State(kmc_tables["fcc_default"], N, n_replicas)
    for r in range(n_replicas):
        st.upload_occ(occ if np.ndim(occ) == 1 else occ[r], r)
        st.set_conditions(T if np.isscalar(T) else T[r], None, r)
    st.set_eci(eci_index, eci_value)
    dev_types = [dict(local_tables=[kmc_tables[n] for n in et["local_tables"]], kra=et["kra"], freq=et["freq"])
                 for et in types]
    return st, _capi.Kmc(st, dev_types, prim), prim


def test_documented_event_state(kmc_tables, systems):
    """python/libcasm/clexmonte/_MonteCalculator.py:186-210."""
    v = dict(np.load(GOLDEN / "vectors_kmc.npz"))
    eci = systems["fcc"]["eci_2"]
    st, kmc, prim = _kmc(kmc_tables, systems, tuple(int(x) for x in v["kat_N"]), v["kat_occ"],
                         eci["index"], eci["value"], float(v["kat_T"]))
    s = kmc.event_states([int(v["kat_unitcell"])], [int(v["kat_prim_event"])])[0]
    assert s["is_allowed"] == 1 and s["is_normal"] == 0
    assert s["dE_final"] == 1.6666666666666665
    assert s["Ekra"] == 0.7375
    assert s["dE_activated"] == 1.6666666666666665
    assert s["freq"] == 1e13
    assert s["rate"] == pytest.approx(1000704.0785393054, rel=1e-13)
    kmc.close()
    st.close()


@pytest.mark.parametrize("key", ["rand", "rand2"])
def test_event_states_match_golden(kmc_tables, systems, key):
    v = dict(np.load(GOLDEN / "vectors_kmc.npz"))
    st, kmc, prim = _kmc(kmc_tables, systems, tuple(int(x) for x in v[f"{key}_N"]), v[f"{key}_occ"],
                         v[f"{key}_eci_index"], v[f"{key}_eci_value"], float(v[f"{key}_T"]))
    s = kmc.event_states(v[f"{key}_unitcell"], v[f"{key}_prim_event"])
    ref = v[f"{key}_states"]
    assert (s["is_allowed"] == ref[:, 0]).all()
    assert (s["is_normal"] == ref[:, 1]).all()
    for c, name in ((2, "dE_final"), (3, "Ekra"), (4, "dE_activated"), (5, "freq")):
        assert (s[name] == ref[:, c]).all(), name
    np.testing.assert_allclose(s["rate"], ref[:, 6], rtol=1e-13, atol=0)
    assert (s["rate"][ref[:, 0] == 0] == 0).all()
    kmc.close()
    st.close()


@pytest.mark.parametrize("sparse", [False, True])
def test_reference_event_state_test(kmc_tables, systems, sparse):
    """events_EventStateCalculator_test.cpp:25-88 and
    events_CompleteEventCalculator_test.cpp:60-102: all A + one vacancy, T = 600 K,
    complete event list: 12 allowed events, dE_final ~ 0, Ekra ~ 1, freq ~ 1e12,
    rate ~ 1e12 exp(-beta); all other rates are exactly 0."""
    N = (8, 8, 8)
    n = 512
    occ = np.zeros(n, dtype=np.int32)
    occ[0] = 2
    eci = systems["fcc"]["eci_sparse" if sparse else "eci_dense"]
    st, kmc, prim = _kmc(kmc_tables, systems, N, occ, eci["index"], eci["value"], 600.0, sparse=sparse)
    uc, pe = K.complete_event_list(n, len(prim))
    assert len(uc) == 24 * n
    s = kmc.event_states(uc, pe)
    assert int(s["is_allowed"].sum()) == 12
    a = s[s["is_allowed"] == 1]
    beta = 1.0 / (8.6173303e-05 * 600.0)
    np.testing.assert_allclose(a["dE_final"], 0.0, atol=1e-5)
    np.testing.assert_allclose(a["Ekra"], 1.0, atol=1e-5)
    np.testing.assert_allclose(a["dE_activated"], 1.0, atol=1e-5)
    np.testing.assert_allclose(a["freq"], 1e12, rtol=1e-12)
    np.testing.assert_allclose(a["rate"], 1e12 * np.exp(-beta), rtol=1e-5)
    # the selector's initial pass over all events gives the same numbers
    rates, total = kmc.all_rates()
    assert (rates[0].reshape(-1) == s["rate"]).all()
    assert total[0] == pytest.approx(s["rate"].sum(), rel=1e-12)
    kmc.close()
    st.close()


def test_batched_trajectories(kmc_tables, systems, oracle):
    """Many independent trajectories = replicas of the state: every (replica,
    cell, event) triple is evaluated against its own configuration and
    temperature; checked against the live oracle on a sample."""
    rng = np.random.default_rng(5)
    N = (6, 6, 6)
    n = 216
    R = 5
    occ = rng.choice(3, size=(R, n), p=[0.6, 0.25, 0.15]).astype(np.int32)
    T = [600.0 + 150.0 * r for r in range(R)]
    eci = systems["fcc"]["eci_dense"]
    st, kmc, prim = _kmc(kmc_tables, systems, N, occ, eci["index"], eci["value"], T, n_replicas=R)
    rates, total = kmc.all_rates()
    assert rates.shape == (R, n, 24)
    rep = rng.integers(0, R, 300).astype(np.int32)
    uc = rng.integers(0, n, 300)
    pe = rng.integers(0, 24, 300).astype(np.int32)
    s = kmc.event_states(uc, pe, rep)
    assert (s["rate"] == rates[rep, uc, pe]).all()
    np.testing.assert_allclose(total, rates.reshape(R, -1).sum(axis=1), rtol=1e-12)
    if oracle is not None:
        types = _types(systems)
        form = oracle.RefClexulator("fcc_default").supercell(N)
        checked = 0
        for q in np.nonzero(s["is_allowed"])[0][:40]:
            p = prim[int(pe[q])]
            y, k = p["event_type"], p["equivalent_index"]
            loc = oracle.RefClexulator(types[y]["local_tables"][k]).supercell(N)
            o = oracle.event_state(form, loc, occ[rep[q]], int(uc[q]),
                                   K.event_linear_site_index(N, int(uc[q]), p["sites"]), p["occ_init"],
                                   p["occ_final"], eci["index"], eci["value"], types[y]["kra"], types[y]["freq"],
                                   T[rep[q]])
            assert o["is_allowed"]
            for name in ("dE_final", "Ekra", "dE_activated", "freq"):
                assert s[name][q] == o[name], name
            assert s["rate"][q] == pytest.approx(o["rate"], rel=1e-13)
            checked += 1
        assert checked > 10
    kmc.close()
    st.close()


def test_kmc_error_paths(kmc_tables, systems):
    types = _types(systems)
    prim = K.make_prim_event_list(types)
    dev_types = [dict(local_tables=[kmc_tables[n] for n in et["local_tables"]], kra=et["kra"], freq=et["freq"])
                 for et in types]
    st = _capi.State(kmc_tables["fcc_default"], (8, 8, 8))
    with pytest.raises(_capi.CmxError):
        _capi.Kmc(st, dev_types, prim)              # no formation energy ECI bound
    st.set_eci([1], [0.5])
    kmc = _capi.Kmc(st, dev_types, prim)
    with pytest.raises(_capi.CmxError):
        kmc.event_states([0], [0])                   # temperature not set
    st.set_conditions(500.0)
    with pytest.raises(_capi.CmxError):
        kmc.event_states([512], [0])
    with pytest.raises(_capi.CmxError):
        kmc.event_states([0], [24])
    bad = [dict(p) for p in prim]
    bad[0] = dict(bad[0], occ_init=[0, 5])
    with pytest.raises(_capi.CmxError):
        _capi.Kmc(st, dev_types, bad)
    kmc.close()
    st.close()


def _impacted_ids(N, n_prim, beg, ent):
    """Absolute impact lists of the complete event list from the relative table."""
    N0, N1, N2 = N
    n_cells = N0 * N1 * N2
    out = []
    for c in range(n_cells):
        i, j, k = c % N0, (c // N0) % N1, c // (N0 * N1)
        for pe in range(n_prim):
            rows = ent[beg[pe]:beg[pe + 1]]
            cells = ((i + rows[:, 1]) % N0) + N0 * (((j + rows[:, 2]) % N1) + N1 * ((k + rows[:, 3]) % N2))
            out.append(np.unique(cells.astype(np.int64) * n_prim + rows[:, 0]))
    return out


def test_kmc_run_matches_reference_selector(kmc_tables, systems, oracle):
    """whole KMC steps on the device reproduce the event sequence of the
    reference's own selector (lotto::RejectionFreeEventSelector, compiled unmodified into
    oracle/_ref/libkmc_lotto.so, seeded std::mt19937_64) fed with event rates from the
    reference's generated kernels (harness.cpp orc_kmc_rate) on the same initial
    occupation: identical (unit cell, prim event) at every one of 500 steps of three
    trajectories, identical final occupation; simulated time to 1e-11 (device exp/log vs
    glibc in the last bit of individual rates)."""
    if oracle is None or not oracle.lotto_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    N = (6, 6, 6)
    n = 216
    R, n_steps = 3, 500
    occ = rng.choice(3, size=(R, n), p=[0.55, 0.35, 0.10]).astype(np.int32)
    occ[2] = rng.choice(3, size=n, p=[0.88, 0.1, 0.02])
    T = [1200.0, 900.0, 600.0]
    seeds = [11, 2026, 5]
    eci = systems["fcc"]["eci_dense"]
    st, kmc, prim = _kmc(kmc_tables, systems, N, occ, eci["index"], eci["value"], T, n_replicas=R)
    kmc.run_begin(seeds)
    out = kmc.run(n_steps, log_cap=n_steps)
    assert (out["n_steps"] == n_steps).all()
    beg, ent = kmc._impact
    impacted = _impacted_ids(N, len(prim), beg, ent)
    types = _types(systems)
    for r in range(R):
        ref = oracle.KmcReference(N, occ[r], prim, types, eci["index"], eci["value"], T[r], impacted, seeds[r])
        t, ev, _ = ref.run(n_steps)
        gpu_ev = out["log"]["unitcell"][r] * len(prim) + out["log"]["prim_event"][r]
        mism = np.nonzero(gpu_ev != ev)[0]
        assert len(mism) == 0, f"replica {r}: first divergence at step {mism[0]}"
        assert out["time"][r] == pytest.approx(t, rel=1e-11)
        assert out["log"]["time_increment"][r].sum() == pytest.approx(t, rel=1e-11)
        assert (st.download_occ(r) == ref.occ).all()
    # the Python-callback form of the same oracle (oracle.event_state) agrees on a short run
    cur = occ[0].copy()
    form = oracle.RefClexulator("fcc_default").supercell(N)
    loc = {name: oracle.RefClexulator(name).supercell(N) for et in types for name in et["local_tables"]}

    def rate(e):
        cell, pe = divmod(e, len(prim))
        p = prim[pe]
        y, k = p["event_type"], p["equivalent_index"]
        s = oracle.event_state(form, loc[types[y]["local_tables"][k]], cur, cell,
                               K.event_linear_site_index(N, cell, p["sites"]), p["occ_init"], p["occ_final"],
                               eci["index"], eci["value"], types[y]["kra"], types[y]["freq"], T[0])
        return s["rate"] if s["is_allowed"] else 0.0

    sel = oracle.LottoSelector(n * len(prim), rate, impacted, seeds[0])
    for step in range(12):
        e, dt, tot = sel.select()
        g = out["log"][0, step]
        assert (int(g["unitcell"]), int(g["prim_event"])) == divmod(e, len(prim))
        assert g["time_increment"] == pytest.approx(dt, rel=1e-12)
        assert g["total_rate"] == pytest.approx(tot, rel=1e-12)
        cell, pe = divmod(e, len(prim))
        for l, o in zip(K.event_linear_site_index(N, cell, prim[pe]["sites"]), prim[pe]["occ_final"]):
            cur[l] = o
    kmc.close()
    st.close()


def test_kmc_run_bookkeeping(kmc_tables, systems):
    """Size-independent properties of the device KMC: after any number of steps the
    selector's leaves equal the rates recomputed from scratch for the current occupation
    (the impact table misses nothing), the roots equal lotto's pairwise tree sums of the
    leaves bit for bit, species are conserved, time grows, a run continues where the
    previous call stopped, and a configuration without events stops at once."""
    rng = np.random.default_rng(8)
    N = (8, 6, 6)
    n = int(np.prod(N))
    R = 6
    occ = rng.choice(3, size=(R, n), p=[0.7, 0.25, 0.05]).astype(np.int32)
    occ[R - 1] = rng.choice(2, size=n)          # no vacancy: no allowed event
    eci = systems["fcc"]["eci_dense"]
    T = [800.0 + 100.0 * r for r in range(R)]
    st, kmc, prim = _kmc(kmc_tables, systems, N, occ, eci["index"], eci["value"], T, n_replicas=R)
    st2, kmc2, _ = _kmc(kmc_tables, systems, N, occ, eci["index"], eci["value"], T, n_replicas=R)
    seeds = np.arange(R) + 5
    kmc.run_begin(seeds)
    kmc2.run_begin(seeds)
    a = kmc.run(300, log_cap=300)
    b1 = kmc2.run(120, log_cap=300)
    b2 = kmc2.run(180, log_cap=300)
    assert (a["n_steps"][:-1] == 300).all() and a["n_steps"][-1] == 0 and a["time"][-1] == 0.0
    assert (b2["n_steps"] == a["n_steps"]).all()
    for name in ("unitcell", "prim_event", "time_increment", "total_rate"):
        # a call logs only the steps it made: the first call's 120 + the second call's 180
        assert (a["log"][name][:, :120] == b1["log"][name][:, :120]).all(), name
        assert (a["log"][name][:, 120:] == b2["log"][name][:, 120:]).all(), name
    assert (a["time"] == b2["time"]).all() and (a["time"][:-1] > 0).all()
    cur, tot = kmc.current_rates()
    fresh, _ = kmc.all_rates()
    assert (cur == fresh).all()
    for r in range(R):
        lvl = cur[r].reshape(-1)
        while len(lvl) > 1:                       # InvertedBinarySumTree: pairwise, odd one carried up
            m = len(lvl) // 2
            nxt = lvl[0:2 * m:2] + lvl[1:2 * m:2]
            lvl = np.concatenate([nxt, lvl[2 * m:]])
        assert lvl[0] == tot[r]
        assert (np.bincount(st.download_occ(r), minlength=3) == np.bincount(occ[r], minlength=3)).all()
        assert (st.download_occ(r) == st2.download_occ(r)).all()
    assert (st.download_occ(0) != occ[0]).any()
    with pytest.raises(_capi.CmxError):
        _kmc(kmc_tables, systems, N, occ, eci["index"], eci["value"], T, n_replicas=R)[1].run(1)   # no run_begin
    for x in (kmc, kmc2, st, st2):
        x.close()
