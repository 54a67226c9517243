"""CPU tests of the KMC rate path: the oracle (reference kernels + restated
_default_event_state_calculation) against the reference's own known answers, and
the host-side prim event list."""
import numpy as np
import pytest

from conftest import GOLDEN

from casmcode_clexmonte_b200 import kmc as K


def _types(systems, sparse=False):
    out = []
    for et in systems["fcc"]["kmc"]["event_types"]:
        kra = et["kra_sparse" if sparse else "kra"]
        freq = et["freq_sparse" if sparse else "freq"]
        out.append(dict(et, kra=(kra["index"], kra["value"]), freq=(freq["index"], freq["value"])))
    return out


def test_prim_event_list_matches_reference_counts(systems):
    """events_System_impact_table_test.cpp:43-47 / events_EventStateCalculator_test.cpp:52:
    the FCC A-B-Va system has 24 prim events (2 types x 6 equivalents x 2 directions);
    the documented event (python/libcasm/clexmonte/_MonteCalculator.py:200-208) is
    prim event 17 = B_Va_1NN, equivalent 2, reverse, occ [2,1] -> [1,2]."""
    prim = K.make_prim_event_list(_types(systems))
    assert len(prim) == 24
    assert [p["prim_event_index"] for p in prim] == list(range(24))
    p17 = prim[17]
    assert (p17["event_type_name"], p17["equivalent_index"], p17["is_forward"]) == ("B_Va_1NN", 2, False)
    assert p17["occ_init"] == [2, 1] and p17["occ_final"] == [1, 2]
    # every hop connects nearest neighbours: the 12 directed NN vectors appear once per type/direction
    for y in range(2):
        vecs = set()
        for p in prim:
            if p["event_type"] == y and p["is_forward"]:
                a, b = np.array(p["sites"][0][1:]), np.array(p["sites"][1][1:])
                vecs.add(tuple(b - a))
        assert len(vecs) == 6


def test_event_linear_site_index_wraps():
    assert K.event_linear_site_index((4, 4, 4), 0, [(0, 0, 0, 0), (0, -1, 0, 1)]) == [0, 3 + 16]
    uc, pe = K.complete_event_list(3, 2)
    assert uc.tolist() == [0, 0, 1, 1, 2, 2] and pe.tolist() == [0, 1, 0, 1, 0, 1]


def test_oracle_reproduces_documented_event_state(oracle, systems):
    """python/libcasm/clexmonte/_MonteCalculator.py:186-210 (T = 1200 K)."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    v = dict(np.load(GOLDEN / "vectors_kmc.npz"))
    types = _types(systems)
    prim = K.make_prim_event_list(types)
    N = tuple(int(x) for x in v["kat_N"])
    pe = prim[int(v["kat_prim_event"])]
    uc = int(v["kat_unitcell"])
    form = oracle.RefClexulator("fcc_default").supercell(N)
    loc = oracle.RefClexulator(types[1]["local_tables"][2]).supercell(N)
    eci = systems["fcc"]["eci_2"]
    st = oracle.event_state(form, loc, v["kat_occ"].astype(np.int32), uc,
                            K.event_linear_site_index(N, uc, pe["sites"]), pe["occ_init"], pe["occ_final"],
                            eci["index"], eci["value"], types[1]["kra"], types[1]["freq"], float(v["kat_T"]))
    assert st["is_allowed"] and not st["is_normal"]
    assert st["dE_final"] == 1.6666666666666665
    assert st["Ekra"] == 0.7375
    assert st["dE_activated"] == 1.6666666666666665
    assert st["freq"] == 1e13
    assert st["rate"] == 1000704.0785393054          # bit for bit: pins KB and the clamp order
    assert (st["local_corr"] == [1.0, 0.5, 0.0, 0.5, 0.0, 0.25, 0.0, 0.5, 0.0]).all()


@pytest.mark.parametrize("sparse", [False, True])
def test_oracle_reproduces_reference_event_state_test(oracle, systems, sparse):
    """events_EventStateCalculator_test.cpp:25-88: all A, one vacancy at site 0,
    T = 600 K: 12 of the events are allowed, each with dE_final ~ 0, Ekra ~ 1.0,
    freq ~ 1e12, rate ~ 1e12 exp(-beta).  (The reference uses the 10x10x10
    conventional FCC supercell; the values do not depend on the supercell as long
    as the vacancy does not see its own images -- a 8^3 primitive box here.)"""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    N = (8, 8, 8)
    n = 512
    types = _types(systems, sparse)
    prim = K.make_prim_event_list(types)
    eci = systems["fcc"]["eci_sparse" if sparse else "eci_dense"]
    occ = np.zeros(n, dtype=np.int32)
    occ[0] = 2
    form = oracle.RefClexulator("fcc_default").supercell(N)
    local = {(y, k): oracle.RefClexulator(types[y]["local_tables"][k]).supercell(N)
             for y in range(2) for k in range(6)}
    beta = 1.0 / (oracle.KB * 600.0)
    n_allowed = 0
    ucs, pes = K.complete_event_list(n, len(prim))
    assert len(ucs) == 24 * n
    for uc, p in zip(ucs, pes):
        pe = prim[p]
        ls = K.event_linear_site_index(N, int(uc), pe["sites"])
        if 0 not in ls:
            continue  # cannot be allowed: no vacancy among the sites
        y, k = pe["event_type"], pe["equivalent_index"]
        st = oracle.event_state(form, local[(y, k)], occ, int(uc), ls, pe["occ_init"], pe["occ_final"],
                                eci["index"], eci["value"], types[y]["kra"], types[y]["freq"], 600.0)
        if st["is_allowed"]:
            n_allowed += 1
            assert st["dE_final"] == pytest.approx(0.0, abs=1e-5)
            assert st["Ekra"] == pytest.approx(1.0, abs=1e-5)
            assert st["dE_activated"] == pytest.approx(1.0, abs=1e-5)
            assert st["freq"] == pytest.approx(1e12, rel=1e-12)
            assert st["rate"] == pytest.approx(1e12 * np.exp(-beta * 1.0), rel=1e-5)
    assert n_allowed == 12


def test_oracle_reproduces_kmc_golden(oracle, systems):
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    v = dict(np.load(GOLDEN / "vectors_kmc.npz"))
    types = _types(systems)
    prim = K.make_prim_event_list(types)
    N = tuple(int(x) for x in v["rand2_N"])
    form = oracle.RefClexulator("fcc_default").supercell(N)
    occ = v["rand2_occ"].astype(np.int32)
    for q in range(0, len(v["rand2_unitcell"]), 7):
        uc, p = int(v["rand2_unitcell"][q]), int(v["rand2_prim_event"][q])
        pe = prim[p]
        y, k = pe["event_type"], pe["equivalent_index"]
        loc = oracle.RefClexulator(types[y]["local_tables"][k]).supercell(N)
        st = oracle.event_state(form, loc, occ, uc, K.event_linear_site_index(N, uc, pe["sites"]),
                                pe["occ_init"], pe["occ_final"], v["rand2_eci_index"], v["rand2_eci_value"],
                                types[y]["kra"], types[y]["freq"], float(v["rand2_T"]))
        row = v["rand2_states"][q]
        assert [float(st["is_allowed"]), float(st["is_normal"]), st["dE_final"], st["Ekra"],
                st["dE_activated"], st["freq"], st["rate"]] == row.tolist()


def test_relative_impact_table_matches_reference_test(systems, load_tables):
    """tests/unit/clexmonte/events_System_impact_table_test.cpp:43-69: for the FCC A-B-Va
    system with its shipped formation-energy / kra / freq coefficients there are 24 prim
    events, every required update neighborhood has 20 sites and every event impacts 708
    events -- reproduced from the exported tables (kmc.required_update_neighborhood,
    kmc.make_relative_impact_table)."""
    from casmcode_clexmonte_b200 import kmc as K
    sysd = systems["fcc"]
    ft = load_tables("fcc_default")
    types = [dict(et, kra=(et["kra"]["index"], et["kra"]["value"]), freq=(et["freq"]["index"], et["freq"]["value"]))
             for et in sysd["kmc"]["event_types"]]
    prim = K.make_prim_event_list(types)
    assert len(prim) == 24
    nbh = []
    for p in prim:
        et = types[p["event_type"]]
        coef = sorted(set(et["kra"][0]) | set(et["freq"][0]))
        nbh.append(K.required_update_neighborhood(ft, sysd["eci_dense"]["index"],
                                                  load_tables(et["local_tables"][p["equivalent_index"]]), coef,
                                                  p["sites"]))
    assert [len(x) for x in nbh] == [20] * 24
    beg, ent = K.make_relative_impact_table(prim, nbh)
    assert (np.diff(beg) == 708).all() and ent.shape == (24 * 708, 4)
    # an event always impacts itself and its reverse
    for j, p in enumerate(prim):
        rows = {tuple(r) for r in ent[beg[j]:beg[j + 1]]}
        assert (j, 0, 0, 0) in rows


def test_lotto_selector_oracle():
    """oracle/_ref/libkmc_lotto.so = the reference's lotto::RejectionFreeEventSelector,
    compiled unmodified: with constant rates the selection frequencies follow the rates,
    the time steps are exponential with mean 1 / total, and an impacted rate is
    re-evaluated before the next selection."""
    from oracle import oracle as O
    if not O.lotto_available():
        pytest.skip("oracle/_ref/libkmc_lotto.so not built")
    rates = [1.0, 2.0, 0.0, 5.0, 2.0]
    calls = []

    def rate(e):
        calls.append(e)
        return rates[e]

    sel = O.LottoSelector(5, rate, [[1], [], [], [0, 4], []], seed=7)
    assert calls == [0, 1, 2, 3, 4]
    n = 4000
    hits, dts = np.zeros(5), []
    for _ in range(n):
        before = len(calls)
        e, dt, tot = sel.select()
        assert tot == 10.0 and dt > 0
        hits[e] += 1
        dts.append(dt)
        last = e
    assert hits[2] == 0
    np.testing.assert_allclose(hits / n, np.array(rates) / 10.0, atol=0.03)
    assert np.mean(dts) == pytest.approx(0.1, rel=0.06)
    # the impact list of the last selected event is evaluated at the start of the next select
    before = len(calls)
    sel.select()
    assert calls[before:] == [[1], [], [], [0, 4], []][last]


def test_relative_impact_table_point_only(systems, load_tables):
    """tests/unit/clexmonte/events_System_impact_table_test.cpp:75-128 (Test2): only the
    A-Va events (12 prim events), point-function formation-energy ECI {0, 1}, constant-only
    kra / freq coefficients: the required update neighborhood is the 2 event sites and every
    event impacts 46 events."""
    from casmcode_clexmonte_b200 import kmc as K
    sysd = systems["fcc"]
    ft = load_tables("fcc_default")
    et = sysd["kmc"]["event_types"][0]
    assert et["name"] == "A_Va_1NN"
    types = [dict(et, kra=([0], [0.0]), freq=([0], [0.0]))]
    prim = K.make_prim_event_list(types)
    assert len(prim) == 12
    nbh = [K.required_update_neighborhood(ft, [0, 1], load_tables(et["local_tables"][p["equivalent_index"]]), [0],
                                          p["sites"]) for p in prim]
    assert [len(x) for x in nbh] == [2] * 12
    beg, ent = K.make_relative_impact_table(prim, nbh)
    assert (np.diff(beg) == 46).all()


def test_kinetic_params_are_accepted_like_the_reference():
    """KineticCalculator.cc:78-100 (keys), :707-830 (values and defaults)."""
    p = K.parse_kinetic_params(None)
    assert (p["event_data_type"], p["event_selector_type"], p["impact_table_type"]) == \
        ("default", "vector_sum_tree", "neighborlist")
    assert p["assign_allowed_events_only"] is True
    for sel in ("vector_sum_tree", "sum_tree", "direct_sum"):
        for imp in ("neighborlist", "relative"):
            for dat in ("high_memory", "default", "low_memory"):
                q = K.parse_kinetic_params({"event_selector_type": sel, "impact_table_type": imp,
                                            "event_data_type": dat})
                assert q["device_equivalent"]["impact_table_type"] == "relative"
    with pytest.raises(ValueError, match="event_selector_type"):
        K.parse_kinetic_params({"event_selector_type": "heap"})
    with pytest.raises(ValueError, match="impact_table_type"):
        K.parse_kinetic_params({"impact_table_type": "absolute"})
    with pytest.raises(ValueError, match="unrecognized"):
        K.parse_kinetic_params({"event_selector": "sum_tree"})


def test_allowed_event_list_order():
    """AllowedEventList.cc:70-98: allowed events in complete-list (unit cell major) order."""
    flags = np.zeros((5, 4), dtype=np.int32)
    flags[0, 3] = flags[2, 1] = flags[2, 2] = flags[4, 0] = 1
    uc, pe = K.allowed_event_list(flags, 4)
    assert uc.tolist() == [0, 2, 2, 4] and pe.tolist() == [3, 1, 2, 0]
