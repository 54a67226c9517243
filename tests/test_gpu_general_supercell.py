"""General supercells (transformation matrices that are not diag(N0, N1, N2)), -m gpu.

The reference runs any integer transformation matrix (StateData::transformation_matrix_to_super,
include/casm/clexmonte/state/Configuration.hh; its KMC known-answer tests use
10 * fcc_conventional, tests/unit/clexmonte/events_CompleteEventCalculator_test.cpp:30-34).
The library stores the Hermite-normal-form box with skewed periodic images.  Checked here:

  * a skewed box against its TILING: 4 * fcc_conventional (256 cells) is periodic under
    diag(8, 8, 8) (512 cells = two copies), so every faithful evaluator must give, cell for
    cell, BIT-identical results on both boxes, and intensive sums must agree to 1e-13;
  * the reference's KMC known answers on the box its tests use;
  * the energy book of checkerboard sweeps (sum of accepted dE == E(after) - E(before)),
    which fails if two sites of one colour interact through a skewed image;
  * the caller-site-order mapping of upload / download.
"""
import numpy as np
import pytest

from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200 import kmc as K
from casmcode_clexmonte_b200.potential import (mol_composition, semigrand_exchange_table,
                                               semigrand_potential_per_supercell)

pytestmark = pytest.mark.gpu

CONV = np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1]])


@pytest.fixture(scope="module")
def dev_tables(load_tables):
    cache = {}

    def _get(name):
        if name not in cache:
            cache[name] = _capi.Tables(load_tables(name))
        return cache[name]

    yield _get
    for t in cache.values():
        t.close()


def _box_coords(N):
    k, j, i = np.meshgrid(np.arange(N[2]), np.arange(N[1]), np.arange(N[0]), indexing="ij")
    return np.stack([i.reshape(-1), j.reshape(-1), k.reshape(-1)], axis=1)   # row c = cell c


def test_diagonal_matrix_is_the_diag_state(dev_tables, systems):
    sysd = systems["fcc"]
    eci = sysd["eci_sparse"]
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.2, -0.1], sysd["n_species"])
    a = _capi.State(dev_tables(sysd["tables"]), (16, 8, 12))
    b = _capi.State(dev_tables(sysd["tables"]), None, transformation_matrix=np.diag([16, 8, 12]))
    assert b.N == (16, 8, 12) and b.skew == (0, 0, 0)
    for st in (a, b):
        st.set_eci(eci["index"], eci["value"])
        st.set_conditions(900.0, ex)
        st.randomize(4)
    assert a.sweep_info()["evaluator"] == b.sweep_info()["evaluator"] == "pair_lut"
    ca, cb = a.sgc_sweep(4, seed=3), b.sgc_sweep(4, seed=3)
    assert ca[0].n_accept == cb[0].n_accept
    assert (a.download_occ() == b.download_occ()).all()
    assert (a.global_corr() == b.global_corr()).all()
    a.close()
    b.close()


@pytest.mark.parametrize("case_sys,eci_key,T", [
    ("fcc", "eci_full", CONV * 4),
    ("fcc", "eci_sparse", [[4, 0, 0], [4, 8, 0], [0, 0, 8]]),
    ("zro", "eci", [[4, 0, 0], [0, 4, 0], [4, 4, 8]]),
])
def test_skewed_box_equals_its_tiling(dev_tables, systems, case_sys, eci_key, T):
    U = 8 * np.linalg.inv(np.asarray(T, dtype=float))
    assert np.allclose(U, np.round(U), atol=1e-9), "diag(8, 8, 8) must be a superlattice of T"
    sysd = systems[case_sys]
    tab = dev_tables(sysd["tables"])
    eci = sysd[eci_key]
    nsub = tab.host.n_sublat
    sk = _capi.State(tab, None, transformation_matrix=T)
    dg = _capi.State(tab, (8, 8, 8))
    assert sk.skew != (0, 0, 0) and 512 % sk.n_cells == 0 and sk.n_cells < 512
    rng = np.random.default_rng(11)
    nocc = np.array(tab.host.n_occ)
    occ_s = np.concatenate([rng.integers(0, nocc[b], sk.n_cells) for b in range(nsub)]).astype(np.int32)
    # every cell of the diag box is an image of one cell of the skewed box
    img = sk.cell_index(_box_coords(dg.N))
    assert sorted(np.bincount(img, minlength=sk.n_cells)) == [512 // sk.n_cells] * sk.n_cells
    occ_d = np.concatenate([occ_s[b * sk.n_cells + img] for b in range(nsub)]).astype(np.int32)
    for st, occ in ((sk, occ_s), (dg, occ_d)):
        st.upload_occ(occ)
        st.set_eci(eci["index"], eci["value"])
    # the same physical cells on both boxes: box coordinates of the skewed box's cells
    ijk = _box_coords(sk.N)
    cells_d = dg.cell_index(ijk)
    cells_s = np.arange(sk.n_cells)
    assert (sk.cell_index(ijk) == cells_s).all()
    assert (sk.cell_corr(cells_s) == dg.cell_corr(cells_d)).all()
    for b in range(nsub):
        if nocc[b] < 2:
            continue
        new = ((occ_s[b * sk.n_cells + cells_s] + 1 + rng.integers(0, nocc[b] - 1, sk.n_cells)) % nocc[b]).astype(np.int32)
        ls, ld = b * sk.n_cells + cells_s, b * dg.n_cells + cells_d
        assert (sk.point_corr(ls) == dg.point_corr(ld)).all()
        assert (sk.delta_corr(ls, new) == dg.delta_corr(ld, new)).all()
        assert (sk.delta_e(ls, new) == dg.delta_e(ld, new)).all()
    np.testing.assert_allclose(sk.global_corr() / sk.n_cells, dg.global_corr() / dg.n_cells, rtol=1e-13, atol=1e-13)
    assert sk.energy() / sk.n_cells == pytest.approx(dg.energy() / dg.n_cells, rel=1e-12, abs=1e-12)
    assert (sk.composition() * (512 // sk.n_cells) == dg.composition()).all()
    sk.close()
    dg.close()


@pytest.mark.parametrize("case_sys,eci_key,T", [
    ("fcc", "eci_sparse", CONV * 6),
    ("fcc", "eci_full", CONV * 8),
    ("zro", "eci", [[6, 0, 0], [3, 9, 0], [0, 3, 6]]),
])
def test_sweep_on_skewed_box_keeps_the_energy_book(dev_tables, systems, case_sys, eci_key, T):
    sysd = systems[case_sys]
    mu = [0.2, -0.1][:len(sysd["axes"]["end_members"])]
    st = _capi.State(dev_tables(sysd["tables"]), None, 2, transformation_matrix=T)
    eci = sysd[eci_key]
    st.set_eci(eci["index"], eci["value"])
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], mu, sysd["n_species"])
    for r in range(2):
        st.set_conditions(800.0 + 300.0 * r, ex, r)
    st.randomize(9)
    st.set_sweep_flags(_capi.CMX_SWEEP_DE_SUM)
    info = st.sweep_info()
    assert info["evaluator"] == ("generic" if case_sys == "zro" else "pair_sum")   # (no pair LUT on skewed boxes)
    S = info["colour_strides"]
    assert st.skew[0] % S[1] == 0 and st.skew[1] % S[2] == 0 and st.skew[2] % S[2] == 0

    def potential(r):
        comp = mol_composition(st.composition(r), sysd["occ_to_species"], sysd["n_species"], st.n_cells)
        return semigrand_potential_per_supercell(st.energy(r), comp, sysd["axes"]["origin"], sysd["axes"]["Rt"],
                                                 mu, st.n_cells)

    p0 = [potential(r) for r in range(2)]
    cnt = st.sgc_sweep(5, seed=21)
    n_mut = st.n_cells * len(sysd["mutable_sublats"])
    for r in range(2):
        assert cnt[r].n_attempt == 5 * n_mut and 0 < cnt[r].n_accept < cnt[r].n_attempt
        assert cnt[r].dE_sum == pytest.approx(potential(r) - p0[r], rel=1e-9, abs=1e-7)
    st.close()


def _kmc_on(tabs, systems, load_tables, T, occ, eci, temperature, sparse=False):
    types = []
    for et in systems["fcc"]["kmc"]["event_types"]:
        kra, freq = et["kra_sparse" if sparse else "kra"], et["freq_sparse" if sparse else "freq"]
        types.append(dict(et, kra=(kra["index"], kra["value"]), freq=(freq["index"], freq["value"])))
    prim = K.make_prim_event_list(types)
    st = _capi.State(tabs("fcc_default"), None, 1, transformation_matrix=T)
    st.upload_occ(occ(st))
    st.set_conditions(temperature, None)
    st.set_eci(eci["index"], eci["value"])
    dev_types = [dict(local_tables=[tabs(n) for n in et["local_tables"]], kra=et["kra"], freq=et["freq"])
                 for et in types]
    return st, _capi.Kmc(st, dev_types, prim), prim


@pytest.mark.parametrize("sparse", [False, True])
def test_reference_kmc_known_answers_on_the_conventional_box(dev_tables, systems, load_tables, sparse):
    """events_CompleteEventCalculator_test.cpp:30-102 on ITS supercell, T = 10 *
    fcc_conventional_transf_mat (4000 unit cells): all A + one vacancy at site 0, 600 K,
    24 prim events; of the 96 000 events exactly 12 are allowed, each with dE_final = 0,
    Ekra = 1, freq = 1e12, rate = 1e12 exp(-beta); every other rate is exactly 0."""
    eci = systems["fcc"]["eci_sparse" if sparse else "eci_dense"]

    def occ(st):
        assert st.N == (10, 20, 20) and st.skew == (10, 10, 0) and st.n_cells == 4000
        o = np.zeros(st.n_sites, dtype=np.int32)
        o[0] = 2
        return o

    st, kmc, prim = _kmc_on(dev_tables, systems, load_tables, CONV * 10, occ, eci, 600.0, sparse)
    assert len(prim) == 24
    uc, pe = K.complete_event_list(st.n_cells, len(prim))
    assert len(uc) == 24 * 4000
    s = kmc.event_states(uc, pe)
    assert int(s["is_allowed"].sum()) == 12 and int((s["is_allowed"] == 0).sum()) == 4000 * 24 - 12
    a = s[s["is_allowed"] == 1]
    beta = 1.0 / (8.6173303e-05 * 600.0)
    np.testing.assert_allclose(a["dE_final"], 0.0, atol=1e-5)
    np.testing.assert_allclose(a["Ekra"], 1.0, atol=1e-5)
    np.testing.assert_allclose(a["dE_activated"], 1.0, atol=1e-5)
    np.testing.assert_allclose(a["freq"], 1e12, rtol=1e-12)
    np.testing.assert_allclose(a["rate"], 1e12 * np.exp(-beta), rtol=1e-5)
    assert (s["rate"][s["is_allowed"] == 0] == 0).all()
    # every allowed event moves the vacancy at site 0: one of the event's two sites, found
    # through the skewed images, is site 0
    box = _box_coords(st.N)
    for u, p in zip(uc[s["is_allowed"] == 1], pe[s["is_allowed"] == 1]):
        sites = np.array(prim[int(p)]["sites"])
        assert (sites[:, 0] == 0).all()
        assert 0 in st.cell_index(box[int(u)] + sites[:, 1:]).tolist()
    kmc.close()
    st.close()


def test_kmc_run_on_a_skewed_box_keeps_its_rates(dev_tables, systems, load_tables):
    """After any number of hops on a skewed box the selector's leaves equal the rates
    recomputed from scratch (the impact list wraps through the skewed images correctly),
    and species are conserved."""
    eci = systems["fcc"]["eci_dense"]
    rng = np.random.default_rng(3)
    store = {}

    def occ(st):
        store["occ"] = rng.choice(3, size=st.n_sites, p=[0.7, 0.25, 0.05]).astype(np.int32)
        return store["occ"]

    st, kmc, prim = _kmc_on(dev_tables, systems, load_tables, CONV * 4, occ, eci, 900.0)
    kmc.run_begin(np.array([17], dtype=np.uint64))
    res = kmc.run(400, log_cap=400)
    assert res["n_steps"][0] == 400 and res["time"][0] > 0
    cur, tot = kmc.current_rates()
    fresh, _ = kmc.all_rates()
    assert (cur == fresh).all()
    now = st.download_occ()
    assert (now != store["occ"]).any()
    assert (np.bincount(now, minlength=3) == np.bincount(store["occ"], minlength=3)).all()
    kmc.close()
    st.close()


def test_reference_order_mode_on_a_skewed_box(dev_tables, systems):
    """The sequential (reference-order) Metropolis on a skewed box keeps its energy book: the
    potential recomputed from scratch after the run equals the potential before plus the
    accepted dE of the step log."""
    sysd = systems["fcc"]
    eci = sysd["eci_full"]
    mu = [0.1, -0.2]
    st = _capi.State(dev_tables(sysd["tables"]), None, transformation_matrix=CONV * 4)
    st.set_eci(eci["index"], eci["value"])
    st.set_occupants(sysd["sublat_to_asym"], sysd["occ_to_species"], 3)
    st.set_conditions(1000.0, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], mu, 3))
    st.randomize(2)

    def potential():
        comp = mol_composition(st.composition(), sysd["occ_to_species"], sysd["n_species"], st.n_cells)
        return semigrand_potential_per_supercell(st.energy(), comp, sysd["axes"]["origin"], sysd["axes"]["Rt"],
                                                 mu, st.n_cells)

    p0 = potential()
    res = st.metropolis_sequential(0, 5000, 5, log_cap=5000)
    assert 0 < res["n_accept"] < 5000
    de = sum(step["dE"] for step in res["log"] if step["accepted"])
    assert de == pytest.approx(potential() - p0, rel=1e-9, abs=1e-7)
    st.close()


def test_caller_site_order(dev_tables, systems):
    sysd = systems["zro"]
    tab = dev_tables(sysd["tables"])
    eci = sysd["eci"]
    a = _capi.State(tab, (6, 4, 8))
    b = _capi.State(tab, (6, 4, 8))
    rng = np.random.default_rng(1)
    nocc = np.array(tab.host.n_occ)
    occ = np.concatenate([rng.integers(0, nocc[s], a.n_cells) for s in range(tab.host.n_sublat)]).astype(np.int32)
    order = rng.permutation(a.n_sites)
    with pytest.raises(_capi.CmxError):
        b.set_site_order(np.zeros(a.n_sites, dtype=np.int64))
    b.set_site_order(order)
    theirs = np.empty_like(occ)
    theirs[np.arange(a.n_sites)] = occ[order]        # the caller's site l_c is our site order[l_c]
    a.upload_occ(occ)
    b.upload_occ(theirs)
    for st in (a, b):
        st.set_eci(eci["index"], eci["value"])
    assert (a.global_corr() == b.global_corr()).all()
    assert (b.download_occ() == theirs).all()
    assert (b.download_occ(dtype=np.int8) == theirs.astype(np.int8)).all()
    b.set_site_order(None)
    assert (b.download_occ() == occ).all()
    a.close()
    b.close()
