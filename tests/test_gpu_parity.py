"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against
(1) the committed golden vectors produced by the reference's own generated
kernels and (2) the live oracle on fresh seeded inputs.

Tolerances: the faithful evaluators (delta corr, point corr, cell corr,
delta E, sequential trajectories) are required to be BIT-EXACT.  Global sums
over all unit cells use a different (tree) summation order than the
reference's ascending loop and are checked to rtol 1e-12 (north_star: 1e-10).
"""
import numpy as np
import pytest

from conftest import CASES

from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200.potential import (KB, mol_composition, semigrand_exchange_table,
                                               semigrand_potential_per_supercell)

pytestmark = pytest.mark.gpu


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


def make_state(dev_tables, systems, load_vectors, case, n_replicas=1):
    sysname, _ = CASES[case]
    sysd = systems[sysname]
    v = load_vectors(case)
    st = _capi.State(dev_tables(sysd["tables"]), tuple(int(x) for x in v["N"]), n_replicas)
    for r in range(n_replicas):
        st.upload_occ(v["occ"], r)
    st.set_eci(v["eci_index"], v["eci_value"])
    return st, v, sysd


@pytest.mark.parametrize("case", list(CASES))
def test_faithful_kernels_match_golden(dev_tables, systems, load_vectors, case):
    st, v, sysd = make_state(dev_tables, systems, load_vectors, case)
    assert (st.download_occ() == v["occ"]).all()
    assert (st.download_occ(dtype=np.int8) == v["occ"].astype(np.int8)).all()
    assert (st.delta_corr(v["l"], v["new_occ"]) == v["delta_corr"]).all()
    assert (st.point_corr(v["l"]) == v["point_corr"]).all()
    assert (st.cell_corr(v["cells"]) == v["cell_corr"]).all()
    assert (st.delta_e(v["l"], v["new_occ"], 1) == v["delta_e_1"]).all()
    assert (st.delta_e(v["l2"], v["new_occ2"], 2) == v["delta_e_2"]).all()
    np.testing.assert_allclose(st.global_corr(), v["global_corr"], rtol=1e-12, atol=1e-12)
    e = st.energy()
    e_ref = float(np.dot(v["eci_value"], v["global_corr"][v["eci_index"]]))
    assert e == pytest.approx(e_ref, rel=1e-12, abs=1e-12)
    # composition + semi-grand potential
    counts = st.composition()
    n_cells = int(np.prod(v["N"]))
    comp = mol_composition(counts, sysd["occ_to_species"], sysd["n_species"], n_cells)
    np.testing.assert_allclose(comp, v["mol_composition"], rtol=0, atol=1e-15)
    pot = semigrand_potential_per_supercell(e, comp, sysd["axes"]["origin"], sysd["axes"]["Rt"],
                                            v["param_chem_pot"], n_cells)
    assert pot == pytest.approx(float(v["potential_per_supercell"]), rel=1e-12)
    st.close()


def test_local_clexulators_match_golden(dev_tables, load_vectors):
    """LocalCorrelations::local for the 12 KMC local basis sets."""
    v = load_vectors("local")
    for ev in ("A_Va_1NN", "B_Va_1NN"):
        for k in range(6):
            name = f"fcc_{ev}_{k}"
            st = _capi.State(dev_tables(name), tuple(int(x) for x in v["N"]))
            st.upload_occ(v["occ"])
            assert (st.cell_corr(v["cells"]) == v[name]).all(), name
            st.close()


@pytest.mark.parametrize("case", list(CASES))
def test_semigrand_delta_potential(dev_tables, systems, load_vectors, oracle, case):
    """occ_delta_per_supercell with the exchange term, against the live oracle's
    single-step log (dE of step 0 of the sequential loop)."""
    st, v, sysd = make_state(dev_tables, systems, load_vectors, case)
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], v["param_chem_pot"],
                                  sysd["n_species"])
    st.set_conditions(float(v["sgc_T"]), ex)
    l0 = v["sgc_log_l0"][:1]
    n0 = v["sgc_log_new0"][:1]
    assert st.delta_e(l0, n0, 1, potential=True)[0] == v["sgc_log_dE"][0]
    st.close()


def test_device_rng_stream_matches_libstdcxx(load_vectors):
    """mt19937_64 + uniform_int/real on the device == libstdc++ draw by draw."""
    v = load_vectors("rng")
    oi, orl, oraw = _capi.rng_stream_test(int(v["seed"]), v["kinds"], v["int_max"], v["real_max"])
    k = v["kinds"]
    assert (oraw[k == 0] == v["out_raw"][k == 0]).all()
    assert (oi[k == 1] == v["out_int"][k == 1]).all()
    assert (orl[k == 2] == v["out_real"][k == 2]).all()


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_sequential_trajectory_matches_golden(dev_tables, systems, load_vectors, case, mode):
    """Reference-order mode reproduces the oracle's occupation trajectory."""
    st, v, sysd = make_state(dev_tables, systems, load_vectors, case)
    ex = None
    if mode == "sgc":
        ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], v["param_chem_pot"],
                                      sysd["n_species"])
    st.set_conditions(float(v[f"{mode}_T"]), ex)
    st.set_occupants(sysd["sublat_to_asym"], sysd["occ_to_species"], sysd["n_species"])
    n_steps = int(v[f"{mode}_steps"])
    res = st.metropolis_sequential(0 if mode == "sgc" else 1, n_steps, int(v[f"{mode}_seed"]), log_cap=256)
    log = res["log"]
    assert [s["l0"] for s in log] == list(v[f"{mode}_log_l0"])
    assert [s["l1"] for s in log] == list(v[f"{mode}_log_l1"])
    assert [s["new0"] for s in log] == list(v[f"{mode}_log_new0"])
    assert [s["accepted"] for s in log] == list(v[f"{mode}_log_acc"])
    assert [s["dE"] for s in log] == list(v[f"{mode}_log_dE"])
    assert res["n_accept"] == int(v[f"{mode}_n_accept"])
    assert res["hash"] == int(v[f"{mode}_hash"])
    assert (st.download_occ(dtype=np.int8) == v[f"{mode}_final_occ"]).all()
    st.close()


def test_sequential_million_steps_vs_live_oracle(dev_tables, systems, load_vectors, oracle):
    """north_star (2): 10^6 reference-order steps, bit-exact, on config 1's size
    (FCC A-B-Va, 4096-site box), canonical and semi-grand."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    sysd = systems["fcc"]
    N = 16
    rng = np.random.default_rng(42)
    occ = rng.integers(0, 3, N ** 3).astype(np.int32)
    eci = sysd["eci_sparse"]
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=3, Rt=np.array(sysd["axes"]["Rt"]))
    mu = np.array([0.1, -0.4])
    sc = oracle.RefClexulator("fcc_default").supercell(N)
    st = _capi.State(dev_tables("fcc_default"), (N, N, N))
    st.set_eci(eci["index"], eci["value"])
    st.set_occupants(sysd["sublat_to_asym"], sysd["occ_to_species"], 3)
    for mode, T in ((0, 1000.0), (1, 600.0)):
        ref = sc.metropolis_run(mode, occ, prim, eci["index"], eci["value"], T, seed=99, n_steps=10 ** 6,
                                param_chem_pot=mu if mode == 0 else None)
        st.upload_occ(occ)
        ex = semigrand_exchange_table(sysd["occ_to_species"], prim["Rt"], mu, 3) if mode == 0 else None
        st.set_conditions(T, ex)
        res = st.metropolis_sequential(mode, 10 ** 6, 99)
        mism = int((st.download_occ() != ref["occ"]).sum())
        # north_star: "any exact-tie accept/reject divergences reported": the device counts the
        # steps whose uniform draw lay within one ulp of exp(-beta dE) -- the only steps where
        # CUDA's and glibc's exp could decide differently -- and names them
        assert (res["n_accept"], res["hash"], mism) == (ref["n_accept"], ref["hash"], 0), \
            (f"mode {mode}: gpu acc {res['n_accept']} ref {ref['n_accept']} mismatching sites {mism}; "
             f"near ties reported: {res['n_near_ties']} at steps {res['tie_steps']}")
        assert res["n_near_ties"] == 0 and res["tie_steps"] == [], \
            f"mode {mode}: trajectory identical, but {res['n_near_ties']} near ties at steps {res['tie_steps']}"
    st.close()


# ---------------------------------------------------------------------------
# checkerboard sweeps
# ---------------------------------------------------------------------------
def _sweep_state(dev_tables, systems, case_sys, eci_key, N, T, mu, n_replicas=1, seed=1, linear_rows=False):
    sysd = systems[case_sys]
    st = _capi.State(dev_tables(sysd["tables"]), N, n_replicas, linear_rows=linear_rows)
    eci = sysd[eci_key]
    st.set_eci(eci["index"], eci["value"])
    ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], mu, sysd["n_species"])
    for r in range(n_replicas):
        st.set_conditions(T, ex, r)
    st.randomize(seed)
    return st, sysd, ex


@pytest.mark.parametrize("case_sys,eci_key,N,expect", [
    ("fcc", "eci_sparse", (16, 16, 16), "pair_lut"),
    ("fcc", "eci_full", (16, 16, 16), "pair_lut2"),   # two neighbor classes: two-class count table
    ("fcc", "eci_full", (8, 8, 8), "pair_sum"),       # ... on linear rows: per-neighbor tables
    ("fcc", "eci_sparse", (12, 12, 12), "pair_sum"),  # N0 % 16 != 0: outside the pair-LUT path
    ("fcc", "eci_sparse", (64, 6, 10), "pair_lut"),
    ("zro", "eci", (8, 8, 8), "generic"),
    ("fcc_syn", "eci", (16, 8, 8), "generic"),        # synthetic FCC binary pair + triplet basis
])
def test_sweep_energy_bookkeeping(dev_tables, systems, case_sys, eci_key, N, expect):
    """Size-independent property: the sum of accepted dE reported by the sweep
    equals E_potential(after) - E_potential(before) recomputed from scratch by
    the (faithful) global evaluation; occupants stay in range; counters add up."""
    mu = [0.2, -0.1][:len(systems[case_sys]["axes"]["end_members"])]
    st, sysd, ex = _sweep_state(dev_tables, systems, case_sys, eci_key, N, 900.0, mu)
    st.set_sweep_flags(_capi.CMX_SWEEP_DE_SUM)
    info = st.sweep_info()
    assert info["evaluator"] == expect
    n_cells = int(np.prod(N))

    def potential():
        comp = mol_composition(st.composition(), sysd["occ_to_species"], sysd["n_species"], n_cells)
        return semigrand_potential_per_supercell(st.energy(), comp, sysd["axes"]["origin"],
                                                 sysd["axes"]["Rt"], mu, n_cells)

    p0 = potential()
    cnt = st.sgc_sweep(5, seed=77)
    p1 = potential()
    n_mut = n_cells * len(sysd["mutable_sublats"])
    assert cnt[0].n_attempt == 5 * n_mut
    assert 0 < cnt[0].n_accept < cnt[0].n_attempt
    assert cnt[0].dE_sum == pytest.approx(p1 - p0, rel=1e-9, abs=1e-7)
    occ = st.download_occ()
    nocc = np.array(st.tables.host.n_occ)
    for b in range(len(nocc)):
        seg = occ[b * n_cells:(b + 1) * n_cells]
        assert seg.min() >= 0 and seg.max() < nocc[b]
    st.close()


def test_sweep_is_deterministic_and_replica_independent(dev_tables, systems):
    """Counter-based RNG: same seed -> same trajectory; replicas with identical
    conditions but different replica index diverge; a batched run equals single runs."""
    N = (16, 16, 16)
    mu = [0.2, -0.1]
    a, _, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 900.0, mu, n_replicas=3, seed=5)
    b, _, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 900.0, mu, n_replicas=3, seed=5)
    a.sgc_sweep(3, seed=9)
    b.sgc_sweep(2, seed=9)
    b.sgc_sweep(1, seed=9, first_sweep=2)   # continuing the stream == one call
    for r in range(3):
        assert (a.download_occ(r) == b.download_occ(r)).all()
    assert (a.download_occ(0) != a.download_occ(1)).any()
    a.close()
    b.close()


@pytest.mark.parametrize("kernel", ["pass", "stream", "block"])
@pytest.mark.parametrize("N,n_replicas", [((16, 16, 16), 2), ((32, 8, 6), 1), ((48, 4, 4), 1), ((512, 2, 2), 1),
                                          ((128, 32, 4), 2), ((128, 24, 20), 3), ((64, 2, 2), 1), ((256, 6, 12), 1)])
def test_pair_lut_kernels_equal_generic_kernel_bit_for_bit(dev_tables, systems, N, n_replicas, kernel):
    """The pair-LUT kernels -- on x4-interleaved rows the colour-pass kernel (the default: one
    cooperative launch per call, grid barriers) and the streaming kernel (CMX_SWEEP_STREAM:
    per-layer completion counters instead of barriers), and the block kernel on linear rows --
    and the one-site-per-thread generic evaluator (whose delta E is checked against the
    reference kernels) draw the same random bits and must make the same decisions:
    identical occupation after several sweeps, identical acceptance counts.  Covers one
    chunk per row (N0=16), rows that do not fill a block (N0=48: block kernel only), a row
    spanning a whole warp (N0=512), partial row-steps (J not a multiple of the rows per
    warp), the smallest boxes (two layers, one row per colour), replicas with different
    conditions and the int8 transfer path.  The generic evaluator runs on the SAME layout
    as the kernel under test, so the layout-dependent assignment of random fields to sites
    is covered on both."""
    mu = [0.2, -0.1]
    lin = kernel == "block"
    a, sysd, ex = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 900.0, mu, n_replicas=n_replicas, seed=5,
                               linear_rows=lin)
    b, _, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 900.0, mu, n_replicas=n_replicas, seed=5,
                           linear_rows=lin)
    if n_replicas > 1:
        ex2 = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [-0.3, 0.4], 3)
        a.set_conditions(500.0, ex2, replica=1)
        b.set_conditions(500.0, ex2, replica=1)
    kflag = _capi.CMX_SWEEP_STREAM if kernel == "stream" else 0
    a.set_sweep_flags(_capi.CMX_SWEEP_DE_SUM | kflag)
    b.set_sweep_flags(_capi.CMX_SWEEP_FORCE_GENERIC | _capi.CMX_SWEEP_DE_SUM)
    ia = a.sweep_info()
    assert ia["evaluator"] == "pair_lut" and b.sweep_info()["evaluator"] == "generic"
    pow2 = N[0] <= 512 and (N[0] & (N[0] - 1)) == 0
    assert ia["stream"] == (kernel == "stream" and pow2)
    assert ia["one_launch_per_call"] == (kernel != "block" and pow2)
    for r in range(n_replicas):
        assert (a.download_occ(r) == b.download_occ(r)).all()
    ca = a.sgc_sweep(6, seed=9)
    cb = b.sgc_sweep(6, seed=9)
    for r in range(n_replicas):
        oa, ob = a.download_occ(r), b.download_occ(r)
        assert (oa == ob).all(), f"replica {r}: {(oa != ob).sum()} sites differ"
        assert (a.download_occ(r, dtype=np.int8) == oa).all()
        assert ca[r].n_accept == cb[r].n_accept and ca[r].n_attempt == cb[r].n_attempt
        assert ca[r].dE_sum == pytest.approx(cb[r].dE_sum, rel=1e-9, abs=1e-9)
    # without the dE accumulation (the default) the trajectory is the same
    a.set_sweep_flags(kflag)
    ca = a.sgc_sweep(2, seed=9, first_sweep=6)
    cb = b.sgc_sweep(2, seed=9, first_sweep=6)
    for r in range(n_replicas):
        assert (a.download_occ(r) == b.download_occ(r)).all()
        assert ca[r].n_accept == cb[r].n_accept and ca[r].dE_sum == 0.0
    a.close()
    b.close()


def test_streaming_and_block_kernel_share_their_statistics(dev_tables, systems):
    """The two pair-LUT kernels assign the random fields to the sites differently (the
    layouts differ), so their trajectories differ -- but they sample the same ensemble:
    acceptance rate and composition agree statistically on a 64^3 box."""
    N, mu = (64, 64, 64), [0.1, -0.2]
    res = []
    for lin in (False, True):
        st, sysd, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 1000.0, mu, seed=3, linear_rows=lin)
        st.sgc_sweep(30, seed=21)
        cnt = st.sgc_sweep(60, seed=21, first_sweep=30)
        comp = st.composition()[0] / float(np.prod(N))
        res.append((cnt[0].n_accept / cnt[0].n_attempt, comp))
        st.close()
    assert res[0][0] == pytest.approx(res[1][0], abs=2e-3)
    np.testing.assert_allclose(res[0][1], res[1][1], atol=3e-3)


def test_conditions_change_after_a_sweep_is_honoured(dev_tables, systems):
    """the pair-LUT acceptance tables are functions of (T, mu); changing
    the conditions of a state that has already swept must rebuild them.  sweep ->
    set_conditions -> sweep equals (bit for bit) a fresh state at the new conditions that
    starts from the same occupation, and the generic evaluator, which reads beta and the
    exchange table directly."""
    N, mu0, mu1 = (32, 16, 8), [0.2, -0.1], [-0.4, 0.3]
    a, sysd, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 900.0, mu0, seed=5)
    a.sgc_sweep(3, seed=9)
    occ_mid = a.download_occ()
    ex1 = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], mu1, 3)
    a.set_conditions(400.0, ex1)
    ca = a.sgc_sweep(3, seed=9, first_sweep=3)
    for flags in (0, _capi.CMX_SWEEP_FORCE_GENERIC):
        b, _, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 400.0, mu1, seed=5)
        b.upload_occ(occ_mid)
        b.set_sweep_flags(flags)
        cb = b.sgc_sweep(3, seed=9, first_sweep=3)
        assert (a.download_occ() == b.download_occ()).all(), f"flags {flags}"
        assert ca[0].n_accept == cb[0].n_accept
        b.close()
    a.close()


@pytest.mark.parametrize("N,n_replicas,n_sweeps", [((512, 512, 512), 1, 6), ((128, 128, 128), 8, 6)])
def test_full_size_sweep_variants_agree(dev_tables, systems, N, n_replicas, n_sweeps):
    """BASELINE sizes (configs[2]: 512^3; configs[1]: 128^3 replicas with a (mu, T) grid):
    the colour-pass kernel (the default), the streaming kernel and the generic
    one-site-per-thread evaluator must leave the SAME occupation and the same acceptance
    counts.  Only at these sizes are all SMs busy and the layer counters, the wavefront
    order and the seam of the periodic box really exercised (thousands of warps in flight
    across tens of units), and the persisting-L2 window of the pass kernel in effect."""
    mu = [0.0, 0.0]
    variants = {"pass": 0, "stream": _capi.CMX_SWEEP_STREAM, "generic": _capi.CMX_SWEEP_FORCE_GENERIC}
    ref_occ, ref_cnt = None, None
    for name, flags in variants.items():
        st, sysd, ex = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 800.0, mu,
                                    n_replicas=n_replicas, seed=2026)
        for r in range(1, n_replicas):
            exr = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"],
                                           [-1.0 + 2.0 * r / n_replicas, 0.0], 3)
            st.set_conditions(400.0 + 200.0 * r, exr, replica=r)
        st.set_sweep_flags(flags)
        info = st.sweep_info()
        assert info["stream"] == (name == "stream")
        cnt = st.sgc_sweep(n_sweeps - 2, seed=7)
        cnt2 = st.sgc_sweep(2, seed=7, first_sweep=n_sweeps - 2)   # a second call continues the counters
        occ = [st.download_occ(r, dtype=np.int8) for r in range(n_replicas)]
        acc = [cnt[r].n_accept + cnt2[r].n_accept for r in range(n_replicas)]
        st.close()
        if ref_occ is None:
            ref_occ, ref_cnt = occ, acc
            continue
        for r in range(n_replicas):
            assert (occ[r] == ref_occ[r]).all(), f"{name} vs pass, replica {r}: {(occ[r] != ref_occ[r]).sum()} sites differ"
        assert acc == ref_cnt, name


@pytest.mark.parametrize("case_sys,eci_key,N,flags", [
    ("fcc", "eci_sparse", (16, 8, 8), 0),                                  # pair-LUT table (streaming kernel)
    ("fcc", "eci_sparse", (48, 8, 8), 0),                                  # pair-LUT table (block kernel, linear rows)
    ("fcc", "eci_sparse", (16, 8, 8), _capi.CMX_SWEEP_FORCE_GENERIC),      # folded terms, thread evaluator
    ("fcc", "eci_full", (8, 8, 8), 0),                                     # 1NN + 2NN pairs: pair-sum tables
    ("fcc", "eci_full", (16, 8, 8), 0),                                    # 1NN + 2NN pairs: two-class count table
    ("fcc", "eci_full", (32, 6, 4), 0),
    ("fcc", "eci_full", (16, 8, 8), _capi.CMX_SWEEP_PAIR_SUM),
    ("fcc", "eci_full", (8, 8, 8), _capi.CMX_SWEEP_THREAD_GENERIC),        # the same through the term lists
    ("fcc", "eci_full", (12, 6, 10), 0),                                   # linear rows
    ("fcc", "eci_2", (8, 8, 8), 0),
    ("zro", "eci", (8, 8, 8), 0),                                          # warp evaluator (quadruplets)
    ("zro", "eci", (8, 8, 8), _capi.CMX_SWEEP_THREAD_GENERIC),
    ("fcc_syn", "eci", (8, 8, 8), 0),                                      # FCC triplets: term lists per thread
])
def test_sweep_evaluators_delta_e_per_proposal(dev_tables, systems, case_sys, eci_key, N, flags):
    """north_star (1): dE within 1e-10 relative, proposal by proposal, for the evaluators
    the SWEEPS use: the pair-LUT table entry, the ECI-folded merged
    term lists (one site per thread) and the warp-staged evaluator -- every mutable site x
    every alternative occupant of the box -- against cmx_delta_e (the faithful evaluator,
    bit-exact to the reference's generated kernels, test_faithful_kernels_match_golden)
    with the semi-grand exchange term of SemiGrandCanonicalCalculator.cc:186-213."""
    sysd = systems[case_sys]
    mu = [0.2, -0.1][:len(sysd["axes"]["end_members"])]
    st, sysd, ex = _sweep_state(dev_tables, systems, case_sys, eci_key, N, 900.0, mu, seed=13)
    st.set_sweep_flags(flags)
    n_cells = int(np.prod(N))
    nocc = np.array(st.tables.host.n_occ)
    occ = st.download_occ()
    ls, news = [], []
    for b in sysd["mutable_sublats"]:
        for alt in range(1, nocc[b]):
            l = b * n_cells + np.arange(n_cells)
            ls.append(l)
            news.append((occ[l] + alt) % nocc[b])
    l = np.concatenate(ls)
    new = np.concatenate(news).astype(np.int32)
    got = st.sweep_debug_delta_e(l, new)
    want = st.delta_e(l, new, 1, potential=True)
    scale = np.abs(want).max()
    assert np.isfinite(got).all()
    err = np.abs(got - want).max()
    assert err <= 1e-10 * scale, f"{st.sweep_info()['evaluator']}: max |dE - dE_ref| = {err:g} (scale {scale:g})"
    st.close()


@pytest.mark.parametrize("N", [(16, 6, 4), (48, 10, 8), (512, 4, 2)])
def test_fast_energy_equals_faithful_global_evaluation(dev_tables, systems, load_vectors, N):
    """ClusterExpansion::per_supercell through the count-table streaming kernel
    (cmx_energy.cu) against the sum of faithful per-cell contributions (rtol 1e-12;
    the reference sums cells in ascending order, both kernels use tree sums)."""
    sysd = systems["fcc"]
    for eci_key in ("eci_sparse", "eci_2"):
        a, _, _ = _sweep_state(dev_tables, systems, "fcc", eci_key, N, 900.0, [0.1, 0.2], n_replicas=2, seed=11)
        for r in range(2):
            fast = a.energy(r)
            a.set_sweep_flags(_capi.CMX_SWEEP_FORCE_GENERIC)
            slow = a.energy(r)
            a.set_sweep_flags(0)
            g = a.global_corr(r)
            eci = sysd[eci_key]
            ref = float(np.dot(eci["value"], g[eci["index"]]))
            assert slow == pytest.approx(ref, rel=1e-13)
            assert fast == pytest.approx(ref, rel=1e-12, abs=1e-9)
        a.close()
    # golden configuration of the reference kernels
    v = load_vectors("fcc_sparse")
    st = _capi.State(dev_tables("fcc_default"), tuple(int(x) for x in v["N"]))
    st.upload_occ(v["occ"])
    st.set_eci(v["eci_index"], v["eci_value"])
    ref = float(np.dot(v["eci_value"], v["global_corr"][v["eci_index"]]))
    assert st.energy() == pytest.approx(ref, rel=1e-12)
    st.close()


def test_int8_round_trip_of_coded_states(dev_tables):
    """Ternary single-sublattice states store occupant 2 as 16 on the device
    (Geom::coded); every transfer path must hide that."""
    st = _capi.State(dev_tables("fcc_default"), (16, 4, 2))
    occ = np.random.default_rng(3).integers(0, 3, 128).astype(np.int32)
    st.upload_occ(occ)
    assert (st.download_occ() == occ).all()
    assert (st.download_occ(dtype=np.int8) == occ).all()
    st.upload_occ(occ.astype(np.int8))
    assert (st.download_occ() == occ).all()
    assert (st.composition()[0] == np.bincount(occ, minlength=3)).all()
    with pytest.raises(_capi.CmxError):
        st.upload_occ(np.full(128, 3, dtype=np.int8))
    st.close()


def test_pair_lut_sweep_equals_generic_sweep(dev_tables, systems):
    """The LUT fast path and the pair-sum evaluator (a box outside the LUT path) make the same
    decisions when they see the same random numbers?  They use different RNG counters, so we
    check the PHYSICS instead: acceptance rate and energy after equilibration
    agree within statistics (two independent evaluators, same ensemble)."""
    mu = [0.2, -0.1]
    T = 1200.0
    res = {}
    for N in ((16, 16, 16), (12, 12, 12)):
        st, sysd, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, T, mu, seed=3)
        st.sgc_sweep(200, seed=11)
        acc, e = [], []
        for k in range(20):
            cnt = st.sgc_sweep(10, seed=11, first_sweep=200 + 10 * k)
            acc.append(cnt[0].n_accept / cnt[0].n_attempt)
            e.append(st.energy() / np.prod(N))
        res[st.sweep_info()["evaluator"]] = (np.mean(acc), np.std(acc) / np.sqrt(20), np.mean(e),
                                             np.std(e) / np.sqrt(20))
        st.close()
    (a1, sa1, e1, se1), (a2, sa2, e2, se2) = res["pair_lut"], res["pair_sum"]
    assert abs(a1 - a2) < 5 * np.hypot(sa1, sa2) + 2e-3
    assert abs(e1 - e2) < 5 * np.hypot(se1, se2) + 2e-3


@pytest.mark.parametrize("case_sys,eci_key,N,linear_rows,n_replicas", [
    ("fcc", "eci_full", (16, 8, 8), False, 1),     # x4-interleaved rows
    ("fcc", "eci_full", (64, 6, 4), False, 3),     # several replicas with different conditions
    ("fcc", "eci_full", (12, 10, 6), False, 1),    # linear rows (N0 not a power of two)
    ("fcc", "eci_sparse", (12, 12, 12), False, 2),
    ("fcc", "eci_2", (24, 8, 8), False, 1),
    ("fcc", "eci_full", (4, 4, 4), False, 1),      # every site on a periodic seam
])
def test_pair_sum_kernel_equals_term_list_kernel(dev_tables, systems, case_sys, eci_key, N, linear_rows, n_replicas):
    """The pair-sum evaluator (per-neighbor tables folded from the term lists; the evaluator
    of the reference's dense FCC ECI -- points, 1NN and 2NN pairs, SURVEY 8d's 19-site
    neighbourhood) draws the same random bits as the one-site-per-thread term-list kernel
    and must leave the SAME occupation and counters: its dE differs only in the summation
    order (checked per proposal in test_sweep_evaluators_delta_e_per_proposal), which moves
    a 47/53-bit acceptance threshold with probability ~1e-15 per step.  Boxes with the
    seam-free fast path and the wrapped path, both row layouts, replicas, several calls."""
    mu = [0.2, -0.1]
    sts = []
    for flags in (_capi.CMX_SWEEP_PAIR_SUM, _capi.CMX_SWEEP_THREAD_GENERIC):
        st, sysd, ex = _sweep_state(dev_tables, systems, case_sys, eci_key, N, 900.0, mu, n_replicas=n_replicas,
                                    seed=5, linear_rows=linear_rows)
        for r in range(n_replicas):
            st.set_conditions(700.0 + 250.0 * r, ex, r)
        st.set_sweep_flags(flags | _capi.CMX_SWEEP_DE_SUM)
        sts.append(st)
    a, b = sts
    assert a.sweep_info()["evaluator"] == "pair_sum" and b.sweep_info()["evaluator"] == "generic"
    for call in range(3):
        ca = a.sgc_sweep(4, seed=21, first_sweep=4 * call)
        cb = b.sgc_sweep(4, seed=21, first_sweep=4 * call)
        for r in range(n_replicas):
            assert (a.download_occ(r) == b.download_occ(r)).all(), f"call {call}, replica {r}"
            assert (ca[r].n_attempt, ca[r].n_accept) == (cb[r].n_attempt, cb[r].n_accept)
            assert 0 < ca[r].n_accept < ca[r].n_attempt
            assert ca[r].dE_sum == pytest.approx(cb[r].dE_sum, rel=1e-9, abs=1e-9)
    a.close()
    b.close()


@pytest.mark.parametrize("N,n_replicas", [
    ((16, 8, 8), 1),       # one chunk per row: 32 rows per row-step, partial row-steps
    ((64, 6, 4), 3),       # several replicas with different conditions (one table each)
    ((32, 32, 6), 2),      # whole row-steps only
    ((128, 4, 2), 1),      # the smallest box in j and k: every row on both seams
    ((512, 8, 4), 1),      # BASELINE row length: one row per row-step
    ((64, 2, 2), 1),       # two layers, one row per colour: the diagonal rows coincide
    ((512, 256, 32), 1),   # every warp of a co-resident grid busy for several rounds
])
def test_two_class_table_kernel_equals_pair_sum_and_term_list_kernels(dev_tables, systems, N, n_replicas):
    """The reference's dense FCC ECI (points + 1NN + 2NN pairs, SURVEY 8d's 19-site neighbourhood)
    on x4-interleaved rows runs the colour-pass kernel with a two-class count table
    (evaluator "pair_lut2": k_sweep_pass16 with a second byte-lane sum).  Its table entries are
    dE values summed in the pair-sum evaluator's order, it draws the same random bits as the
    pair-sum kernel (CMX_SWEEP_PAIR_SUM) and the term-list kernel (CMX_SWEEP_THREAD_GENERIC),
    and must leave the SAME occupation and counters as both, call after call; a change of
    conditions between calls rebuilds its acceptance tables."""
    mu = [0.2, -0.1]
    sts = []
    for flags in (0, _capi.CMX_SWEEP_PAIR_SUM, _capi.CMX_SWEEP_THREAD_GENERIC):
        st, sysd, ex = _sweep_state(dev_tables, systems, "fcc", "eci_full", N, 900.0, mu, n_replicas=n_replicas, seed=5)
        for r in range(n_replicas):
            st.set_conditions(700.0 + 250.0 * r, ex, r)
        st.set_sweep_flags(flags | _capi.CMX_SWEEP_DE_SUM)
        sts.append(st)
    a, b, c = sts
    assert [x.sweep_info()["evaluator"] for x in sts] == ["pair_lut2", "pair_sum", "generic"]
    assert a.sweep_info()["one_launch_per_call"]
    ex2 = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [-0.3, 0.25], sysd["n_species"])
    for call in range(4):
        if call == 2:   # new T and mu on every state: the acceptance tables must follow
            for st in sts:
                for r in range(n_replicas):
                    st.set_conditions(1100.0 - 150.0 * r, ex2, r)
        cnts = [st.sgc_sweep(3, seed=21, first_sweep=3 * call) for st in sts]
        for r in range(n_replicas):
            occ_a = a.download_occ(r)
            for name, st, cn in (("pair_sum", b, cnts[1]), ("generic", c, cnts[2])):
                diff = int((occ_a != st.download_occ(r)).sum())
                assert diff == 0, f"call {call}, replica {r}: {diff} sites differ from the {name} kernel"
                assert (cnts[0][r].n_attempt, cnts[0][r].n_accept) == (cn[r].n_attempt, cn[r].n_accept)
                assert cnts[0][r].dE_sum == pytest.approx(cn[r].dE_sum, rel=1e-9, abs=1e-9)
            assert 0 < cnts[0][r].n_accept < cnts[0][r].n_attempt
    # without the dE sum (the instantiation the benchmark runs) the trajectory is the same
    a.set_sweep_flags(0)
    b.set_sweep_flags(_capi.CMX_SWEEP_PAIR_SUM)
    ca, cb = a.sgc_sweep(2, seed=22), b.sgc_sweep(2, seed=22)
    for r in range(n_replicas):
        assert (a.download_occ(r) == b.download_occ(r)).all()
        assert ca[r].n_accept == cb[r].n_accept and ca[r].dE_sum == 0.0
    for st in sts:
        st.close()


def test_two_class_table_kernel_gives_way_when_the_replicas_do_not_fit(dev_tables, systems):
    """The two-class kernel is one cooperative launch with a grid row per replica (2 blocks per
    SM: 296 rows on a B200).  A state with more replicas sweeps with the pair-sum kernel
    instead of failing -- and leaves the trajectory the two-class kernel leaves on a state
    that holds only the first replicas."""
    mu = [0.2, -0.1]
    N, many, few = (16, 4, 4), 320, 3
    a, sysd, ex = _sweep_state(dev_tables, systems, "fcc", "eci_full", N, 900.0, mu, n_replicas=many, seed=5)
    b, _, _ = _sweep_state(dev_tables, systems, "fcc", "eci_full", N, 900.0, mu, n_replicas=few, seed=5)
    ca = a.sgc_sweep(3, seed=21)
    cb = b.sgc_sweep(3, seed=21)
    assert a.sweep_info()["evaluator"] == "pair_sum" and b.sweep_info()["evaluator"] == "pair_lut2"
    for r in range(few):
        assert (a.download_occ(r) == b.download_occ(r)).all()
        assert ca[r].n_accept == cb[r].n_accept
    a.close()
    b.close()


@pytest.mark.parametrize("T", [500.0, 900.0, 1500.0])
@pytest.mark.parametrize("mu", [(0.3, -0.4), (0.0, 0.0), (-0.2, 0.2)])
def test_checkerboard_matches_sequential_thermodynamics(dev_tables, systems, oracle, T, mu):
    """north_star (3): checkerboard-mode thermodynamic averages -- energy, composition, heat
    capacity, susceptibility -- agree with the reference's sequential random-site Metropolis
    (the restated loop around the reference's compiled kernels) within 3 sigma over
    independent runs, on a 3 x 3 grid of (T, param_chem_pot) that reaches down to 500 K.
    sigma = run-to-run scatter of both sides; no extra slack on the fluctuation quantities."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    sysd = systems["fcc"]
    eci = sysd["eci_sparse"]
    N = 8
    n_cells = N ** 3
    mu = np.array(mu)
    Rt = np.array(sysd["axes"]["Rt"])
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=3, Rt=Rt, origin=np.array(sysd["axes"]["origin"]))
    sc = oracle.RefClexulator("fcc_default").supercell(N)
    n_runs, n_eq, n_smp, period = 8, 150, 60, 4
    kT = KB * T

    def stats(e_pot, xb):
        """per-run estimates: <E_pot>/cell, <x_B>, C_v, chi_BB (analysis_functions.cc:43-173)"""
        return [np.mean(e_pot) / n_cells, np.mean(xb), np.var(e_pot) / (kT * T * n_cells), np.var(xb) * n_cells / kT]

    ref = []
    for run in range(n_runs):
        occ = np.random.default_rng(100 + run).integers(0, 3, n_cells).astype(np.int32)
        occ = sc.metropolis_run(0, occ, prim, eci["index"], eci["value"], T, seed=1000 + run,
                                n_steps=n_eq * n_cells, param_chem_pot=mu)["occ"]
        ep, xb = [], []
        for k in range(n_smp):
            occ = sc.metropolis_run(0, occ, prim, eci["index"], eci["value"], T, seed=5000 + 97 * run + k,
                                    n_steps=period * n_cells, param_chem_pot=mu)["occ"]
            g = sc.global_corr(occ)
            n = np.bincount(occ, minlength=3) / n_cells
            x = Rt @ (n - prim["origin"])
            ep.append(float(np.dot(eci["value"], g[eci["index"]])) - n_cells * float(mu @ x))
            xb.append(n[1])
        ref.append(stats(np.array(ep), np.array(xb)))
    st, _, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", (N, N, N), T, mu, n_replicas=n_runs, seed=8)
    st.sgc_sweep(n_eq, seed=21)
    ge = np.zeros((n_runs, n_smp))
    gx = np.zeros((n_runs, n_smp))
    for k in range(n_smp):
        st.sgc_sweep(period, seed=21, first_sweep=n_eq + period * k)
        for r in range(n_runs):
            n = st.composition(r)[0] / n_cells
            ge[r, k] = st.energy(r) - n_cells * float(mu @ (Rt @ (n - prim["origin"])))
            gx[r, k] = n[1]
    gpu = [stats(ge[r], gx[r]) for r in range(n_runs)]
    st.close()
    ref, gpu = np.array(ref), np.array(gpu)
    floors = [1e-4, 1e-3, 0.0, 0.0]   # resolution floors of the two means (E per cell [eV], x_B); none on C_v, chi
    for q, name in enumerate(("potential energy per cell", "x_B", "heat capacity", "chi_BB")):
        se = np.hypot(np.std(ref[:, q], ddof=1), np.std(gpu[:, q], ddof=1)) / np.sqrt(n_runs)
        d = abs(ref[:, q].mean() - gpu[:, q].mean())
        assert d <= 3 * se + floors[q], f"{name} at T={T}, mu={mu}: reference {ref[:, q].mean():.6g} gpu {gpu[:, q].mean():.6g} (3 sigma = {3 * se:.3g})"


@pytest.mark.parametrize("N,linear", [((16, 6, 4), False), ((48, 10, 8), True), ((512, 4, 2), False), ((64, 64, 64), False)])
def test_streaming_global_corr_equals_faithful_sum(dev_tables, systems, N, linear):
    """Correlations::per_supercell (the `corr.<bset>` sampler) through the streaming bond-count
    passes (cmx_energy.cu: every function of a point + pair basis is linear in the species
    counts over its orbit's forward neighbors; one integer pass per forward-neighbor set)
    against the faithful term-by-term sum over all cells: ALL nine functions of the FCC basis
    (constant, points, 1NN and 2NN pairs), both row layouts, rtol 1e-12."""
    st, sysd, _ = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 900.0, [0.1, 0.2], n_replicas=2, seed=11,
                               linear_rows=linear)
    for r in range(2):
        fast = st.global_corr(r)
        st.set_sweep_flags(_capi.CMX_SWEEP_FORCE_GENERIC)
        slow = st.global_corr(r)
        st.set_sweep_flags(0)
        assert np.abs(slow).min() > 0      # every function is exercised
        np.testing.assert_allclose(fast, slow, rtol=1e-12, atol=1e-9)
    # not applicable (triplets / several sublattices): the faithful kernel serves the call
    st.close()


def test_error_paths(dev_tables, load_tables):
    """Error behaviour mirrors the reference: bad input -> exception, state intact."""
    st = _capi.State(dev_tables("fcc_default"), (8, 8, 8))
    with pytest.raises(_capi.CmxError):
        st.energy()                      # no ECI bound
    with pytest.raises(_capi.CmxError):
        st.upload_occ(np.full(512, 7, dtype=np.int32))   # occupant index out of range
    with pytest.raises(_capi.CmxError):
        st.upload_occ(np.zeros(5, dtype=np.int32))
    with pytest.raises(_capi.CmxError):
        st.delta_corr([10 ** 9], [1])
    with pytest.raises(_capi.CmxError):
        st.set_eci([99], [1.0])
    with pytest.raises(_capi.CmxError):
        st.set_conditions(-5.0)
    st.set_eci([1], [0.5])
    with pytest.raises(_capi.CmxError):
        st.sgc_sweep(1, seed=1)           # conditions not set
    assert st.delta_e(np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int32)).shape == (0,)
    with pytest.raises(_capi.CmxError):
        _capi.State(dev_tables("fcc_default"), (0, 8, 8))
    st.close()


def test_async_pipeline_equals_synchronous_calls(dev_tables, systems):
    """Asynchronous upload / sweeps / download on several states (one stream each, pinned
    host buffers) leave the same occupations and counters as the synchronous calls, and an
    occupant index out of range in an asynchronous upload surfaces at cmx_state_synchronize."""
    import torch
    N = (64, 16, 16)
    n = int(np.prod(N))
    rng = np.random.default_rng(3)
    jobs = [rng.integers(0, 3, n).astype(np.int8) for _ in range(5)]
    ref, ref_acc = [], []
    st, sysd, ex = _sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 800.0, [0.1, -0.2], seed=1)
    for k, occ in enumerate(jobs):
        st.upload_occ(occ)
        c = st.sgc_sweep(3, seed=5, first_sweep=3 * k)
        ref.append(st.download_occ(dtype=np.int8))
        ref_acc.append(c[0].n_accept)
    states = [st] + [_sweep_state(dev_tables, systems, "fcc", "eci_sparse", N, 800.0, [0.1, -0.2], seed=1)[0]
                     for _ in range(2)]
    h_in = [torch.empty(n, dtype=torch.int8).pin_memory().numpy() for _ in range(3)]
    h_out = [torch.empty(n, dtype=torch.int8).pin_memory().numpy() for _ in range(3)]
    got, got_acc = {}, {}
    for k in range(len(jobs) + 3):
        b = k % 3
        if k >= 3:
            got_acc[k - 3] = states[b].counters_read()[0].n_accept
            got[k - 3] = h_out[b].copy()
        if k < len(jobs):
            h_in[b][:] = jobs[k]
            states[b].upload_occ_async(h_in[b])
            states[b].sgc_sweep_async(3, seed=5, first_sweep=3 * k)
            states[b].download_occ_async(h_out[b])
    for k in range(len(jobs)):
        assert (got[k] == ref[k]).all() and got_acc[k] == ref_acc[k]
    h_in[0][:] = 0
    h_in[0][7] = 5
    states[0].upload_occ_async(h_in[0])
    with pytest.raises(_capi.CmxError):
        states[0].synchronize()
    states[0].synchronize()            # the error is reported once
    for s2 in states:
        s2.close()
