"""GPU tests of the canonical pair-exchange sweeps (csrc/cmx_canonical.cu)."""
import numpy as np
import pytest

from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200.potential import canonical_swap_types

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


def _state(dev_tables, systems, case_sys, eci_key, N, T, occ, n_replicas=1):
    sysd = systems[case_sys]
    st = _capi.State(dev_tables(sysd["tables"]), N, n_replicas)
    eci = sysd[eci_key]
    st.set_eci(eci["index"], eci["value"])
    mo = st.tables.host.max_occ
    o2s = np.full((len(sysd["occ_to_species"]), mo), -1, dtype=np.int32)
    for b, row in enumerate(sysd["occ_to_species"]):
        o2s[b, :len(row)] = row
    st.set_occupants(sysd["sublat_to_asym"], o2s, sysd["n_species"])
    for r in range(n_replicas):
        st.set_conditions(T, None, r)
        st.upload_occ(occ, r)
    swaps = canonical_swap_types(st.tables.host, sysd["sublat_to_asym"], o2s.tolist(), N)
    # the library's default swap table (what a C / C++ caller such as the canonical plugin gets)
    assert st.canonical_default_swaps() == [(a, b, tuple(t)) for a, b, t in swaps]
    return st, sysd, swaps


@pytest.mark.parametrize("case_sys,eci_key,N", [
    ("fcc", "eci_sparse", (16, 8, 8)),
    ("fcc", "eci_full", (8, 8, 12)),
    ("zro", "eci", (8, 8, 8)),
])
def test_canonical_sweep_conserves_composition_and_energy(dev_tables, systems, case_sys, eci_key, N):
    """Size-independent properties: occupant counts per sublattice group are
    conserved exactly; the accepted delta E sum to E(after) - E(before) computed
    from scratch by the faithful global evaluation (so the two-site sequential
    delta E and the conflict-free colourings are right); sites stay in range."""
    rng = np.random.default_rng(2)
    sysd = systems[case_sys]
    n_cells = int(np.prod(N))
    n_sub = len(sysd["occ_to_species"])
    occ = np.zeros(n_cells * n_sub, dtype=np.int32)
    for b in sysd["mutable_sublats"]:
        nocc = sum(1 for s in sysd["occ_to_species"][b] if s >= 0)
        occ[b * n_cells:(b + 1) * n_cells] = rng.integers(0, nocc, n_cells)
    st, sysd, swaps = _state(dev_tables, systems, case_sys, eci_key, N, 1000.0, occ)
    info = st.canonical_set_swaps(swaps)
    assert len(info) == len(swaps) and all(nc >= 4 for _, nc in info)
    e0 = st.energy()
    c0 = st.composition()
    cnt = st.canonical_sweep(4, seed=5)
    e1 = st.energy()
    c1 = st.composition()
    # species counts: sum over the sublattices of one asymmetric unit
    def species_counts(c):
        out = np.zeros(sysd["n_species"], dtype=np.int64)
        for b, row in enumerate(sysd["occ_to_species"]):
            for o, sp in enumerate(row):
                if sp >= 0:
                    out[sp] += c[b, o]
        return out
    assert (species_counts(c0) == species_counts(c1)).all()
    assert 0 < cnt[0].n_accept < cnt[0].n_attempt <= 4 * len(swaps) * n_cells
    assert cnt[0].dE_sum == pytest.approx(e1 - e0, rel=1e-9, abs=1e-7)
    new = st.download_occ()
    assert (new != occ).any()
    for b in range(n_sub):
        nocc = sum(1 for s in sysd["occ_to_species"][b] if s >= 0)
        seg = new[b * n_cells:(b + 1) * n_cells]
        assert seg.min() >= 0 and seg.max() < nocc
    # deterministic: same seed, same trajectory; continuing the stream == one call
    st2, _, _ = _state(dev_tables, systems, case_sys, eci_key, N, 1000.0, occ)
    st2.canonical_set_swaps(swaps)
    st2.canonical_sweep(3, seed=5)
    st2.canonical_sweep(1, seed=5, first_sweep=3)
    assert (st2.download_occ() == new).all()
    st.close()
    st2.close()


def test_canonical_colourings_are_conflict_free(dev_tables, systems):
    """A long translation with t_i = 2 mod 4 needs only the stride-4 colouring;
    nearest-neighbour exchanges need a stride > range + |t| along the hop."""
    occ = np.zeros(16 * 8 * 8, dtype=np.int32)
    st, sysd, _ = _state(dev_tables, systems, "fcc", "eci_sparse", (16, 8, 8), 800.0, occ)
    info = st.canonical_set_swaps([(0, 0, (1, 0, 0)), (0, 0, (10, 3, 5)), (0, 0, (0, 1, -1))])
    (S0, n0), (S1, n1), (S2, n2) = info
    assert S0[0] >= 3 and n1 <= 16
    with pytest.raises(_capi.CmxError):
        st.canonical_set_swaps([(0, 0, (0, 0, 0))])
    with pytest.raises(_capi.CmxError):
        st.canonical_set_swaps([(0, 3, (1, 0, 0))])
    st.close()


def test_canonical_matches_sequential_thermodynamics(dev_tables, systems, oracle):
    """north_star (3) for the canonical ensemble: averages of the parallel pair
    exchanges agree with the reference's sequential any-two-sites swaps
    (oracle: propose_canonical_event restated, reference kernels) within 3 sigma."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    sysd = systems["fcc"]
    eci = sysd["eci_sparse"]
    N = 8
    n_cells = N ** 3
    T = 1200.0
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=3, Rt=np.array(sysd["axes"]["Rt"]), origin=np.array(sysd["axes"]["origin"]))
    sc = oracle.RefClexulator("fcc_default").supercell(N)
    n_runs = 8
    base = np.array([0] * (n_cells // 2) + [1] * (n_cells // 4) + [2] * (n_cells - n_cells // 2 - n_cells // 4),
                    dtype=np.int32)
    ref_e = []
    inits = []
    for run in range(n_runs):
        occ = np.random.default_rng(100 + run).permutation(base).astype(np.int32)
        inits.append(occ.copy())
        out = sc.metropolis_run(1, occ, prim, eci["index"], eci["value"], T, seed=1000 + run,
                                n_steps=60 * n_cells)
        occ = out["occ"]
        es = []
        for k in range(30):
            out = sc.metropolis_run(1, occ, prim, eci["index"], eci["value"], T, seed=5000 + 97 * run + k,
                                    n_steps=4 * n_cells)
            occ = out["occ"]
            g = sc.global_corr(occ)
            es.append(float(np.dot(eci["value"], g[eci["index"]])) / n_cells)
        ref_e.append(np.mean(es))
    st, _, swaps = _state(dev_tables, systems, "fcc", "eci_sparse", (N, N, N), T, inits[0], n_replicas=n_runs)
    for r in range(n_runs):
        st.upload_occ(inits[r], r)
    st.canonical_set_swaps(swaps)
    st.canonical_sweep(40, seed=21)
    ge = np.zeros((n_runs, 30))
    for k in range(30):
        st.canonical_sweep(3, seed=21, first_sweep=40 + 3 * k)
        for r in range(n_runs):
            ge[r, k] = st.energy(r) / n_cells
    for r in range(n_runs):
        assert (np.bincount(st.download_occ(r), minlength=3) == np.bincount(base, minlength=3)).all()
    gpu_e = ge.mean(axis=1)
    se = np.hypot(np.std(ref_e, ddof=1), np.std(gpu_e, ddof=1)) / np.sqrt(n_runs)
    assert abs(np.mean(ref_e) - np.mean(gpu_e)) < 3 * se + 1e-4, (np.mean(ref_e), np.mean(gpu_e), se)
    st.close()


def test_config0_binary_pair_triplet_canonical_temperature_path(dev_tables, systems, oracle):
    """BASELINE configs[0]: FCC binary A-B canonical Metropolis on a pair + triplet basis along a
    cooling path.  The reference ships no such clexulator (SURVEY 8c "Gap"): the basis is the
    synthetic one emitted in the generated-source grammar (make_synthetic_clexulator.py) and
    compiled into oracle/_ref like the reference's own.  x_B = 0.5, every temperature starts
    from the final state of the one before (dependent runs); mean formation energy of the
    parallel pair exchanges against the restated sequential any-two-sites loop driving the
    compiled kernels, 3 sigma over independent runs at every temperature."""
    if oracle is None or not oracle.available("fcc_synthetic"):
        pytest.skip("oracle/_ref has no synthetic clexulator")
    sysd = systems["fcc_syn"]
    eci = sysd["eci"]
    N = 8
    n_cells = N ** 3
    temps = [1500.0, 700.0, 300.0]
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=2, Rt=np.array(sysd["axes"]["Rt"]), origin=np.array(sysd["axes"]["origin"]))
    sc = oracle.RefClexulator("fcc_synthetic").supercell(N)
    n_runs = 8
    base = np.array([0] * (n_cells // 2) + [1] * (n_cells - n_cells // 2), dtype=np.int32)
    inits = [np.random.default_rng(300 + r).permutation(base).astype(np.int32) for r in range(n_runs)]
    ref_e = np.zeros((len(temps), n_runs))
    for run in range(n_runs):
        occ = inits[run].copy()
        for ti, T in enumerate(temps):
            occ = sc.metropolis_run(1, occ, prim, eci["index"], eci["value"], T, seed=7000 + 31 * run + ti,
                                    n_steps=60 * n_cells)["occ"]
            es = []
            for k in range(24):
                occ = sc.metropolis_run(1, occ, prim, eci["index"], eci["value"], T,
                                        seed=9000 + 977 * run + 37 * ti + k, n_steps=4 * n_cells)["occ"]
                g = sc.global_corr(occ)
                es.append(float(np.dot(eci["value"], g[eci["index"]])) / n_cells)
            ref_e[ti, run] = np.mean(es)
    st, _, swaps = _state(dev_tables, systems, "fcc_syn", "eci", (N, N, N), temps[0], inits[0], n_replicas=n_runs)
    for r in range(n_runs):
        st.upload_occ(inits[r], r)
    st.canonical_set_swaps(swaps)
    sweep = 0
    for ti, T in enumerate(temps):
        for r in range(n_runs):
            st.set_conditions(T, None, r)
        st.canonical_sweep(40, seed=33, first_sweep=sweep)
        sweep += 40
        ge = np.zeros((n_runs, 24))
        for k in range(24):
            st.canonical_sweep(3, seed=33, first_sweep=sweep)
            sweep += 3
            for r in range(n_runs):
                ge[r, k] = st.energy(r) / n_cells
        gpu_e = ge.mean(axis=1)
        se = np.hypot(np.std(ref_e[ti], ddof=1), np.std(gpu_e, ddof=1)) / np.sqrt(n_runs)
        assert abs(np.mean(ref_e[ti]) - np.mean(gpu_e)) < 3 * se + 1e-4, (T, np.mean(ref_e[ti]), np.mean(gpu_e), se)
    for r in range(n_runs):
        assert (np.bincount(st.download_occ(r), minlength=2) == np.bincount(base, minlength=2)).all()
    # the path really cools: the mean energy drops from the first to the last temperature
    assert np.mean(ref_e[-1]) < np.mean(ref_e[0])
    st.close()


def test_zro_canonical_matches_sequential_thermodynamics(dev_tables, systems, oracle):
    """BASELINE configs[3] (HCP Zr-O, O/Va on two interstitial sublattices, pairs + triplets +
    quadruplets): the parallel pair exchanges -- swap types restricted to a set of
    translations, coloured conflict-free -- sample the same canonical ensemble as the
    reference's sequential any-two-sites swaps (oracle: propose_canonical_event restated
    around the reference's compiled ZrO kernels): <formation energy> per cell and its
    variance (the heat capacity) within 3 sigma over independent runs, 12^3 cells."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    sysd = systems["zro"]
    eci = sysd["eci"]
    N = (12, 12, 12)
    n_cells = int(np.prod(N))
    n_sub = len(sysd["occ_to_species"])
    T = 900.0
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=sysd["n_species"], Rt=np.array(sysd["axes"]["Rt"]), origin=np.array(sysd["axes"]["origin"]))
    sc = oracle.RefClexulator(sysd["tables"]).supercell(N)
    n_runs, n_eq, n_smp, period = 4, 30, 40, 2
    n_mut = n_cells * len(sysd["mutable_sublats"])
    inits = []
    for run in range(n_runs):
        rng = np.random.default_rng(300 + run)
        occ = np.zeros(n_cells * n_sub, dtype=np.int32)
        flat = np.zeros(n_mut, dtype=np.int32)
        flat[rng.permutation(n_mut)[:n_mut // 4]] = 1          # x_O = 0.25 over the interstitial sites
        for q, b in enumerate(sysd["mutable_sublats"]):
            occ[b * n_cells:(b + 1) * n_cells] = flat[q * n_cells:(q + 1) * n_cells]
        inits.append(occ)
    ref = []
    for run in range(n_runs):
        occ = sc.metropolis_run(1, inits[run], prim, eci["index"], eci["value"], T, seed=1000 + run,
                                n_steps=n_eq * n_mut)["occ"]
        es = []
        for k in range(n_smp):
            occ = sc.metropolis_run(1, occ, prim, eci["index"], eci["value"], T, seed=5000 + 97 * run + k,
                                    n_steps=period * n_mut)["occ"]
            g = sc.global_corr(occ)
            es.append(float(np.dot(eci["value"], g[eci["index"]])) / n_cells)
        ref.append([np.mean(es), np.var(es)])
    st, _, swaps = _state(dev_tables, systems, "zro", "eci", N, T, inits[0], n_replicas=n_runs)
    for r in range(n_runs):
        st.upload_occ(inits[r], r)
    st.canonical_set_swaps(swaps)
    # one device sweep visits n_swap_types * n_cells pairs: match the reference's attempts per pass
    per_sweep = len(swaps) * n_cells
    eq = max(1, round(n_eq * n_mut / per_sweep))
    per = max(1, round(period * n_mut / per_sweep))
    st.canonical_sweep(eq, seed=21)
    ge = np.zeros((n_runs, n_smp))
    for k in range(n_smp):
        st.canonical_sweep(per, seed=21, first_sweep=eq + per * k)
        for r in range(n_runs):
            ge[r, k] = st.energy(r) / n_cells
    for r in range(n_runs):
        assert np.bincount(st.download_occ(r)).tolist() == np.bincount(inits[r]).tolist()
    st.close()
    gpu = np.array([[ge[r].mean(), ge[r].var()] for r in range(n_runs)])
    ref = np.array(ref)
    for q, (name, floor) in enumerate((("formation energy per cell", 2e-4), ("variance of the energy", 0.0))):
        se = np.hypot(ref[:, q].std(ddof=1), gpu[:, q].std(ddof=1)) / np.sqrt(n_runs)
        d = abs(ref[:, q].mean() - gpu[:, q].mean())
        assert d <= 3 * se + floor, f"{name}: reference {ref[:, q].mean():.6g} gpu {gpu[:, q].mean():.6g} (3 sigma {3 * se:.3g})"


@pytest.mark.parametrize("ensemble", ["semigrand", "canonical"])
def test_warp_evaluator_equals_thread_evaluator(dev_tables, systems, ensemble):
    """Wide orbit sets (ZrO: ~700 merged terms per site) are evaluated one site per warp
    (cmx_warp_site_delta; canonical: all colours of a swap type in one cooperative launch
    with grid barriers).  Same random bits and decision rule as the one-site-per-thread
    kernels: identical trajectories and counters, dE sums equal to rounding."""
    from casmcode_clexmonte_b200.potential import semigrand_exchange_table
    sysd = systems["zro"]
    N = (12, 8, 8)
    n_cells = int(np.prod(N))
    rng = np.random.default_rng(9)
    occ = np.zeros(n_cells * len(sysd["occ_to_species"]), dtype=np.int32)
    for b in sysd["mutable_sublats"]:
        occ[b * n_cells:(b + 1) * n_cells] = rng.random(n_cells) < 0.3
    out = []
    for flags in (_capi.CMX_SWEEP_DE_SUM, _capi.CMX_SWEEP_DE_SUM | _capi.CMX_SWEEP_THREAD_GENERIC):
        st, _, swaps = _state(dev_tables, systems, "zro", "eci", N, 900.0, occ, n_replicas=2)
        ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0.3], sysd["n_species"])
        st.set_conditions(900.0, ex, 0)
        st.set_conditions(500.0, ex, 1)
        st.set_sweep_flags(flags)
        if ensemble == "canonical":
            st.canonical_set_swaps(swaps)
            c1 = st.canonical_sweep(2, seed=3)
            c2 = st.canonical_sweep(1, seed=3, first_sweep=2)
        else:
            c1 = st.sgc_sweep(2, seed=3)
            c2 = st.sgc_sweep(1, seed=3, first_sweep=2)
        out.append(([st.download_occ(r) for r in range(2)],
                    [(c1[r].n_attempt + c2[r].n_attempt, c1[r].n_accept + c2[r].n_accept) for r in range(2)],
                    [c1[r].dE_sum + c2[r].dE_sum for r in range(2)]))
        st.close()
    (occ_w, cnt_w, de_w), (occ_t, cnt_t, de_t) = out
    for r in range(2):
        assert (occ_w[r] == occ_t[r]).all()
        assert cnt_w[r] == cnt_t[r] and cnt_w[r][1] > 0
        assert de_w[r] == pytest.approx(de_t[r], rel=1e-10, abs=1e-9)


@pytest.mark.parametrize("N,binary", [((32, 16, 16), False), ((16, 8, 8), True), ((64, 4, 6), False)])
def test_canonical_pair_lut_equals_generic(dev_tables, systems, N, binary):
    """Pair-LUT canonical kernel (two table lookups per swap, the second with the first
    site's new occupant) against the generic term-list kernel: same proposals and random
    bits, dE equal to rounding -> identical trajectories and counters."""
    rng = np.random.default_rng(12)
    n_cells = int(np.prod(N))
    occ = rng.choice(2 if binary else 3, size=n_cells).astype(np.int32)
    out = []
    for flags in (0, _capi.CMX_SWEEP_FORCE_GENERIC):
        st, sysd, swaps = _state(dev_tables, systems, "fcc", "eci_sparse", N, 700.0, occ, n_replicas=2)
        st.set_conditions(1500.0, None, 1)
        st.set_sweep_flags(flags)
        st.canonical_set_swaps(swaps)
        c1 = st.canonical_sweep(3, seed=4)
        c2 = st.canonical_sweep(2, seed=4, first_sweep=3)
        out.append(([st.download_occ(r) for r in range(2)],
                    [(c1[r].n_attempt + c2[r].n_attempt, c1[r].n_accept + c2[r].n_accept) for r in range(2)],
                    [c1[r].dE_sum + c2[r].dE_sum for r in range(2)], st.energy(0)))
        st.close()
    (occ_l, cnt_l, de_l, e_l), (occ_g, cnt_g, de_g, e_g) = out
    for r in range(2):
        assert (occ_l[r] == occ_g[r]).all()
        assert cnt_l[r] == cnt_g[r] and 0 < cnt_l[r][1] < cnt_l[r][0]
        assert de_l[r] == pytest.approx(de_g[r], rel=1e-10, abs=1e-9)
        assert (np.bincount(occ_l[r], minlength=3) == np.bincount(occ, minlength=3)).all()
