"""CPU tests (-m "not gpu"): the oracle against the golden vectors, the table
exporter against the reference's generated sources, and host logic."""
import numpy as np
import pytest

from conftest import CASES, GOLDEN, REFERENCE

from casmcode_clexmonte_b200 import clexulator_tables as CT
from oracle import tables_eval as TE


# ---------------------------------------------------------------------------
# known-answer checks that pin the oracle itself
# ---------------------------------------------------------------------------
def test_documented_event_state_kat(oracle, systems):
    """python/libcasm/clexmonte/_MonteCalculator.py:186-210 documents one KMC event
    state (B_Va_1NN, occ_init [2,1] -> occ_final [1,2]) computed with
    formation_energy_eci.2.json:
      formation_energy_delta_corr = [0,0,0,-0.8333..,0.5892558333..,0,0,0,0]
      dE_final = 1.6666666666666665, Ekra = 0.7375, freq = 1e13,
      rate = 1000704.0785393054          
    -5/6 and 0.707107*5/6 mean: the B atom has 5 more B nearest neighbours than
    the vacancy, and no other vacancy is near.  The printed BITS
    (-0.8333333333333333, not -0.8333333333333334; 1.6666666666666665) are
    reproduced only by 3 B neighbours of the vacancy vs 8 of the B atom,
    evaluated site after site in the reference's association order
    (3/6 - 8/6), so this also pins the sequential two-site semantics.  Build
    such a configuration and check the oracle (reference kernels + our
    restatement of occ_delta) bit for bit; the rate pins KB."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    clex = oracle.RefClexulator("fcc_default")
    N = 6
    sc = clex.supercell(N)
    nn = clex.cells[1:13]  # the 12 nearest-neighbour cells of the prim neighbor list

    def site(c):
        return int(c[0] % N + N * (c[1] % N + N * (c[2] % N)))

    X = np.array([2, 2, 2])
    Y = X + nn[6]
    nbX = {site(X + d) for d in nn}
    exclusive = [site(Y + d) for d in nn if site(Y + d) not in nbX and site(Y + d) != site(X)]
    common = [site(Y + d) for d in nn if site(Y + d) in nbX]
    assert len(exclusive) == 7 and len(common) == 4
    occ = np.zeros(sc.n_sites, dtype=np.int32)
    occ[site(X)] = 2  # Va
    occ[site(Y)] = 1  # B
    for l in common[:2] + exclusive[:5]:
        occ[l] = 1
    eci = systems["fcc"]["eci_2"]
    e, dcorr = sc.occ_delta_value(occ, [site(X), site(Y)], [1, 2], eci["index"], eci["value"],
                                  return_dcorr=True)
    expect = np.array([0.0, 0.0, 0.0, -0.8333333333333333, 0.5892558333333333, 0.0, 0.0, 0.0, 0.0])
    idx = np.array(eci["index"])
    assert (dcorr[idx] == expect[idx]).all()
    assert e == 1.6666666666666665
    # rate = freq * exp(-dE_activated / (KB T)), dE_activated clamped up to dE_final
    # (BaseMonteEventData.cc:148-155); T = 1200 K
    dEa = max(0.5 * e + 0.7375, e)
    assert dEa == 1.6666666666666665
    rate = 1e13 * np.exp(-dEa / (oracle.KB * 1200.0))
    assert rate == pytest.approx(1000704.0785393054, rel=1e-12)


def test_delta_equals_difference_of_global(oracle):
    """The neighbor-list convention: sum over cells of the
    global contribution after - before == delta point corr."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    for name, N, mut, nocc, tol in (("fcc_default", 6, [0], 3, 5e-14), ("zro", 8, [2, 3], 2, 5e-13),
                                    ("fcc_synthetic", 6, [0], 2, 5e-14)):
        if not oracle.available(name):
            pytest.skip(f"oracle/_ref has no {name}")
        sc = oracle.RefClexulator(name).supercell(N)
        rng = np.random.default_rng(5)
        occ = np.zeros(sc.n_sites, dtype=np.int32)
        for b in mut:
            occ[b * sc.n_cells:(b + 1) * sc.n_cells] = rng.integers(0, nocc, sc.n_cells)
        g0 = sc.global_corr(occ)
        for _ in range(10):
            l = mut[rng.integers(len(mut))] * sc.n_cells + rng.integers(sc.n_cells)
            new = (occ[l] + 1 + rng.integers(nocc - 1)) % nocc
            d = sc.delta_corr(occ, l, new)
            occ2 = occ.copy()
            occ2[l] = new
            assert np.abs((sc.global_corr(occ2) - g0) - d).max() < tol


@pytest.mark.parametrize("case", list(CASES))
def test_oracle_reproduces_golden(oracle, systems, load_vectors, case):
    """The live oracle regenerates the committed golden vectors bit for bit."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    sysname, _ = CASES[case]
    if not oracle.available(systems[sysname]["tables"]):
        pytest.skip(f"oracle/_ref has no {systems[sysname]['tables']}")
    v = load_vectors(case)
    sc = oracle.RefClexulator(systems[sysname]["tables"]).supercell(tuple(v["N"]))
    occ = v["occ"]
    for q in range(0, len(v["l"]), 5):
        assert (sc.delta_corr(occ, v["l"][q], v["new_occ"][q]) == v["delta_corr"][q]).all()
        assert (sc.point_corr(occ, v["l"][q]) == v["point_corr"][q]).all()
        assert (sc.cell_corr(occ, v["cells"][q]) == v["cell_corr"][q]).all()
        assert sc.occ_delta_value(occ, v["l2"][q], v["new_occ2"][q], v["eci_index"], v["eci_value"]) \
            == v["delta_e_2"][q]
    assert (sc.global_corr(occ) == v["global_corr"]).all()


def test_libstdcxx_rng_golden(oracle, load_vectors):
    if oracle is None:
        pytest.skip("oracle harness not built")
    v = load_vectors("rng")
    oi, orl, oraw = oracle.rng_stream(int(v["seed"]), v["kinds"], v["int_max"], v["real_max"])
    assert (oi == v["out_int"]).all() and (orl == v["out_real"]).all() and (oraw == v["out_raw"]).all()


# ---------------------------------------------------------------------------
# exporter: flat tables vs golden vectors (always) and vs the sources (here)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("case", list(CASES))
def test_tables_evaluate_to_golden(systems, load_tables, load_vectors, case):
    """Evaluating the committed tables in the reference's operation order on the
    CPU gives the reference kernels' values bit for bit."""
    sysname, _ = CASES[case]
    t = load_tables(systems[sysname]["tables"])
    v = load_vectors(case)
    N = tuple(int(x) for x in v["N"])
    occ = v["occ"]
    step = 8 if sysname == "zro" else 4
    for q in range(0, len(v["l"]), step):
        assert (TE.delta_corr(t, N, occ, int(v["l"][q]), int(v["new_occ"][q])) == v["delta_corr"][q]).all()
        assert (TE.point_corr(t, N, occ, int(v["l"][q])) == v["point_corr"][q]).all()
        assert (TE.cell_corr(t, N, occ, int(v["cells"][q])) == v["cell_corr"][q]).all()


def test_local_tables_evaluate_to_golden(load_tables, load_vectors):
    v = load_vectors("local")
    N = tuple(int(x) for x in v["N"])
    for ev in ("A_Va_1NN", "B_Va_1NN"):
        for k in range(6):
            name = f"fcc_{ev}_{k}"
            t = load_tables(name)
            assert t.is_local
            for q in range(0, len(v["cells"]), 4):
                assert (TE.cell_corr(t, N, v["occ"], int(v["cells"][q])) == v[name][q]).all()


@pytest.mark.skipif(not REFERENCE.exists(), reason="/root/reference not present")
def test_exporter_matches_committed_tables(load_tables):
    """Re-export from the reference sources and compare with tests/golden/tables."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    # only need SOURCES; avoid running main()
    src = (GOLDEN / "make_golden.py").read_text()
    assert "SOURCES" in src
    data = REFERENCE / "tests/unit/clexmonte/data"
    pairs = {
        "fcc_default": data / "FCC_binary_vacancy/basis_sets/bset.default/FCC_binary_vacancy_Clexulator_default.cc",
        "zro": data / "Clex_ZrO_Occ/basis_sets/bset.formation_energy/ZrO_Clexulator_formation_energy.cc",
        "fcc_A_Va_1NN_3": data / "FCC_binary_vacancy/basis_sets/bset.A_Va_1NN/3/FCC_binary_vacancy_Clexulator_A_Va_1NN_3.cc",
    }
    for name, path in pairs.items():
        fresh = CT.parse_clexulator_source(path, name=name)
        old = load_tables(name)
        for k in CT.ClexulatorTables._ARRAYS:
            assert np.array_equal(getattr(fresh, k), getattr(old, k)), (name, k)
        assert fresh.nlist_size == old.nlist_size and fresh.corr_size == old.corr_size


def test_synthetic_source_exports_to_committed_tables(load_tables, tmp_path):
    """The synthetic FCC binary pair + triplet basis (SURVEY 8c "Gap", BASELINE configs[0]):
    the emitter writes the generated-source grammar, the exporter reads it back into exactly
    the committed tables -- 1 + 12 + 6 pair and 24 triplet clusters around a site, divisors
    6 / 3 / 8.  Runs anywhere (no reference, no oracle build needed)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_synthetic_clexulator", GOLDEN / "make_synthetic_clexulator.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    src = tmp_path / f"{mod.NAME}.cc"
    src.write_text(mod.emit())
    fresh = CT.parse_clexulator_source(src, name="fcc_synthetic")
    old = load_tables("fcc_synthetic")
    for k in CT.ClexulatorTables._ARRAYS:
        assert np.array_equal(getattr(fresh, k), getattr(old, k)), k
    assert (fresh.nlist_size, fresh.corr_size, fresh.n_point_corr) == (19, 5, 1)
    # delta function of the triplet orbit: 24 two-factor terms under one divisor
    g = fresh.delta_gbeg[4]
    assert fresh.group_div[g] == 8.0
    terms = range(fresh.elem_tbeg[fresh.group_ebeg[g]], fresh.elem_tbeg[fresh.group_ebeg[g + 1]])
    assert len(terms) == 24 and all(fresh.term_fbeg[t + 1] - fresh.term_fbeg[t] == 2 for t in terms)


def test_table_shapes(load_tables):
    fcc = load_tables("fcc_default")
    assert (fcc.nlist_size, fcc.corr_size, fcc.n_point_corr) == (19, 9, 1)
    assert fcc.nlist_len == 19 and fcc.n_sublat == 1 and fcc.max_occ == 3 and fcc.n_func == 2
    # literals are kept verbatim (6 digits, not 1/sqrt(2))
    assert 0.707107 in set(np.round(fcc.term_coef, 12))
    zro = load_tables("zro")
    assert (zro.nlist_size, zro.corr_size, zro.n_point_corr) == (225, 74, 2)
    assert zro.nlist_len == 226 and list(zro.nlist_sublat) == [2, 3] and list(zro.n_occ) == [1, 1, 2, 2]
    loc = load_tables("fcc_A_Va_1NN_0")
    assert (loc.nlist_size, loc.corr_size, loc.n_point_corr) == (50, 9, 50) and loc.is_local
    # work statistics used for the roofline bookkeeping
    w = fcc.delta_work(0, [1, 2, 3, 4, 5])
    assert w["neighbors"] == 12
    w = fcc.delta_work(0, range(9))
    assert w["neighbors"] == 18


def test_parser_rejects_unknown_grammar():
    good = "(occ_func_0_0(1) + occ_func_0_0(2)) / 2."
    CT.canonicalize(CT.parse_expression(good))
    for bad in ("pow(occ_func_0_0(1), 2)", "occ_func_0_0(1) * 0.5 * occ_func_0_0(2)",
                "(occ_func_0_0(1) - occ_func_0_0(2)) / 2.", "((occ_func_0_0(1) + (occ_func_0_0(2) + occ_func_0_0(3)))) / x"):
        with pytest.raises(CT.ClexulatorParseError):
            CT.canonicalize(CT.parse_expression(bad))
    with pytest.raises(CT.ClexulatorParseError):
        CT.parse_clexulator_source("int main() { return 0; }\n")


def test_canonical_form_equals_ast():
    """The canonical (flat) form evaluates exactly like the C++ expression tree."""
    rng = np.random.default_rng(0)
    exprs = [
        "(occ_func_0_0(0) * occ_func_0_0(9) + occ_func_0_0(10) * occ_func_0_0(0)) / 6.",
        "((0.707107 * occ_func_0_0(0) * occ_func_0_1(9) + 0.707107 * occ_func_0_1(0) * occ_func_0_0(9)) + "
        "(0.707107 * occ_func_0_0(10) * occ_func_0_1(0) + 0.707107 * occ_func_0_1(10) * occ_func_0_0(0))) / 6.",
        "(m_occ_func_0_0[occ_f] - m_occ_func_0_0[occ_i]) * (0.707107 * occ_func_0_1(9) + 0.707107 * occ_func_0_1(4)) / 6. + "
        "(m_occ_func_0_1[occ_f] - m_occ_func_0_1[occ_i]) * (0.707107 * occ_func_0_0(9) + 0.707107 * occ_func_0_0(4)) / 6.",
        "(m_occ_func_0_0[occ_f] - m_occ_func_0_0[occ_i])",
        "(m_occ_func_0_1[occ_f] - m_occ_func_0_1[occ_i]) * (1) / 2.",
        "1", "occ_func_0_1(0)",
    ]
    for e in exprs:
        ast = CT.parse_expression(e)
        can = CT.canonicalize(ast)
        for _ in range(20):
            vals = rng.uniform(-1.3, 1.3, size=(2, 2, 32))
            tab = rng.uniform(-1.3, 1.3, size=(2, 2, 3))
            of = lambda b, f, n: vals[b, f, n]
            mo = lambda b, f, o: tab[b, f, o]
            a = CT.eval_ast(ast, of, mo, 0, 2)
            c = CT.eval_canonical(can, of, mo, 0, 2)
            assert a == c


def test_read_eci_formats(systems):
    idx, val = CT.read_eci([[1, -0.1], [2, 0.3]])
    assert list(idx) == [1, 2] and list(val) == [-0.1, 0.3]
    dense = {"orbits": [{"cluster_functions": [{"linear_function_index": 0}, ]},
                        {"cluster_functions": [{"linear_function_index": 1, "eci": 0.5}]}]}
    idx, val = CT.read_eci(dense)
    assert list(idx) == [1] and list(val) == [0.5]
    with pytest.raises(ValueError):
        CT.read_eci([[99, 1.0]], corr_size=9)
    assert systems["fcc"]["eci_sparse"]["index"] == [1, 2, 3, 4, 5]


def test_semigrand_exchange_table(systems):
    from casmcode_clexmonte_b200.potential import dparam_dmol, semigrand_exchange_table
    ax = systems["fcc"]["axes"]
    Rt = dparam_dmol(ax["origin"], ax["end_members"])
    np.testing.assert_allclose(Rt, np.array([[-1, 2, -1], [-1, -1, 2]]) / 3.0, atol=1e-15)
    ex = semigrand_exchange_table(systems["fcc"]["occ_to_species"], Rt, [0.3, -0.2], 3)
    # A -> B changes x_a by +1: exch = mu_a
    assert ex[0, 0, 1] == pytest.approx(0.3) and ex[0, 0, 2] == pytest.approx(-0.2)
    assert ex[0, 1, 0] == pytest.approx(-0.3) and ex[0, 1, 2] == pytest.approx(-0.5)
    with pytest.raises(ValueError):
        semigrand_exchange_table(systems["fcc"]["occ_to_species"], Rt, [0.3], 3)


def test_point_pair_delta_e_depends_on_shell_counts_only(oracle, systems, load_tables):
    """The premise of the count-table sweep kernels, checked on the reference's own generated
    kernels: with the dense FCC ECI (points + 1NN + 2NN pairs) the delta E of a site flip is a
    function of (occupant, proposal, species counts of the first shell, species counts of the
    second shell) -- two arrangements with the same counts give the same delta E (to rounding),
    whatever the rest of the box holds -- and with the sparse ECI (points + 1NN pairs) of the
    first-shell counts alone."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    N = 8
    sc = oracle.RefClexulator("fcc_default").supercell(N)
    nbr = np.asarray(load_tables("fcc_default").nbr).reshape(-1, 4)
    shell1, shell2 = nbr[1:13, :3], nbr[13:19, :3]
    assert len(shell1) == 12 and len(shell2) == 6

    def site(c):
        return int((c[0] % N) + N * ((c[1] % N) + N * (c[2] % N)))

    c0 = np.array([4, 3, 5])
    l0 = site(c0)
    l1 = [site(c0 + d) for d in shell1]
    l2 = [site(c0 + d) for d in shell2]
    near = set(l1) | set(l2) | {l0}
    rng = np.random.default_rng(17)

    def arrangement(nB1, nV1, nB2, nV2, oi):
        occ = rng.integers(0, 3, sc.n_sites).astype(np.int32)      # the rest of the box: anything
        for ls, nB, nV in ((l1, nB1, nV1), (l2, nB2, nV2)):
            vals = np.array([1] * nB + [2] * nV + [0] * (len(ls) - nB - nV), dtype=np.int32)
            occ[ls] = rng.permutation(vals)
        occ[l0] = oi
        return occ

    for eci_key, second_shell_matters in (("eci_full", True), ("eci_sparse", False)):
        eci = systems["fcc"][eci_key]
        scale = float(np.abs(eci["value"]).sum())
        seen_second = False
        for _ in range(60):
            nV1 = int(rng.integers(0, 13))
            nB1 = int(rng.integers(0, 13 - nV1))
            nV2 = int(rng.integers(0, 7))
            nB2 = int(rng.integers(0, 7 - nV2))
            oi = int(rng.integers(0, 3))
            of = (oi + 1 + int(rng.integers(0, 2))) % 3
            dE = [sc.occ_delta_value(arrangement(nB1, nV1, nB2, nV2, oi), [l0], [of], eci["index"], eci["value"])
                  for _ in range(3)]
            assert max(dE) - min(dE) <= 1e-12 * scale, (eci_key, nB1, nV1, nB2, nV2, oi, of, dE)
            # a different second-shell count changes dE exactly when 2NN pairs carry coefficients
            if nB2 + nV2 < 6:
                other = sc.occ_delta_value(arrangement(nB1, nV1, nB2 + 1, nV2, oi), [l0], [of], eci["index"], eci["value"])
                if abs(other - dE[0]) > 1e-9 * scale:
                    seen_second = True
                if not second_shell_matters:
                    assert abs(other - dE[0]) <= 1e-12 * scale
        assert seen_second == second_shell_matters
    assert len(near) == 19       # SURVEY 8d's 19-site neighbourhood
