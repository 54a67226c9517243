"""GPU tests of the device-side samplers and the sampled run (csrc/cmx_sampler.cu).

The series rows must equal the reference's state sampling functions
(monte_calculator/sampling_functions.cc:37-288) evaluated on the same occupation:
checked against the live oracle (the reference's generated kernels + the restated
SemiGrandCanonicalPotential::per_supercell) when it is built, and always against
the per-replica entry points cmx_energy / cmx_composition / cmx_global_corr, whose
own parity is pinned by test_gpu_parity.py."""
import numpy as np
import pytest

from casmcode_clexmonte_b200 import _capi
from casmcode_clexmonte_b200.potential import (canonical_swap_types, mol_composition, param_composition,
                                               semigrand_exchange_table, semigrand_potential_per_supercell)

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


def _o2s(sysd, max_occ):
    o2s = np.full((len(sysd["occ_to_species"]), max_occ), -1, dtype=np.int32)
    for b, row in enumerate(sysd["occ_to_species"]):
        o2s[b, :len(row)] = row
    return o2s


def _state(dev_tables, systems, case_sys, eci_key, N, conds, seed=3):
    """conds: list of (T, mu) per replica."""
    sysd = systems[case_sys]
    st = _capi.State(dev_tables(sysd["tables"]), N, len(conds))
    eci = sysd[eci_key]
    st.set_eci(eci["index"], eci["value"])
    st.set_occupants(sysd["sublat_to_asym"], _o2s(sysd, st.tables.host.max_occ), sysd["n_species"])
    for r, (T, mu) in enumerate(conds):
        ex = semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], mu, sysd["n_species"])
        st.set_conditions(T, ex, r)
    st.randomize(seed)
    return st, sysd


def _expected_row(st, sysd, r, mu, n_cells):
    e = st.energy(r)
    n = mol_composition(st.composition(r), sysd["occ_to_species"], sysd["n_species"], n_cells)
    x = param_composition(n, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    pot = semigrand_potential_per_supercell(e, n, sysd["axes"]["origin"], sysd["axes"]["Rt"], mu, n_cells)
    return e / n_cells, pot / n_cells, n, x


@pytest.mark.parametrize("case_sys,eci_key,N,with_corr", [
    ("fcc", "eci_sparse", (32, 16, 8), False),    # pair-LUT model: one fused streaming pass
    ("fcc", "eci_sparse", (32, 16, 8), True),
    ("fcc", "eci_full", (16, 8, 8), True),        # generic model: faithful global correlations
    ("fcc", "eci_sparse", (12, 12, 12), False),   # N0 % 16 != 0
    ("zro", "eci", (8, 8, 8), False),
])
def test_sampled_run_equals_sweeps_plus_direct_evaluation(dev_tables, systems, oracle, case_sys, eci_key, N,
                                                          with_corr):
    """cmx_sweep_run == the same sweeps issued call by call, and every series row ==
    the sampling functions evaluated on the occupation at that point."""
    n_par = len(systems[case_sys]["axes"]["end_members"])
    conds = [(900.0, [0.2, -0.1][:n_par]), (600.0, [-0.3, 0.4][:n_par]), (1500.0, [0.0, 0.0][:n_par])]
    a, sysd = _state(dev_tables, systems, case_sys, eci_key, N, conds)
    b, _ = _state(dev_tables, systems, case_sys, eci_key, N, conds)
    n_cells = int(np.prod(N))
    sm = _capi.Sampler(a, 8, sysd["axes"]["origin"], sysd["axes"]["Rt"], with_corr=with_corr)
    for r, (_, mu) in enumerate(conds):
        sm.set_param_chem_pot(mu, r)
    assert sm.n_quantities == 2 + sysd["n_species"] + n_par + (a.tables.host.corr_size if with_corr else 0)
    cnt = sm.run(3, 2, seed=11)
    assert sm.n_samples == 3
    tot = [0] * len(conds)
    rows = []
    for k in range(3):
        c = b.sgc_sweep(2, seed=11, first_sweep=2 * k)
        rows.append([_expected_row(b, sysd, r, conds[r][1], n_cells) + ((b.global_corr(r) / n_cells,) if with_corr else ())
                     for r in range(len(conds))])
        for r in range(len(conds)):
            tot[r] += c[r].n_accept
    for r in range(len(conds)):
        assert (a.download_occ(r) == b.download_occ(r)).all()
        assert cnt[r].n_accept == tot[r] and cnt[r].n_attempt == 6 * n_cells * len(sysd["mutable_sublats"])
        ser = sm.series(r)
        for k in range(3):
            e, pot, n, x = rows[k][r][:4]
            assert ser["clex.formation_energy"][k] == pytest.approx(e, rel=1e-12, abs=1e-14)
            assert ser["potential_energy"][k] == pytest.approx(pot, rel=1e-12, abs=1e-14)
            np.testing.assert_allclose(ser["mol_composition"][k], n, rtol=1e-15, atol=0)
            np.testing.assert_allclose(ser["param_composition"][k], x, rtol=1e-13, atol=1e-15)
            if with_corr:
                np.testing.assert_allclose(ser["corr"][k], rows[k][r][4], rtol=1e-15, atol=0)
    if oracle is not None and case_sys == "fcc":
        # the reference's kernels + restated potential on the final occupation
        prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                    n_species=3, Rt=np.array(sysd["axes"]["Rt"]), origin=np.array(sysd["axes"]["origin"]))
        sc = oracle.RefClexulator("fcc_default").supercell(N)
        eci = sysd[eci_key]
        for r in range(len(conds)):
            occ = a.download_occ(r)
            ref, _ = sc.potential_per_supercell(occ, prim, eci["index"], eci["value"], np.array(conds[r][1]))
            assert sm.series(r)["potential_energy"][-1] == pytest.approx(ref / n_cells, rel=1e-11, abs=1e-13)
            if with_corr:
                np.testing.assert_allclose(sm.series(r)["corr"][-1], sc.global_corr(occ) / n_cells, rtol=1e-12, atol=1e-14)
    # the series is full at `capacity`; reset starts over
    sm.run(5, 0, seed=11, first_sweep=6)
    with pytest.raises(_capi.CmxError):
        sm.sample()
    sm.reset()
    sm.sample()
    assert sm.n_samples == 1
    assert sm.series(0)["potential_energy"][0] == pytest.approx(rows[2][0][1], rel=1e-12)
    sm.close()
    a.close()
    b.close()


def test_canonical_sampled_run(dev_tables, systems):
    """Canonical ensemble (mu = 0: potential_energy == formation energy,
    CanonicalCalculator.cc:126-133): composition is conserved along the series and the
    run equals cmx_canonical_sweep issued call by call."""
    sysd = systems["zro"]
    N = (8, 8, 8)
    n_cells = 512
    rng = np.random.default_rng(4)
    occ = np.zeros(n_cells * len(sysd["occ_to_species"]), dtype=np.int32)
    for bsub in sysd["mutable_sublats"]:
        occ[bsub * n_cells:(bsub + 1) * n_cells] = rng.random(n_cells) < 0.25
    states = []
    for _ in range(2):
        st = _capi.State(dev_tables(sysd["tables"]), N, 2)
        st.set_eci(sysd["eci"]["index"], sysd["eci"]["value"])
        o2s = _o2s(sysd, st.tables.host.max_occ)
        st.set_occupants(sysd["sublat_to_asym"], o2s, sysd["n_species"])
        for r in range(2):
            st.set_conditions(700.0 + 300.0 * r, None, r)
            st.upload_occ(occ, r)
        st.canonical_set_swaps(canonical_swap_types(st.tables.host, sysd["sublat_to_asym"], o2s.tolist(), N))
        states.append(st)
    a, b = states
    sm = _capi.Sampler(a, 4, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    cnt = sm.run(4, 2, seed=5, ensemble="canonical")
    acc = [0, 0]
    for k in range(4):
        c = b.canonical_sweep(2, seed=5, first_sweep=2 * k)
        for r in range(2):
            acc[r] += c[r].n_accept
            ser = sm.series(r)
            assert ser["clex.formation_energy"][k] == pytest.approx(b.energy(r) / n_cells, rel=1e-12)
            assert ser["potential_energy"][k] == ser["clex.formation_energy"][k]
    for r in range(2):
        assert (a.download_occ(r) == b.download_occ(r)).all()
        assert cnt[r].n_accept == acc[r] > 0
        ser = sm.series(r)
        assert (ser["mol_composition"] == ser["mol_composition"][0]).all()
        n0 = mol_composition(a.composition(r), sysd["occ_to_species"], sysd["n_species"], n_cells)
        np.testing.assert_allclose(ser["mol_composition"][0], n0, rtol=1e-15)
    sm.close()
    a.close()
    b.close()


def test_sampler_errors(dev_tables, systems):
    sysd = systems["fcc"]
    st = _capi.State(dev_tables("fcc_default"), (16, 8, 8))
    with pytest.raises(_capi.CmxError):
        _capi.Sampler(st, 4, sysd["axes"]["origin"], sysd["axes"]["Rt"])      # no ECI
    st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
    with pytest.raises(_capi.CmxError):
        _capi.Sampler(st, 4, sysd["axes"]["origin"], sysd["axes"]["Rt"])      # occupants unknown
    st.set_occupants(sysd["sublat_to_asym"], _o2s(sysd, 3), 3)
    with pytest.raises(_capi.CmxError):
        _capi.Sampler(st, 0, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    sm = _capi.Sampler(st, 2, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    with pytest.raises(_capi.CmxError):
        sm.run(1, 1, seed=1)                                                  # conditions not set
    st.set_conditions(800.0, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], [0, 0], 3))
    with pytest.raises(_capi.CmxError):
        sm.run(3, 1, seed=1)                                                  # exceeds the capacity
    with pytest.raises(_capi.CmxError):
        sm.run(1, 1, seed=1, ensemble="canonical")                           # no swap types
    with pytest.raises(_capi.CmxError):
        sm.set_param_chem_pot([0.0, 0.0], replica=3)
    sm.close()
    st.close()


def test_heat_capacity_and_susceptibility_match_reference_run(dev_tables, systems, oracle):
    """north_star (3), through the sampler: <potential_energy>, <param_composition>,
    heat_capacity and param_susc of checkerboard runs agree with the reference's
    sequential Metropolis (oracle: reference kernels in the restated loop, sampled every
    pass, analysis functions of analysis_functions.cc:43-173) within 3 sigma over
    independent runs."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    sysd = systems["fcc"]
    eci = sysd["eci_sparse"]
    N, T, mu = 8, 1500.0, np.array([0.3, -0.4])
    n_cells = N ** 3
    n_runs, n_samp = 8, 150
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=3, Rt=np.array(sysd["axes"]["Rt"]), origin=np.array(sysd["axes"]["origin"]))
    sc = oracle.RefClexulator("fcc_default").supercell(N)
    c_heat, c_susc = _capi.KB * T * T / n_cells, _capi.KB * T / n_cells

    def stats(e, x):
        return np.array([e.mean(), x[:, 0].mean(), x[:, 1].mean(), e.var() / c_heat,
                         x[:, 0].var() / c_susc, x[:, 1].var() / c_susc])

    ref = []
    for run in range(n_runs):
        occ = np.random.default_rng(100 + run).integers(0, 3, n_cells).astype(np.int32)
        occ = sc.metropolis_run(0, occ, prim, eci["index"], eci["value"], T, seed=1000 + run,
                                n_steps=100 * n_cells, param_chem_pot=mu)["occ"]
        e, x = np.zeros(n_samp), np.zeros((n_samp, 2))
        for k in range(n_samp):
            occ = sc.metropolis_run(0, occ, prim, eci["index"], eci["value"], T, seed=5000 + 977 * run + k,
                                    n_steps=2 * n_cells, param_chem_pot=mu)["occ"]
            e[k] = sc.potential_per_supercell(occ, prim, eci["index"], eci["value"], mu)[0] / n_cells
            n = np.bincount(occ, minlength=3) / n_cells
            x[k] = param_composition(n, prim["origin"], prim["Rt"])
        ref.append(stats(e, x))
    ref = np.array(ref)
    st, _ = _state(dev_tables, systems, "fcc", "eci_sparse", (N, N, N), [(T, mu)] * n_runs, seed=8)
    sm = _capi.Sampler(st, n_samp, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    for r in range(n_runs):
        sm.set_param_chem_pot(mu, r)
    st.sgc_sweep(100, seed=21)
    sm.run(n_samp, 2, seed=21, first_sweep=100)
    gpu = []
    for r in range(n_runs):
        ser = sm.series(r)
        gpu.append(stats(ser["potential_energy"], ser["param_composition"]))
        an = sm.analysis(r)
        assert an["heat_capacity"] == pytest.approx(gpu[-1][3], rel=1e-9)
        assert an["param_susc"][0, 0] == pytest.approx(gpu[-1][4], rel=1e-9)
    gpu = np.array(gpu)
    se = np.hypot(ref.std(axis=0, ddof=1), gpu.std(axis=0, ddof=1)) / np.sqrt(n_runs)
    diff = np.abs(ref.mean(axis=0) - gpu.mean(axis=0))
    names = ["<potential_energy>", "<x_a>", "<x_b>", "heat_capacity", "param_susc(a,a)", "param_susc(b,b)"]
    for q, name in enumerate(names):
        assert diff[q] < 3 * se[q] + 1e-4 * max(1.0, abs(ref.mean(axis=0)[q])), \
            (name, ref.mean(axis=0)[q], gpu.mean(axis=0)[q], se[q])
    sm.close()
    st.close()


def test_binary_canonical_temperature_path_matches_reference(dev_tables, systems, oracle):
    """BASELINE configs[0]: FCC A-B canonical Metropolis (vacancy fraction 0) on a
    4 000-site box along a temperature path.  The conditions of the path are the replicas
    of ONE state (a run series with independent runs, run/functions.hh:83-166), sampled
    on the device; <formation energy> and the heat capacity at every temperature agree
    with the reference's sequential any-two-sites swaps (oracle: propose_canonical_event
    restated, reference kernels) within 3 sigma over independent runs."""
    if oracle is None:
        pytest.skip("oracle/_ref not built")
    sysd = systems["fcc"]
    eci = sysd["eci_sparse"]
    N = (20, 20, 10)                      # 4 000 sites, as the 10^3 conventional FCC box
    n_cells = int(np.prod(N))
    temps = [2000.0, 1100.0, 600.0]
    n_runs, n_samp = 8, 100
    prim = dict(sublat_to_asym=sysd["sublat_to_asym"], occ_to_species=sysd["occ_to_species"],
                n_species=3, Rt=np.array(sysd["axes"]["Rt"]), origin=np.array(sysd["axes"]["origin"]))
    sc = oracle.RefClexulator("fcc_default").supercell(N)
    base = np.array([0] * (n_cells // 2) + [1] * (n_cells - n_cells // 2), dtype=np.int32)
    inits = [np.random.default_rng(40 + q).permutation(base).astype(np.int32) for q in range(n_runs)]
    ref = np.zeros((len(temps), n_runs, 2))
    for ti, T in enumerate(temps):
        c_heat = _capi.KB * T * T / n_cells
        for q in range(n_runs):
            occ = sc.metropolis_run(1, inits[q], prim, eci["index"], eci["value"], T, seed=900 + 17 * ti + q,
                                    n_steps=40 * n_cells)["occ"]
            es = np.zeros(n_samp)
            for k in range(n_samp):
                occ = sc.metropolis_run(1, occ, prim, eci["index"], eci["value"], T,
                                        seed=7000 + 1000 * ti + 100 * q + k, n_steps=2 * n_cells)["occ"]
                g = sc.global_corr(occ)
                es[k] = float(np.dot(eci["value"], g[eci["index"]])) / n_cells
            ref[ti, q] = es.mean(), es.var() / c_heat
    R = len(temps) * n_runs
    st = _capi.State(dev_tables("fcc_default"), N, R)
    st.set_eci(eci["index"], eci["value"])
    o2s = _o2s(sysd, 3)
    st.set_occupants(sysd["sublat_to_asym"], o2s, 3)
    for ti, T in enumerate(temps):
        for q in range(n_runs):
            st.set_conditions(T, None, ti * n_runs + q)
            st.upload_occ(inits[q], ti * n_runs + q)
    st.canonical_set_swaps(canonical_swap_types(st.tables.host, sysd["sublat_to_asym"], o2s.tolist(), N))
    sm = _capi.Sampler(st, n_samp, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    st.canonical_sweep(40, seed=77)
    sm.run(n_samp, 2, seed=77, first_sweep=40, ensemble="canonical")
    for ti, T in enumerate(temps):
        gpu = np.zeros((n_runs, 2))
        for q in range(n_runs):
            r = ti * n_runs + q
            ser = sm.series(r)
            assert (ser["mol_composition"][:, 2] == 0).all()          # no vacancies appear
            assert (ser["mol_composition"][:, 1] == ser["mol_composition"][0, 1]).all()
            gpu[q] = ser["clex.formation_energy"].mean(), sm.analysis(r)["heat_capacity"]
        se = np.hypot(ref[ti].std(axis=0, ddof=1), gpu.std(axis=0, ddof=1)) / np.sqrt(n_runs)
        diff = np.abs(ref[ti].mean(axis=0) - gpu.mean(axis=0))
        assert diff[0] < 3 * se[0] + 2e-4, ("energy", T, ref[ti].mean(axis=0), gpu.mean(axis=0), se)
        assert diff[1] < 3 * se[1], ("heat capacity", T, ref[ti].mean(axis=0), gpu.mean(axis=0), se)
    sm.close()
    st.close()


def test_run_series_batches_the_conditions_path(dev_tables, systems):
    """run_series with independent runs = the states of an incremental conditions path as
    replicas of one device state; every state equals the same run done on its own, and a
    dependent series hands the final configuration to the next state."""
    from casmcode_clexmonte_b200.run_series import run_series
    sysd = systems["fcc"]
    eci = sysd["eci_sparse"]
    N = (16, 8, 8)
    occ = np.random.default_rng(2).integers(0, 3, int(np.prod(N))).astype(np.int32)
    init = {"temperature": 600.0, "param_chem_pot": [-0.5, 0.0]}
    inc = {"temperature": 300.0, "param_chem_pot": [0.25, 0.0]}
    kw = dict(n_equilibration_passes=5, n_samples=6, sample_period=2, seed=9)
    res = run_series(dev_tables("fcc_default"), N, sysd, eci["index"], eci["value"], init, inc, 4, occ, **kw)
    assert [r["conditions"]["temperature"] for r in res] == [600.0, 900.0, 1200.0, 1500.0]
    assert all(r["potential_energy"]["n_samples"] == 6 and 0 < r["acceptance_rate"] < 1 for r in res)
    assert res[0]["acceptance_rate"] < res[-1]["acceptance_rate"]          # hotter: more acceptances
    # replica 0 of the batch == the one-state series (replica index enters the RNG counter)
    one = run_series(dev_tables("fcc_default"), N, sysd, eci["index"], eci["value"], init, inc, 1, occ, **kw)
    assert one[0]["potential_energy"]["mean"] == res[0]["potential_energy"]["mean"]
    assert (one[0]["final_occupation"] == res[0]["final_occupation"]).all()
    assert one[0]["heat_capacity"] == res[0]["heat_capacity"]
    dep = run_series(dev_tables("fcc_default"), N, sysd, eci["index"], eci["value"], init, inc, 2, occ,
                     dependent_runs=True, **kw)
    assert (dep[0]["final_occupation"] == res[0]["final_occupation"]).all()
    assert dep[1]["conditions"]["temperature"] == 900.0 and dep[1]["potential_energy"]["n_samples"] == 6


def test_run_series_writes_the_reference_result_files_and_restarts(dev_tables, systems, tmp_path):
    """summary.json in the layout python/tests/conftest.py:180-240 (validate_summary_file) of
    the reference checks, completed_runs.json per RunData_json_io.hh, and the restart of
    python/tests/run_management/test_run_series.py: a series started again on the same output
    directory continues after the completed states -- dependent runs from the saved final
    state -- and ends with the files an uninterrupted series writes."""
    import json
    from test_run_series_cpu import validate_summary_file
    from casmcode_clexmonte_b200.results_io import RunDataOutputParams
    from casmcode_clexmonte_b200.run_series import run_series
    sysd = systems["zro"]
    eci = sysd["eci"]
    N = (6, 6, 4)
    n_cells = int(np.prod(N))
    occ = np.zeros(4 * n_cells, dtype=np.int32)
    init = {"temperature": 300.0, "param_chem_pot": [-4.0]}
    inc = {"temperature": 0.0, "param_chem_pot": [0.5]}
    kw = dict(n_equilibration_passes=4, n_samples=8, sample_period=1, seed=3, dependent_runs=True)

    def series(out, n_states):
        p = RunDataOutputParams(do_save_all_initial_states=True, do_save_all_final_states=True,
                                write_initial_states=True, write_final_states=True, output_dir=str(out))
        return run_series(dev_tables(sysd["tables"]), N, sysd, eci["index"], eci["value"], init, inc, n_states, occ,
                          output_params=p, **kw)

    full = series(tmp_path / "full", 5)
    assert len(full) == 5
    first = series(tmp_path / "parts", 2)
    rest = series(tmp_path / "parts", 5)          # reads completed_runs.json: runs states 2, 3, 4 only
    assert len(first) == 2 and len(rest) == 3
    assert [r["conditions"]["param_chem_pot"] for r in rest] == [[-3.0], [-2.5], [-2.0]]
    assert series(tmp_path / "parts", 5) == []    # complete: nothing left to run
    a = validate_summary_file(tmp_path / "full" / "summary.json", 5)
    b = validate_summary_file(tmp_path / "parts" / "summary.json", 5)
    for sec in ("analysis", "conditions", "statistics"):
        assert a[sec] == b[sec], sec
    assert a["completion_check_results"]["count"] == b["completion_check_results"]["count"] == [8 * 2 * n_cells] * 5
    assert a["conditions"]["param_chem_pot"]["a"] == [-4.0, -3.5, -3.0, -2.5, -2.0]
    ra = json.loads((tmp_path / "full" / "completed_runs.json").read_text())
    rb = json.loads((tmp_path / "parts" / "completed_runs.json").read_text())
    assert ra == rb and len(ra) == 5
    assert ra[3]["initial_state"]["configuration"]["dof"]["occ"] == ra[2]["final_state"]["configuration"]["dof"]["occ"]
    assert ra[0]["n_unitcells"] == n_cells and ra[0]["transformation_matrix_to_supercell"] == np.diag(N).tolist()
    assert (np.array(ra[4]["final_state"]["configuration"]["dof"]["occ"]) == full[4]["final_occupation"]).all()


def test_device_moments_equal_the_series_moments(dev_tables, systems):
    """cmx_sampler_moments ({n, sum q, sum q q^T} per replica, accumulated on the device: what
    ranks all-reduce for a replica grid) against the same sums of the downloaded series, and
    the analysis functions computed from them against Sampler.analysis."""
    from casmcode_clexmonte_b200 import replicas as R
    sysd = systems["fcc"]
    st = _capi.State(dev_tables(sysd["tables"]), (16, 16, 16), 3)
    st.set_eci(sysd["eci_sparse"]["index"], sysd["eci_sparse"]["value"])
    o2s = np.array(sysd["occ_to_species"], dtype=np.int32)
    st.set_occupants(sysd["sublat_to_asym"], o2s, sysd["n_species"])
    sm = _capi.Sampler(st, 40, sysd["axes"]["origin"], sysd["axes"]["Rt"])
    for r, (T, mu) in enumerate(((700.0, [0.1, 0.0]), (1000.0, [-0.2, 0.1]), (1500.0, [0.0, 0.3]))):
        st.set_conditions(T, semigrand_exchange_table(sysd["occ_to_species"], sysd["axes"]["Rt"], mu, 3), r)
        sm.set_param_chem_pot(mu, r)
    st.randomize(3)
    sm.run(40, 2, seed=5)
    got = sm.moments(first=4)
    for r in range(3):
        ser = sm.series(r, first=4)
        q = np.column_stack([ser["clex.formation_energy"], ser["potential_energy"], ser["mol_composition"],
                             ser["param_composition"]])
        np.testing.assert_allclose(got[r], R.moments_from_series(q), rtol=1e-12, atol=1e-12)
        a = R.analysis_from_moments(got[r], st.temperature(r), st.n_cells, sm.n_species, sm.n_param)
        b = sm.analysis(r, first=4)
        assert a["heat_capacity"] == pytest.approx(b["heat_capacity"], rel=1e-6, abs=1e-9)
        np.testing.assert_allclose(a["param_susc"], b["param_susc"], rtol=1e-6, atol=1e-8)
    sm.close()
    st.close()
