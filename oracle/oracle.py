"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes front-end to ``oracle/liboracle.so`` (harness.cpp: our restatement of the
[EXT] call chain) driving ``oracle/_ref/*.so`` (the reference's own
CASM-generated Clexulator kernels, compiled unmodified by ``oracle/Makefile``).

May be imported ONLY from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product
package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_DIR = HERE / "_ref"
LIB_PATH = HERE / "liboracle.so"

# name -> (shared object stem, factory symbol); the factory names are the
# reference's own (e.g. FCC_binary_vacancy_Clexulator_default.cc:1065-1071).
CLEXULATORS = {
    "fcc_default": "FCC_binary_vacancy_Clexulator_default",
    "zro": "ZrO_Clexulator_formation_energy",
}
# the synthetic FCC binary pair + triplet basis (tests/golden/make_synthetic_clexulator.py, SURVEY 8c "Gap")
CLEXULATORS["fcc_synthetic"] = "FCC_binary_Clexulator_synthetic"
for _ev in ("A_Va_1NN", "B_Va_1NN"):
    for _k in range(6):
        CLEXULATORS[f"fcc_{_ev}_{_k}"] = f"FCC_binary_vacancy_Clexulator_{_ev}_{_k}"


def build(ref: bool = True) -> None:
    """Compile the harness (always) and oracle/_ref (when /root/reference exists)."""
    targets = ["harness"]
    if ref and Path("/root/reference").exists():
        targets.append("ref")
        targets.append("synthetic")
    subprocess.run(["make", "-s", "-j8", "-C", str(HERE)] + targets, check=True)


def available(name: str = "fcc_default") -> bool:
    return LIB_PATH.exists() and (REF_DIR / (CLEXULATORS[name] + ".so")).exists()


_lib = None


class OrcStep(C.Structure):
    _fields_ = [
        ("l0", C.c_long),
        ("l1", C.c_long),
        ("new0", C.c_int),
        ("new1", C.c_int),
        ("accepted", C.c_int),
        ("pad", C.c_int),
        ("dE", C.c_double),
    ]


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build(ref=False)
        L = C.CDLL(str(LIB_PATH))
        L.orc_kb.restype = C.c_double
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.c_char_p, C.c_char_p, C.c_long]
        L.orc_close.argtypes = [C.c_void_p]
        L.orc_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_nlist_cells.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_sublat_indices.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_weight_matrix.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_orbit_site_neighborhood.restype = C.c_long
        L.orc_orbit_site_neighborhood.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.orc_supercell.restype = C.c_void_p
        L.orc_supercell.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_long]
        L.orc_supercell_free.argtypes = [C.c_void_p]
        L.orc_delta_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        L.orc_restricted_delta_corr.argtypes = [
            C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_long, C.c_void_p]
        L.orc_point_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.orc_cell_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.orc_global_corr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_occ_delta_value.restype = C.c_double
        L.orc_occ_delta_value.argtypes = [
            C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_long, C.c_void_p]
        L.orc_rng_stream.argtypes = [
            C.c_uint64, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p]
        L.orc_metropolis_run.restype = C.c_long
        L.orc_metropolis_run.argtypes = [
            C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
            C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
            C.c_double, C.c_uint64, C.c_long, C.c_void_p, C.c_long, C.c_void_p,
            C.c_void_p, C.c_void_p]
        L.orc_potential_per_supercell.restype = C.c_double
        L.orc_potential_per_supercell.argtypes = [
            C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
            C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_long, C.c_void_p]
        _lib = L
    return _lib


KB = 8.6173303e-05


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefClexulator:
    """One of the reference's generated Clexulators, loaded from oracle/_ref."""

    def __init__(self, name: str, rmax_override: int = -1):
        stem = CLEXULATORS[name]
        so = REF_DIR / (stem + ".so")
        if not so.exists():
            raise FileNotFoundError(f"{so} missing: run `make -C oracle ref` where /root/reference exists")
        self._h = lib().orc_open(str(so).encode(), ("make_" + stem).encode(), rmax_override)
        if not self._h:
            raise RuntimeError("orc_open failed")
        info = np.zeros(8, dtype=np.int64)
        lib().orc_info(self._h, _p(info))
        (self.nlist_size, self.corr_size, self.n_point_corr, self.n_sublat,
         self.n_nlist_sublat, self.n_nlist_cells, self.rmax, self.n_neighborhood) = map(int, info)
        self.cells = np.zeros((self.n_nlist_cells, 3), dtype=np.int64)
        lib().orc_nlist_cells(self._h, _p(self.cells))
        self.sublats = np.zeros(self.n_nlist_sublat, dtype=np.int32)
        lib().orc_sublat_indices(self._h, _p(self.sublats))
        self.weight_matrix = np.zeros((3, 3), dtype=np.int64)
        lib().orc_weight_matrix(self._h, _p(self.weight_matrix))

    def orbit_site_neighborhood(self, corr: int) -> np.ndarray:
        buf = np.zeros((4096, 4), dtype=np.int64)
        n = lib().orc_orbit_site_neighborhood(self._h, corr, _p(buf), 4096)
        return buf[:n].copy()

    def supercell(self, N) -> "RefSupercell":
        return RefSupercell(self, N)


class RefSupercell:
    def __init__(self, clex: RefClexulator, N):
        if np.isscalar(N):
            N = (N, N, N)
        self.clex = clex
        self.N = tuple(int(x) for x in N)
        self.n_cells = self.N[0] * self.N[1] * self.N[2]
        self.n_sites = self.n_cells * clex.n_sublat
        self._h = lib().orc_supercell(clex._h, *self.N)

    def __del__(self):
        try:
            lib().orc_supercell_free(self._h)
        except Exception:
            pass

    @staticmethod
    def _occ(occ):
        occ = np.ascontiguousarray(occ, dtype=np.int32)
        return occ

    def delta_corr(self, occ, l: int, new_occ: int) -> np.ndarray:
        occ = self._occ(occ)
        out = np.zeros(self.clex.corr_size)
        lib().orc_delta_corr(self._h, _p(occ), int(l), int(new_occ), _p(out))
        return out

    def restricted_delta_corr(self, occ, l, new_occ, idx) -> np.ndarray:
        occ = self._occ(occ)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        out = np.zeros(self.clex.corr_size)
        lib().orc_restricted_delta_corr(self._h, _p(occ), int(l), int(new_occ), _p(idx), len(idx), _p(out))
        return out

    def point_corr(self, occ, l: int) -> np.ndarray:
        occ = self._occ(occ)
        out = np.zeros(self.clex.corr_size)
        lib().orc_point_corr(self._h, _p(occ), int(l), _p(out))
        return out

    def cell_corr(self, occ, cell: int) -> np.ndarray:
        occ = self._occ(occ)
        out = np.zeros(self.clex.corr_size)
        lib().orc_cell_corr(self._h, _p(occ), int(cell), _p(out))
        return out

    def global_corr(self, occ) -> np.ndarray:
        """Correlations per supercell (sum over unit cells)."""
        occ = self._occ(occ)
        out = np.zeros(self.clex.corr_size)
        lib().orc_global_corr(self._h, _p(occ), _p(out))
        return out

    def occ_delta_value(self, occ, l, new_occ, eci_idx, eci_val, return_dcorr=False):
        occ = self._occ(occ).copy()
        l = np.ascontiguousarray(np.atleast_1d(l), dtype=np.int64)
        new_occ = np.ascontiguousarray(np.atleast_1d(new_occ), dtype=np.int32)
        eci_idx = np.ascontiguousarray(eci_idx, dtype=np.uint32)
        eci_val = np.ascontiguousarray(eci_val, dtype=np.float64)
        dcorr = np.zeros(self.clex.corr_size)
        e = lib().orc_occ_delta_value(self._h, _p(occ), len(l), _p(l), _p(new_occ), _p(eci_idx),
                                      _p(eci_val), len(eci_idx), _p(dcorr))
        return (e, dcorr) if return_dcorr else e

    def metropolis_run(self, mode, occ, prim, eci_idx, eci_val, temperature, seed, n_steps,
                       param_chem_pot=None, log_cap=0):
        """Sequential loop.  ``prim`` = dict(sublat_to_asym, occ_to_species[n_sublat][max_occ],
        n_species, Rt[n_param][n_species]).  Returns dict(occ, n_accept, hash, log, seconds)."""
        occ = self._occ(occ).copy()
        s2a = np.ascontiguousarray(prim["sublat_to_asym"], dtype=np.int32)
        o2s = np.ascontiguousarray(prim["occ_to_species"], dtype=np.int32)
        n_sublat, max_occ = o2s.shape
        Rt = np.ascontiguousarray(prim.get("Rt", np.zeros((0, prim["n_species"]))), dtype=np.float64)
        mu = np.ascontiguousarray(param_chem_pot if param_chem_pot is not None else np.zeros(0),
                                  dtype=np.float64)
        n_param = len(mu)
        eci_idx = np.ascontiguousarray(eci_idx, dtype=np.uint32)
        eci_val = np.ascontiguousarray(eci_val, dtype=np.float64)
        log = (OrcStep * max(log_cap, 1))()
        n_acc = C.c_long(0)
        h = C.c_uint64(0)
        sec = C.c_double(0)
        rc = lib().orc_metropolis_run(
            self._h, int(mode), _p(occ), n_sublat, _p(s2a), _p(o2s), max_occ, int(prim["n_species"]),
            _p(mu), n_param, _p(Rt), _p(eci_idx), _p(eci_val), len(eci_idx), float(temperature),
            int(seed), int(n_steps), C.byref(log), int(log_cap), C.byref(n_acc), C.byref(h),
            C.byref(sec))
        if rc != n_steps:
            raise RuntimeError(f"orc_metropolis_run failed rc={rc}")
        steps = [dict(l0=s.l0, l1=s.l1, new0=s.new0, new1=s.new1, accepted=s.accepted, dE=s.dE)
                 for s in log[:min(log_cap, n_steps)]]
        return dict(occ=occ, n_accept=n_acc.value, hash=h.value, log=steps, seconds=sec.value)

    def potential_per_supercell(self, occ, prim, eci_idx, eci_val, param_chem_pot=None):
        occ = self._occ(occ)
        s2a = np.ascontiguousarray(prim["sublat_to_asym"], dtype=np.int32)
        o2s = np.ascontiguousarray(prim["occ_to_species"], dtype=np.int32)
        n_sublat, max_occ = o2s.shape
        Rt = np.ascontiguousarray(prim.get("Rt", np.zeros((0, prim["n_species"]))), dtype=np.float64)
        origin = np.ascontiguousarray(prim.get("origin", np.zeros(prim["n_species"])), dtype=np.float64)
        mu = np.ascontiguousarray(param_chem_pot if param_chem_pot is not None else np.zeros(0),
                                  dtype=np.float64)
        eci_idx = np.ascontiguousarray(eci_idx, dtype=np.uint32)
        eci_val = np.ascontiguousarray(eci_val, dtype=np.float64)
        comp = np.zeros(prim["n_species"])
        e = lib().orc_potential_per_supercell(
            self._h, _p(occ), n_sublat, _p(s2a), _p(o2s), max_occ, int(prim["n_species"]), _p(mu),
            len(mu), _p(Rt), _p(origin), _p(eci_idx), _p(eci_val), len(eci_idx), _p(comp))
        return e, comp


def rng_stream(seed: int, kinds, int_max=None, real_max=None):
    """Replay std::mt19937_64 + libstdc++ distributions (kind 0 raw, 1 int, 2 real)."""
    kinds = np.ascontiguousarray(kinds, dtype=np.int32)
    n = len(kinds)
    int_max = np.ascontiguousarray(int_max if int_max is not None else np.zeros(n), dtype=np.int64)
    real_max = np.ascontiguousarray(real_max if real_max is not None else np.ones(n), dtype=np.float64)
    oi = np.zeros(n, dtype=np.int64)
    orl = np.zeros(n, dtype=np.float64)
    oraw = np.zeros(n, dtype=np.uint64)
    lib().orc_rng_stream(int(seed), n, _p(int_max), _p(real_max), _p(kinds), _p(oi), _p(orl), _p(oraw))
    return oi, orl, oraw


def event_state(sc_formation: "RefSupercell", sc_local: "RefSupercell", occ, unitcell_index: int,
                linear_site_index, occ_init, occ_final, eci_idx, eci_val, kra, freq,
                temperature: float) -> dict:
    """EventStateCalculator::calculate_event_state + _default_event_state_calculation
    (src/casm/clexmonte/monte_calculator/BaseMonteEventData.cc:87-156) restated on
    top of the reference's generated kernels:
      is_allowed   event_is_allowed (events/event_methods.cc:340-351)
      dE_final     coefficients . Correlations::occ_delta(sites, occ_final)   :135-139
      local_corr   LocalCorrelations::local(unitcell_index, equivalent_index) :143-144
                   = calc_global_corr_contribution of the equivalent's local clexulator
      Ekra, freq   sparse dot products with the local correlations            :145-149
      dE_activated dE_final*0.5 + Ekra, is_normal, the two clamps             :152-156
      rate         freq * exp(-beta * dE_activated), beta = 1/(KB T)          :158
    `sc_local` is the supercell of the local clexulator of the event's equivalent."""
    occ = np.ascontiguousarray(occ, dtype=np.int32)
    st = dict(is_allowed=True, is_normal=False, dE_final=0.0, Ekra=0.0, dE_activated=0.0, freq=0.0, rate=0.0)
    for l, o in zip(linear_site_index, occ_init):
        if int(occ[l]) != int(o):
            st["is_allowed"] = False
            return st
    st["dE_final"] = sc_formation.occ_delta_value(occ, list(linear_site_index), list(occ_final), eci_idx, eci_val)
    local_corr = sc_local.cell_corr(occ, int(unitcell_index))
    ekra = 0.0
    for i, v in zip(*kra):
        ekra = ekra + float(v) * float(local_corr[int(i)])
    fr = 0.0
    for i, v in zip(*freq):
        fr = fr + float(v) * float(local_corr[int(i)])
    st["Ekra"], st["freq"] = ekra, fr
    dEa = st["dE_final"] * 0.5 + ekra
    st["is_normal"] = (dEa > 0.0) and (dEa > st["dE_final"])
    if dEa < st["dE_final"]:
        dEa = st["dE_final"]
    if dEa < 0.0:
        dEa = 0.0
    st["dE_activated"] = dEa
    beta = 1.0 / (KB * temperature)
    st["rate"] = fr * float(np.exp(-beta * dEa))
    st["local_corr"] = local_corr
    return st


# ---------------------------------------------------------------------------
# the reference's own KMC event selector (oracle/_ref/libkmc_lotto.so, kmc_lotto.cpp)
# ---------------------------------------------------------------------------
LOTTO_PATH = REF_DIR / "libkmc_lotto.so"
_RATE_CB = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_long)
_lotto = None


def lotto_available() -> bool:
    return LOTTO_PATH.exists()


def _lotto_lib():
    global _lotto
    if _lotto is None:
        L = C.CDLL(str(LOTTO_PATH))
        L.lotto_create.restype = C.c_void_p
        L.lotto_create.argtypes = [C.c_long, _RATE_CB, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong]
        L.lotto_destroy.argtypes = [C.c_void_p]
        L.lotto_select.restype = C.c_long
        L.lotto_select.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.lotto_get_rate.restype = C.c_double
        L.lotto_get_rate.argtypes = [C.c_void_p, C.c_long]
        _lotto = L
    return _lotto


class LottoSelector:
    """lotto::RejectionFreeEventSelector<long, ., std::mt19937_64> (the reference's
    submodules/kmc-lotto, compiled unmodified) over event ids 0..n_events-1.
    rate_fn(event_id) -> float is called for the initial rates and for every impacted
    event; impacted[e] lists the event ids to update after event e."""

    def __init__(self, n_events: int, rate_fn, impacted, seed: int):
        self._cb = _RATE_CB(lambda ctx, e: float(rate_fn(int(e))))
        beg = np.zeros(n_events + 1, dtype=np.int64)
        for e in range(n_events):
            beg[e + 1] = beg[e] + len(impacted[e])
        imp = (np.concatenate([np.asarray(impacted[e], dtype=np.int64) for e in range(n_events)])
               if beg[-1] else np.zeros(1, dtype=np.int64))
        self._h = _lotto_lib().lotto_create(n_events, self._cb, None, _p(beg), _p(np.ascontiguousarray(imp)), int(seed))

    def select(self):
        dt, tot = C.c_double(), C.c_double()
        e = _lotto_lib().lotto_select(self._h, C.byref(dt), C.byref(tot))
        return int(e), dt.value, tot.value

    def rate(self, event_id: int) -> float:
        return _lotto_lib().lotto_get_rate(self._h, int(event_id))

    def __del__(self):
        if getattr(self, "_h", None):
            _lotto_lib().lotto_destroy(self._h)
            self._h = None


class KmcReference:
    """The KMC loop of the reference on the CPU with nothing of this repository's product in
    it: lotto::RejectionFreeEventSelector (compiled unmodified) + event rates from the
    reference's generated Clexulator kernels (harness.cpp: orc_kmc_rate) + the apply step.
    prim: list of prim events (kmc.make_prim_event_list); types: event types with
    local_tables / kra / freq as (index, value); impacted[e]: event ids to update after e."""

    def __init__(self, N, occ, prim, types, eci_idx, eci_val, temperature, impacted, seed):
        L = lib()
        L.orc_kmc_ctx.restype = C.c_void_p
        L.orc_kmc_ctx.argtypes = [C.c_void_p, C.c_void_p, C.c_long] + [C.c_void_p] * 12 + [C.c_long, C.c_double]
        L.orc_kmc_ctx_free.argtypes = [C.c_void_p]
        K = _lotto_lib()
        K.lotto_run.restype = C.c_double
        K.lotto_run.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        self.N = tuple(int(x) for x in N)
        self.occ = np.ascontiguousarray(occ, dtype=np.int32).copy()
        self.n_prim = len(prim)
        self._form = RefClexulator("fcc_default").supercell(self.N)
        self._local = {}
        n_sites = np.array([len(p["sites"]) for p in prim], dtype=np.int32)
        sites = np.zeros((self.n_prim, 4, 4), dtype=np.int64)
        pocc = np.zeros((self.n_prim, 2, 4), dtype=np.int32)
        handles = (C.c_void_p * self.n_prim)()
        kb, ki, kv, fb, fi, fv = [0], [], [], [0], [], []
        for q, p in enumerate(prim):
            for s_, site in enumerate(p["sites"]):
                sites[q, s_] = site
            pocc[q, 0, :len(p["occ_init"])] = p["occ_init"]
            pocc[q, 1, :len(p["occ_final"])] = p["occ_final"]
            et = types[p["event_type"]]
            name = et["local_tables"][p["equivalent_index"]]
            if name not in self._local:
                self._local[name] = RefClexulator(name).supercell(self.N)
            handles[q] = self._local[name]._h
            ki += list(et["kra"][0]); kv += list(et["kra"][1]); kb.append(len(ki))
            fi += list(et["freq"][0]); fv += list(et["freq"][1]); fb.append(len(fi))
        arrs = [n_sites, sites, pocc, np.array(kb, dtype=np.int64), np.array(ki, dtype=np.uint32),
                np.array(kv, dtype=np.float64), np.array(fb, dtype=np.int64), np.array(fi, dtype=np.uint32),
                np.array(fv, dtype=np.float64), np.ascontiguousarray(eci_idx, dtype=np.uint32),
                np.ascontiguousarray(eci_val, dtype=np.float64)]
        self._keep = arrs + [handles]
        self._ctx = L.orc_kmc_ctx(self._form._h, _p(self.occ), self.n_prim, _p(arrs[0]), _p(arrs[1]), _p(arrs[2]),
                                  handles, _p(arrs[3]), _p(arrs[4]), _p(arrs[5]), _p(arrs[6]), _p(arrs[7]),
                                  _p(arrs[8]), _p(arrs[9]), _p(arrs[10]), len(arrs[9]), float(temperature))
        n_events = int(np.prod(self.N)) * self.n_prim
        beg = np.zeros(n_events + 1, dtype=np.int64)
        beg[1:] = np.cumsum([len(x) for x in impacted])
        imp = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.int64) for x in impacted]))
        rate = C.cast(L.orc_kmc_rate, C.c_void_p)
        K.lotto_create.argtypes = [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong]
        self._sel = K.lotto_create(n_events, rate, self._ctx, _p(beg), _p(imp), int(seed))
        K.lotto_create.argtypes = [C.c_long, _RATE_CB, C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulonglong]

    def run(self, n_steps: int):
        """(simulated time, selected event ids, wall seconds); self.occ is advanced."""
        import time as _t
        ev = np.zeros(n_steps, dtype=np.int64)
        t0 = _t.perf_counter()
        t = _lotto_lib().lotto_run(self._sel, int(n_steps), C.cast(lib().orc_kmc_apply, C.c_void_p), self._ctx, _p(ev))
        return t, ev, _t.perf_counter() - t0

    def __del__(self):
        try:
            if getattr(self, "_sel", None):
                _lotto_lib().lotto_destroy(self._sel)
                self._sel = None
            if getattr(self, "_ctx", None):
                lib().orc_kmc_ctx_free(self._ctx)
                self._ctx = None
        except Exception:
            pass
