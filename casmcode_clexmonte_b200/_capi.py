"""ctypes binding of the C ABI declared in ``include/cmx_b200.h``.

This is the same binding a reference-side plugin would make;
everything above it in this package is host bookkeeping.  There is no CPU
implementation behind these calls: if ``libcmx_b200.so`` is missing or no CUDA
device is present, they raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Tuple, Optional, Sequence

import numpy as np

from .clexulator_tables import ClexulatorTables

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libcmx_b200.so"

CMX_OK, CMX_ERR_INVALID, CMX_ERR_CUDA, CMX_ERR_UNSUPPORTED, CMX_ERR_STATE = range(5)
CMX_SWEEP_DE_SUM, CMX_SWEEP_FORCE_GENERIC = 1, 2
CMX_SWEEP_THREAD_GENERIC = 16
CMX_SWEEP_PAIR_SUM = 32
CMX_SWEEP_STREAM = 8
CMX_STATE_LINEAR_ROWS = 1

# every symbol include/cmx_b200.h declares
EXPORTED_SYMBOLS = [
    "cmx_last_error", "cmx_version", "cmx_device_count",
    "cmx_tables_create", "cmx_tables_create_from_file", "cmx_tables_destroy",
    "cmx_state_create", "cmx_state_create_opts", "cmx_state_create_general", "cmx_state_box", "cmx_supercell_box",
    "cmx_state_cell_index", "cmx_state_set_site_order", "cmx_state_destroy",
    "cmx_state_upload_occ", "cmx_state_download_occ",
    "cmx_state_upload_occ_i8", "cmx_state_download_occ_i8",
    "cmx_state_upload_occ_i8_async", "cmx_state_download_occ_i8_async", "cmx_state_synchronize", "cmx_sgc_sweep_async", "cmx_sgc_sweep_continue",
    "cmx_state_randomize", "cmx_state_set_k_offset", "cmx_state_stream", "cmx_state_device_ptr",
    "cmx_state_ipc_export", "cmx_state_ipc_attach", "cmx_state_p2p_active",
    "cmx_state_set_eci", "cmx_state_set_conditions", "cmx_state_set_occupants",
    "cmx_delta_corr", "cmx_point_corr", "cmx_cell_corr", "cmx_delta_e",
    "cmx_global_corr", "cmx_energy", "cmx_composition",
    "cmx_sgc_sweep", "cmx_sgc_sweep_kgroup", "cmx_sgc_sweep_slab", "cmx_state_set_sweep_flags", "cmx_counters_reset", "cmx_counters_read", "cmx_sweep_info", "cmx_sweep_launches", "cmx_sweep_term_counts", "cmx_sweep_stream_info", "cmx_sweep_debug_delta_e",
    "cmx_metropolis_sequential", "cmx_metropolis_sequential_ties", "cmx_rng_stream_test",
    "cmx_canonical_set_swaps", "cmx_canonical_default_swaps", "cmx_canonical_sweep", "cmx_canonical_info",
    "cmx_kmc_create", "cmx_kmc_destroy", "cmx_kmc_event_states", "cmx_kmc_all_rates",
    "cmx_kmc_set_impact_table", "cmx_kmc_run_begin", "cmx_kmc_run", "cmx_kmc_current_rates",
    "cmx_sampler_create", "cmx_sampler_destroy", "cmx_sampler_set_param_chem_pot", "cmx_sampler_info",
    "cmx_sampler_reset", "cmx_sampler_sample", "cmx_sampler_read", "cmx_sweep_run", "cmx_sampler_moments",
]


class CmxError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference throws
    std::runtime_error at the same places, e.g. CanonicalCalculator.cc:380-391)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[cmx {code}] {message}")
        self.code = code


class TableDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_sublat", "max_occ", "n_func", "corr_size", "n_point_corr", "nlist_len",
        "n_nlist_sublat", "n_factors", "n_terms", "n_elems", "n_groups")] + [
        ("nlist_sublat", C.c_void_p), ("n_occ", C.c_void_p), ("phi", C.c_void_p),
        ("nbr", C.c_void_p), ("factor_f", C.c_void_p), ("factor_n", C.c_void_p),
        ("term_coef", C.c_void_p), ("term_fbeg", C.c_void_p), ("elem_tbeg", C.c_void_p),
        ("group_ebeg", C.c_void_p), ("group_dphi", C.c_void_p), ("group_has_sum", C.c_void_p),
        ("group_div", C.c_void_p), ("global_gbeg", C.c_void_p), ("point_gbeg", C.c_void_p),
        ("delta_gbeg", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("n_attempt", C.c_int64), ("n_accept", C.c_int64), ("dE_sum", C.c_double),
                ("reserved", C.c_int64)]


class StepRecord(C.Structure):
    _fields_ = [("l0", C.c_int64), ("l1", C.c_int64), ("new0", C.c_int32), ("new1", C.c_int32),
                ("accepted", C.c_int32), ("pad", C.c_int32), ("dE", C.c_double)]


class PrimEvent(C.Structure):
    """cmx_prim_event <-> PrimEventData (include/casm/clexmonte/events/event_data.hh:47-72)"""
    _fields_ = [("n_sites", C.c_int32), ("site", (C.c_int32 * 4) * 4), ("occ_init", C.c_int32 * 4),
                ("occ_final", C.c_int32 * 4), ("event_type", C.c_int32), ("equivalent_index", C.c_int32)]


class EventType(C.Structure):
    _fields_ = [("n_equivalents", C.c_int32), ("local_tables", C.POINTER(C.c_void_p)),
                ("n_kra", C.c_int32), ("kra_index", C.c_void_p), ("kra_value", C.c_void_p),
                ("n_freq", C.c_int32), ("freq_index", C.c_void_p), ("freq_value", C.c_void_p)]


class EventState(C.Structure):
    """cmx_event_state <-> EventState (events/event_data.hh:19-30)"""
    _fields_ = [("is_allowed", C.c_int32), ("is_normal", C.c_int32), ("dE_final", C.c_double),
                ("Ekra", C.c_double), ("dE_activated", C.c_double), ("freq", C.c_double),
                ("rate", C.c_double)]


EVENT_STATE_DTYPE = np.dtype([("is_allowed", np.int32), ("is_normal", np.int32), ("dE_final", np.float64),
                              ("Ekra", np.float64), ("dE_activated", np.float64), ("freq", np.float64),
                              ("rate", np.float64)])

KMC_STEP_DTYPE = np.dtype([("unitcell", np.int64), ("prim_event", np.int32), ("pad", np.int32),
                           ("time_increment", np.float64), ("total_rate", np.float64)])

_lib = None


def lib():
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise CmxError(CMX_ERR_CUDA,
                       f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'`; "
                       "there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    L.cmx_last_error.restype = C.c_char_p
    L.cmx_tables_create.argtypes = [C.POINTER(TableDesc), C.c_int, C.POINTER(vp)]
    L.cmx_tables_destroy.argtypes = [vp]
    L.cmx_tables_destroy.restype = None
    L.cmx_state_create.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(vp)]
    L.cmx_state_destroy.argtypes = [vp]
    L.cmx_state_destroy.restype = None
    for f in ("cmx_state_upload_occ", "cmx_state_download_occ", "cmx_state_upload_occ_i8",
              "cmx_state_download_occ_i8"):
        getattr(L, f).argtypes = [vp, i32, vp]
    L.cmx_state_upload_occ_i8_async.argtypes = [vp, i32, vp]
    L.cmx_state_download_occ_i8_async.argtypes = [vp, i32, vp]
    L.cmx_state_synchronize.argtypes = [vp]
    L.cmx_sgc_sweep_async.argtypes = [vp, i64, u64, i64]
    L.cmx_sgc_sweep_continue.argtypes = [vp, i64, u64, i64]
    L.cmx_state_randomize.argtypes = [vp, u64]
    L.cmx_state_set_k_offset.argtypes = [vp, i32]
    L.cmx_state_device_ptr.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t)]
    L.cmx_state_stream.argtypes = [vp, C.POINTER(vp)]
    L.cmx_state_set_eci.argtypes = [vp, i32, vp, vp]
    L.cmx_state_ipc_export.argtypes = [vp, vp]
    L.cmx_state_ipc_attach.argtypes = [vp, vp, vp]
    L.cmx_state_p2p_active.argtypes = [vp, C.POINTER(i32)]
    L.cmx_state_set_conditions.argtypes = [vp, i32, dbl, vp]
    L.cmx_state_set_occupants.argtypes = [vp, vp, vp, i32]
    L.cmx_delta_corr.argtypes = [vp, i32, i64, vp, vp, vp]
    L.cmx_point_corr.argtypes = [vp, i32, i64, vp, vp]
    L.cmx_cell_corr.argtypes = [vp, i32, i64, vp, vp]
    L.cmx_delta_e.argtypes = [vp, i32, i64, i32, vp, vp, i32, vp]
    L.cmx_global_corr.argtypes = [vp, i32, vp]
    L.cmx_energy.argtypes = [vp, i32, C.POINTER(dbl)]
    L.cmx_composition.argtypes = [vp, i32, vp]
    L.cmx_sgc_sweep.argtypes = [vp, i64, u64, i64, vp]
    L.cmx_sgc_sweep_kgroup.argtypes = [vp, u64, i64, i32]
    L.cmx_sgc_sweep_slab.argtypes = [vp, i64, u64, i64]
    L.cmx_counters_reset.argtypes = [vp]
    L.cmx_state_set_sweep_flags.argtypes = [vp, C.c_uint32]
    L.cmx_counters_read.argtypes = [vp, vp]
    L.cmx_sweep_info.argtypes = [vp, C.c_char_p, C.c_size_t, C.POINTER(dbl), C.POINTER(dbl),
                                 C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.cmx_sweep_launches.argtypes = [vp, C.POINTER(i32)]
    L.cmx_sweep_term_counts.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
    L.cmx_sweep_stream_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.cmx_sweep_debug_delta_e.argtypes = [vp, i32, i64, vp, vp, vp]
    L.cmx_state_create_opts.argtypes = [vp, i32, i32, i32, i32, i32, C.c_uint32, C.POINTER(vp)]
    L.cmx_state_create_general.argtypes = [vp, vp, i32, C.c_uint32, C.POINTER(vp)]
    L.cmx_state_box.argtypes = [vp, vp]
    L.cmx_supercell_box.argtypes = [vp, vp]
    L.cmx_state_cell_index.argtypes = [vp, i64, vp, vp]
    L.cmx_state_set_site_order.argtypes = [vp, vp]
    L.cmx_metropolis_sequential.argtypes = [vp, i32, i32, i64, u64, vp, i64, C.POINTER(i64),
                                            C.POINTER(u64)]
    L.cmx_metropolis_sequential_ties.argtypes = [vp, C.POINTER(i64), vp, i32]
    L.cmx_rng_stream_test.argtypes = [u64, i64, vp, vp, vp, vp, vp, vp]
    L.cmx_canonical_set_swaps.argtypes = [vp, i32, vp]
    L.cmx_canonical_default_swaps.argtypes = [vp, i32, i32, i32, vp, C.POINTER(i32)]
    L.cmx_canonical_sweep.argtypes = [vp, i64, u64, i64, vp]
    L.cmx_canonical_info.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    L.cmx_kmc_create.argtypes = [vp, i32, C.POINTER(EventType), i32, C.POINTER(PrimEvent), C.POINTER(vp)]
    L.cmx_kmc_destroy.argtypes = [vp]
    L.cmx_kmc_destroy.restype = None
    L.cmx_kmc_event_states.argtypes = [vp, i64, vp, vp, vp, vp]
    L.cmx_kmc_all_rates.argtypes = [vp, vp, vp, C.POINTER(vp)]
    L.cmx_kmc_set_impact_table.argtypes = [vp, i32, vp, vp]
    L.cmx_kmc_run_begin.argtypes = [vp, vp]
    L.cmx_kmc_run.argtypes = [vp, i64, vp, i64, vp, vp, vp]
    L.cmx_kmc_current_rates.argtypes = [vp, vp, vp]
    L.cmx_sampler_create.argtypes = [vp, i32, i32, vp, vp, i32, C.POINTER(vp)]
    L.cmx_sampler_destroy.argtypes = [vp]
    L.cmx_sampler_set_param_chem_pot.argtypes = [vp, i32, vp]
    L.cmx_sampler_info.argtypes = [vp] + [C.POINTER(i32)] * 6
    L.cmx_sampler_reset.argtypes = [vp]
    L.cmx_sampler_sample.argtypes = [vp]
    L.cmx_sampler_read.argtypes = [vp, i32, i32, i32, vp]
    L.cmx_sweep_run.argtypes = [vp, vp, i32, i64, i64, u64, i64, vp]
    L.cmx_sampler_moments.argtypes = [vp, i32, vp, i32]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != CMX_OK:
        raise CmxError(rc, lib().cmx_last_error().decode())


def device_count() -> int:
    return int(lib().cmx_device_count())


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Tables:
    """Device copy of one basis set (``cmx_tables``)."""

    def __init__(self, t: ClexulatorTables, device: int = 0):
        self.host = t
        keep = {}

        def arr(name, dtype):
            a = np.ascontiguousarray(getattr(t, name), dtype=dtype)
            keep[name] = a
            return _p(a)

        d = TableDesc()
        d.n_sublat, d.max_occ, d.n_func = t.n_sublat, t.max_occ, t.n_func
        d.corr_size, d.n_point_corr = t.corr_size, t.n_point_corr
        d.nlist_len, d.n_nlist_sublat = t.nlist_len, t.n_nlist_sublat
        d.n_factors, d.n_terms = len(t.factor_f), len(t.term_coef)
        d.n_elems, d.n_groups = len(t.elem_tbeg) - 1, len(t.group_div)
        for name in ("nlist_sublat", "n_occ", "nbr", "factor_f", "factor_n", "term_fbeg",
                     "elem_tbeg", "group_ebeg", "group_dphi", "group_has_sum", "global_gbeg",
                     "point_gbeg", "delta_gbeg"):
            setattr(d, name, arr(name, np.int32))
        for name in ("phi", "term_coef", "group_div"):
            setattr(d, name, arr(name, np.float64))
        self._h = C.c_void_p()
        check(lib().cmx_tables_create(C.byref(d), device, C.byref(self._h)))
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            lib().cmx_tables_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def supercell_box(transformation_matrix) -> Tuple[int, ...]:
    """(N0, N1, N2, s10, s20, s21): the Hermite-normal-form box of a transformation matrix
    (host arithmetic only)."""
    T = np.ascontiguousarray(np.asarray(transformation_matrix).reshape(3, 3), dtype=np.int32)
    box = np.zeros(6, dtype=np.int32)
    check(lib().cmx_supercell_box(_p(T), _p(box)))
    return tuple(int(x) for x in box)


class State:
    """Device-resident supercell(s) (``cmx_state``)."""

    def __init__(self, tables: Tables, N: Sequence[int], n_replicas: int = 1, halo: int = 0,
                 linear_rows: bool = False, transformation_matrix=None):
        """N: the diag(N0, N1, N2) supercell, or (``transformation_matrix`` 3x3 integers,
        columns = supercell lattice vectors in prim coordinates, as the reference's
        ``transformation_matrix_to_super``) a general supercell -- N is then ignored and
        ``self.N`` / ``self.skew`` report the Hermite-normal-form box."""
        self.tables = tables
        self.n_replicas = int(n_replicas)
        self.halo = int(halo)
        t = tables.host
        self._temperature = {}
        self._h = C.c_void_p()
        opts = CMX_STATE_LINEAR_ROWS if linear_rows else 0
        if transformation_matrix is not None:
            T = np.ascontiguousarray(np.asarray(transformation_matrix).reshape(3, 3), dtype=np.int32)
            if halo:
                raise CmxError(CMX_ERR_INVALID, "slabs of a general supercell are not supported")
            check(lib().cmx_state_create_general(tables._h, _p(T), self.n_replicas, opts, C.byref(self._h)))
            box = np.zeros(6, dtype=np.int32)
            check(lib().cmx_state_box(self._h, _p(box)))
            self.N = tuple(int(x) for x in box[:3])
            self.skew = tuple(int(x) for x in box[3:])
        else:
            if np.isscalar(N):
                N = (N, N, N)
            self.N = tuple(int(x) for x in N)
            self.skew = (0, 0, 0)
            check(lib().cmx_state_create_opts(tables._h, *self.N, self.n_replicas, self.halo, opts,
                                              C.byref(self._h)))
        self.n_cells = self.N[0] * self.N[1] * self.N[2]
        self.n_sites = self.n_cells * t.n_sublat

    def cell_index(self, ijk) -> np.ndarray:
        """Unit cells by prim-lattice coordinates (any integers, [n][3]) -> cell index."""
        ijk = np.ascontiguousarray(np.asarray(ijk).reshape(-1, 3), dtype=np.int32)
        out = np.empty(ijk.shape[0], dtype=np.int64)
        check(lib().cmx_state_cell_index(self._h, ijk.shape[0], _p(ijk), _p(out)))
        return out

    def set_site_order(self, order) -> None:
        """order[l_caller] = l of this library (b * n_cells + cell); None: identity.  Applies
        to the synchronous upload_occ / download_occ."""
        if order is None:
            check(lib().cmx_state_set_site_order(self._h, None))
            return
        order = np.ascontiguousarray(order, dtype=np.int64)
        if order.size != self.n_sites:
            raise CmxError(CMX_ERR_INVALID, "site order must list every site")
        check(lib().cmx_state_set_site_order(self._h, _p(order)))

    def close(self):
        if getattr(self, "_h", None):
            lib().cmx_state_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data movement ------------------------------------------------------
    def upload_occ(self, occ: np.ndarray, replica: int = 0) -> None:
        if occ.dtype == np.int8:
            occ = np.ascontiguousarray(occ)
            fn = lib().cmx_state_upload_occ_i8
        else:
            occ = np.ascontiguousarray(occ, dtype=np.int32)
            fn = lib().cmx_state_upload_occ
        if occ.size != self.n_sites:
            raise CmxError(CMX_ERR_INVALID, f"occupation has {occ.size} sites, expected {self.n_sites}")
        check(fn(self._h, replica, _p(occ)))

    def download_occ(self, replica: int = 0, dtype=np.int32, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n_sites, dtype=dtype)
        fn = lib().cmx_state_download_occ_i8 if out.dtype == np.int8 else lib().cmx_state_download_occ
        check(fn(self._h, replica, _p(out)))
        return out

    # ---- asynchronous forms (pipelines of independent states; pinned int8 host buffers)
    def upload_occ_async(self, occ: np.ndarray, replica: int = 0) -> None:
        if occ.dtype != np.int8 or not occ.flags.c_contiguous or occ.size != self.n_sites:
            raise CmxError(CMX_ERR_INVALID, "upload_occ_async: contiguous int8 array of n_sites entries required")
        check(lib().cmx_state_upload_occ_i8_async(self._h, int(replica), _p(occ)))

    def download_occ_async(self, out: np.ndarray, replica: int = 0) -> None:
        if out.dtype != np.int8 or not out.flags.c_contiguous or out.size != self.n_sites:
            raise CmxError(CMX_ERR_INVALID, "download_occ_async: contiguous int8 array of n_sites entries required")
        check(lib().cmx_state_download_occ_i8_async(self._h, int(replica), _p(out)))

    def sgc_sweep_async(self, n_sweeps: int, seed: int, first_sweep: int = 0) -> None:
        check(lib().cmx_sgc_sweep_async(self._h, int(n_sweeps), int(seed), int(first_sweep)))

    def synchronize(self) -> None:
        check(lib().cmx_state_synchronize(self._h))

    def randomize(self, seed: int) -> None:
        check(lib().cmx_state_randomize(self._h, int(seed)))

    def set_k_offset(self, k_offset: int) -> None:
        check(lib().cmx_state_set_k_offset(self._h, int(k_offset)))

    def stream(self) -> int:
        """cudaStream_t the state's kernels run on (for external CUDA events)."""
        p = C.c_void_p()
        check(lib().cmx_state_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def device_ptr(self):
        p = C.c_void_p()
        n = C.c_size_t()
        check(lib().cmx_state_device_ptr(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    # -- model --------------------------------------------------------------
    def ipc_export(self) -> bytes:
        buf = C.create_string_buffer(128)
        check(lib().cmx_state_ipc_export(self._h, buf))
        return buf.raw

    def ipc_attach(self, handle_dn: Optional[bytes], handle_up: Optional[bytes]) -> None:
        """Handles of the lower / upper ring neighbour (None: this state itself)."""
        dn = C.create_string_buffer(handle_dn, 128) if handle_dn is not None else None
        up = C.create_string_buffer(handle_up, 128) if handle_up is not None else None
        check(lib().cmx_state_ipc_attach(self._h, dn, up))

    def p2p_active(self) -> bool:
        v = C.c_int32()
        check(lib().cmx_state_p2p_active(self._h, C.byref(v)))
        return bool(v.value)

    def set_eci(self, index, value) -> None:
        index = np.ascontiguousarray(index, dtype=np.uint32)
        value = np.ascontiguousarray(value, dtype=np.float64)
        if index.shape != value.shape:
            raise CmxError(CMX_ERR_INVALID, "ECI index/value length mismatch")
        self.eci_index, self.eci_value = index, value
        check(lib().cmx_state_set_eci(self._h, len(index), _p(index), _p(value)))

    def set_conditions(self, temperature: float, exch: Optional[np.ndarray] = None, replica: int = 0) -> None:
        t = self.tables.host
        if exch is not None:
            exch = np.ascontiguousarray(exch, dtype=np.float64)
            if exch.size != t.n_sublat * t.max_occ * t.max_occ:
                raise CmxError(CMX_ERR_INVALID, "exch must have n_sublat*max_occ*max_occ entries")
        check(lib().cmx_state_set_conditions(self._h, replica, float(temperature), _p(exch)))
        self._temperature[int(replica)] = float(temperature)

    def temperature(self, replica: int = 0) -> float:
        return self._temperature[int(replica)]

    def set_occupants(self, sublat_to_asym, occ_to_species, n_species: int) -> None:
        s2a = np.ascontiguousarray(sublat_to_asym, dtype=np.int32)
        o2s = np.ascontiguousarray(occ_to_species, dtype=np.int32)
        check(lib().cmx_state_set_occupants(self._h, _p(s2a), _p(o2s), int(n_species)))

    # -- potential / correlations ------------------------------------------
    def delta_corr(self, l, new_occ, replica: int = 0) -> np.ndarray:
        l = np.ascontiguousarray(np.atleast_1d(l), dtype=np.int64)
        new_occ = np.ascontiguousarray(np.atleast_1d(new_occ), dtype=np.int32)
        out = np.zeros((len(l), self.tables.host.corr_size))
        check(lib().cmx_delta_corr(self._h, replica, len(l), _p(l), _p(new_occ), _p(out)))
        return out

    def point_corr(self, l, replica: int = 0) -> np.ndarray:
        l = np.ascontiguousarray(np.atleast_1d(l), dtype=np.int64)
        out = np.zeros((len(l), self.tables.host.corr_size))
        check(lib().cmx_point_corr(self._h, replica, len(l), _p(l), _p(out)))
        return out

    def cell_corr(self, cells, replica: int = 0) -> np.ndarray:
        cells = np.ascontiguousarray(np.atleast_1d(cells), dtype=np.int64)
        out = np.zeros((len(cells), self.tables.host.corr_size))
        check(lib().cmx_cell_corr(self._h, replica, len(cells), _p(cells), _p(out)))
        return out

    def delta_e(self, l, new_occ, sites_per_event: int = 1, potential: bool = False,
                replica: int = 0) -> np.ndarray:
        l = np.ascontiguousarray(l, dtype=np.int64).reshape(-1)
        new_occ = np.ascontiguousarray(new_occ, dtype=np.int32).reshape(-1)
        if len(l) != len(new_occ) or len(l) % sites_per_event:
            raise CmxError(CMX_ERR_INVALID, "l/new_occ shape mismatch")
        n = len(l) // sites_per_event
        out = np.zeros(n)
        check(lib().cmx_delta_e(self._h, replica, n, sites_per_event, _p(l), _p(new_occ),
                                1 if potential else 0, _p(out)))
        return out

    def global_corr(self, replica: int = 0) -> np.ndarray:
        out = np.zeros(self.tables.host.corr_size)
        check(lib().cmx_global_corr(self._h, replica, _p(out)))
        return out

    def energy(self, replica: int = 0) -> float:
        e = C.c_double()
        check(lib().cmx_energy(self._h, replica, C.byref(e)))
        return e.value

    def composition(self, replica: int = 0) -> np.ndarray:
        t = self.tables.host
        out = np.zeros((t.n_sublat, t.max_occ), dtype=np.int64)
        check(lib().cmx_composition(self._h, replica, _p(out)))
        return out

    # -- drivers --------------------------------------------------------------
    def sgc_sweep(self, n_sweeps: int, seed: int, first_sweep: int = 0, counters: bool = True):
        cnt = (Counters * self.n_replicas)() if counters else None
        check(lib().cmx_sgc_sweep(self._h, int(n_sweeps), int(seed), int(first_sweep),
                                  C.byref(cnt) if counters else None))
        return cnt

    def sgc_sweep_enqueue(self, n_sweeps: int, seed: int, first_sweep: int = 0) -> None:
        """Asynchronous, counters keep accumulating (cmx_sgc_sweep_continue)."""
        check(lib().cmx_sgc_sweep_continue(self._h, int(n_sweeps), int(seed), int(first_sweep)))

    def sgc_sweep_slab(self, n_sweeps: int, seed: int, first_sweep: int = 0) -> None:
        """Peer-attached slab: whole sweeps in one cooperative launch (asynchronous)."""
        check(lib().cmx_sgc_sweep_slab(self._h, int(n_sweeps), int(seed), int(first_sweep)))

    def sgc_sweep_kgroup(self, seed: int, sweep: int, kgroup: int) -> None:
        """Asynchronous: enqueue one k-colour group of one sweep on the state's stream."""
        check(lib().cmx_sgc_sweep_kgroup(self._h, int(seed), int(sweep), int(kgroup)))

    # -- canonical pair exchanges ----------------------------------------------
    def canonical_set_swaps(self, swaps) -> list:
        """swaps: iterable of (b_a, b_b, (t0, t1, t2)).  Returns [(strides, n_colours)] per type."""
        arr = np.array([[a, b, t[0], t[1], t[2]] for a, b, t in swaps], dtype=np.int32)
        check(lib().cmx_canonical_set_swaps(self._h, len(arr), _p(arr)))
        out = []
        for i in range(len(arr)):
            S = (C.c_int32 * 3)()
            nc = C.c_int32()
            check(lib().cmx_canonical_info(self._h, i, S, C.byref(nc)))
            out.append((tuple(S), nc.value))
        return out

    def canonical_default_swaps(self, n_shell: int = 12, long_range: bool = True) -> list:
        """The library's default swap table [(b_a, b_b, (t0, t1, t2))] (needs set_occupants);
        the same list potential.canonical_swap_types builds on the host."""
        n = C.c_int32()
        check(lib().cmx_canonical_default_swaps(self._h, int(n_shell), int(bool(long_range)), 0, None, C.byref(n)))
        arr = np.zeros((max(1, n.value), 5), dtype=np.int32)
        check(lib().cmx_canonical_default_swaps(self._h, int(n_shell), int(bool(long_range)), n.value, _p(arr), C.byref(n)))
        return [(int(r[0]), int(r[1]), (int(r[2]), int(r[3]), int(r[4]))) for r in arr[:n.value]]

    def canonical_sweep(self, n_sweeps: int, seed: int, first_sweep: int = 0):
        cnt = (Counters * self.n_replicas)()
        check(lib().cmx_canonical_sweep(self._h, int(n_sweeps), int(seed), int(first_sweep), C.byref(cnt)))
        return cnt

    def set_sweep_flags(self, flags: int) -> None:
        check(lib().cmx_state_set_sweep_flags(self._h, int(flags)))

    def counters_reset(self) -> None:
        check(lib().cmx_counters_reset(self._h))

    def counters_read(self):
        cnt = (Counters * self.n_replicas)()
        check(lib().cmx_counters_read(self._h, C.byref(cnt)))
        return cnt

    def sweep_info(self) -> dict:
        name = C.create_string_buffer(32)
        b, f, nc, rk = C.c_double(), C.c_double(), C.c_int32(), C.c_int32()
        S = (C.c_int32 * 3)()
        check(lib().cmx_sweep_info(self._h, name, 32, C.byref(b), C.byref(f), C.byref(nc), S, C.byref(rk)))
        nl = C.c_int32()
        check(lib().cmx_sweep_launches(self._h, C.byref(nl)))
        st, sb, gr, gap = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        try:  # needs a device and conditions on every replica
            check(lib().cmx_sweep_stream_info(self._h, C.byref(st), C.byref(sb), C.byref(gr), C.byref(gap)))
        except CmxError:
            pass
        nt, nn = C.c_double(), C.c_double()
        check(lib().cmx_sweep_term_counts(self._h, C.byref(nt), C.byref(nn)))
        return dict(evaluator=name.value.decode(), bytes_per_step=b.value, flops_per_step=f.value,
                    terms_per_step=nt.value, neighbors_per_step=nn.value,
                    n_colours=nc.value, colour_strides=tuple(S), range_k=rk.value,
                    launches_per_sweep=nl.value, one_launch_per_call=(nl.value == 0), stream=bool(st.value), stream_blocks=sb.value,
                    stream_group_rowsteps=gr.value, stream_gap_units=gap.value)

    def sweep_debug_delta_e(self, l, new_occ, replica: int = 0) -> np.ndarray:
        """Delta potential energy of single-site proposals as the SWEEP's evaluator computes it
        (pair-LUT table entry / folded term lists), for comparison with delta_e()."""
        l = np.ascontiguousarray(l, dtype=np.int64)
        new_occ = np.ascontiguousarray(new_occ, dtype=np.int32)
        out = np.zeros(len(l))
        check(lib().cmx_sweep_debug_delta_e(self._h, replica, len(l), _p(l), _p(new_occ), _p(out)))
        return out

    def metropolis_sequential(self, mode: int, n_steps: int, seed: int, log_cap: int = 0,
                              replica: int = 0) -> dict:
        log = (StepRecord * max(1, log_cap))()
        n_acc = C.c_int64()
        h = C.c_uint64()
        check(lib().cmx_metropolis_sequential(self._h, replica, int(mode), int(n_steps), int(seed),
                                              C.byref(log), int(log_cap), C.byref(n_acc), C.byref(h)))
        steps = [dict(l0=s.l0, l1=s.l1, new0=s.new0, new1=s.new1, accepted=s.accepted, dE=s.dE)
                 for s in log[:min(log_cap, n_steps)]]
        nt = C.c_int64()
        ts = (C.c_int64 * 16)()
        check(lib().cmx_metropolis_sequential_ties(self._h, C.byref(nt), ts, 16))
        return dict(n_accept=n_acc.value, hash=h.value, log=steps, n_near_ties=nt.value,
                    tie_steps=[int(x) for x in ts if x >= 0])


KB = 8.6173303e-05  # eV/K, CASM::KB [EXT libcasm-global], pinned by _MonteCalculator.py:186-210


class Sampler:
    """Device-resident sample series of every replica of a State (cmx_sampler_*):
    the reference's state sampling functions `clex.formation_energy`,
    `potential_energy`, `mol_composition`, `param_composition`, `corr`
    (monte_calculator/sampling_functions.cc:37-288, all per unit cell) and, from
    the series, its analysis functions (monte_calculator/analysis_functions.cc:43-173)."""

    def __init__(self, state: "State", capacity: int, origin=None, Rt=None, with_corr: bool = False):
        self.state = state
        Rt = np.zeros((0, 0)) if Rt is None else np.ascontiguousarray(Rt, dtype=np.float64)
        self.n_param = int(Rt.shape[0])
        org = None if origin is None else np.ascontiguousarray(origin, dtype=np.float64)
        self._h = C.c_void_p()
        check(lib().cmx_sampler_create(state._h, int(capacity), self.n_param, _p(org),
                                       _p(Rt) if self.n_param else None, int(bool(with_corr)), C.byref(self._h)))
        v = [C.c_int32() for _ in range(6)]
        check(lib().cmx_sampler_info(self._h, *[C.byref(x) for x in v]))
        self.n_quantities, self.n_species, _, self.corr_size, _, self.capacity = [int(x.value) for x in v]

    def close(self):
        if getattr(self, "_h", None):
            lib().cmx_sampler_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_samples(self) -> int:
        v = C.c_int32()
        check(lib().cmx_sampler_info(self._h, None, None, None, None, C.byref(v), None))
        return int(v.value)

    def set_param_chem_pot(self, param_chem_pot, replica: int = 0) -> None:
        mu = None if param_chem_pot is None else np.ascontiguousarray(param_chem_pot, dtype=np.float64)
        if mu is not None and mu.size != self.n_param:
            raise CmxError(CMX_ERR_INVALID, "param_chem_pot has the wrong length")
        check(lib().cmx_sampler_set_param_chem_pot(self._h, int(replica), _p(mu)))

    def reset(self) -> None:
        check(lib().cmx_sampler_reset(self._h))

    def sample(self) -> None:
        """Append one sample of every replica (asynchronous)."""
        check(lib().cmx_sampler_sample(self._h))

    def run(self, n_samples: int, sweeps_per_sample: int, seed: int, first_sweep: int = 0,
            ensemble: str = "semigrand_canonical"):
        """n_samples x (sweeps_per_sample passes, one sample): one stream of launches."""
        ens = {"semigrand_canonical": 0, "canonical": 1}[ensemble]
        cnt = (Counters * self.state.n_replicas)()
        check(lib().cmx_sweep_run(self.state._h, self._h, ens, int(n_samples), int(sweeps_per_sample),
                                  int(seed), int(first_sweep), cnt))
        return list(cnt)

    @property
    def n_scalar_quantities(self) -> int:
        return 2 + self.n_species + self.n_param

    def moments(self, first: int = 0, device_ptr: Optional[int] = None) -> Optional[np.ndarray]:
        """{n, sum q, sum q q^T} of every replica over the samples [first, n_samples)
        (cmx_sampler_moments).  device_ptr: write to that device address instead (asynchronous
        on the state's stream; the caller hands the buffer to NCCL)."""
        Q = self.n_scalar_quantities
        if device_ptr is not None:
            check(lib().cmx_sampler_moments(self._h, int(first), C.c_void_p(int(device_ptr)), 1))
            return None
        out = np.zeros((self.state.n_replicas, 1 + Q + Q * Q))
        check(lib().cmx_sampler_moments(self._h, int(first), _p(out), 0))
        return out

    def series(self, replica: int = 0, first: int = 0, n: Optional[int] = None) -> dict:
        """Sampled series by the reference's sampler names."""
        n = self.n_samples - first if n is None else n
        raw = np.empty((n, self.n_quantities), dtype=np.float64)
        check(lib().cmx_sampler_read(self._h, int(replica), int(first), int(n), _p(raw)))
        s, p = self.n_species, self.n_param
        out = {"clex.formation_energy": raw[:, 0], "potential_energy": raw[:, 1],
               "mol_composition": raw[:, 2:2 + s], "param_composition": raw[:, 2 + s:2 + s + p]}
        if self.corr_size:
            out["corr"] = raw[:, 2 + s + p:]
        return out

    def analysis(self, replica: int = 0, first: int = 0) -> dict:
        """heat_capacity, mol_susc, param_susc, mol_thermochem_susc, param_thermochem_susc
        (analysis_functions.cc:43-173): covariances of the sampled series normalised by
        kB T^2 / n_unitcells resp. kB T / n_unitcells (run/io/convariance_functions.cc:29-143)."""
        ser = self.series(replica, first)
        T = self.state.temperature(replica)
        n_cells = self.state.n_cells

        def cov(a, b):
            a = a - a.mean(axis=0)
            b = b - b.mean(axis=0)
            return a.T @ b / a.shape[0]

        e = ser["potential_energy"][:, None]
        n, x = ser["mol_composition"], ser["param_composition"]
        c_heat = KB * T * T / n_cells
        c_susc = KB * T / n_cells
        return {"heat_capacity": float(cov(e, e)[0, 0] / c_heat),
                "mol_susc": cov(n, n) / c_susc, "param_susc": cov(x, x) / c_susc,
                "mol_thermochem_susc": cov(e, n)[0] / c_susc,
                "param_thermochem_susc": cov(e, x)[0] / c_susc}


def rng_stream_test(seed: int, kinds, int_max=None, real_max=None):
    kinds = np.ascontiguousarray(kinds, dtype=np.int32)
    n = len(kinds)
    int_max = np.ascontiguousarray(int_max if int_max is not None else np.zeros(n), dtype=np.int64)
    real_max = np.ascontiguousarray(real_max if real_max is not None else np.ones(n), dtype=np.float64)
    oi = np.zeros(n, dtype=np.int64)
    orl = np.zeros(n, dtype=np.float64)
    oraw = np.zeros(n, dtype=np.uint64)
    check(lib().cmx_rng_stream_test(int(seed), n, _p(int_max), _p(real_max), _p(kinds), _p(oi),
                                    _p(orl), _p(oraw)))
    return oi, orl, oraw


class Kmc:
    """Device-side event-state calculator of one State (cmx_kmc).

    event_types: list of dicts {local_tables: [Tables per equivalent], kra: (index, value),
    freq: (index, value)}; prim_events: list of dicts {sites: [(b,i,j,k)], occ_init, occ_final,
    event_type, equivalent_index} -- see casmcode_clexmonte_b200.kmc for builders."""

    def __init__(self, state: State, event_types, prim_events):
        self.state = state
        self._keep = []
        types = (EventType * len(event_types))()
        for y, et in enumerate(event_types):
            hs = (C.c_void_p * len(et["local_tables"]))(*[t._h for t in et["local_tables"]])
            ki = np.ascontiguousarray(et["kra"][0], dtype=np.uint32)
            kv = np.ascontiguousarray(et["kra"][1], dtype=np.float64)
            fi = np.ascontiguousarray(et["freq"][0], dtype=np.uint32)
            fv = np.ascontiguousarray(et["freq"][1], dtype=np.float64)
            self._keep += [hs, ki, kv, fi, fv, et["local_tables"]]
            types[y] = EventType(len(et["local_tables"]), hs, len(ki), _p(ki), _p(kv), len(fi), _p(fi), _p(fv))
        prim = (PrimEvent * len(prim_events))()
        for p, ev in enumerate(prim_events):
            e = prim[p]
            e.n_sites = len(ev["sites"])
            if e.n_sites > 4:
                raise CmxError(CMX_ERR_INVALID, "events with more than 4 sites are not supported")
            for q, site in enumerate(ev["sites"]):
                for c in range(4):
                    e.site[q][c] = int(site[c])
                e.occ_init[q] = int(ev["occ_init"][q])
                e.occ_final[q] = int(ev["occ_final"][q])
            e.event_type = int(ev["event_type"])
            e.equivalent_index = int(ev["equivalent_index"])
        self.n_prim = len(prim_events)
        self.event_types, self.prim_events = event_types, prim_events
        h = C.c_void_p()
        check(lib().cmx_kmc_create(state._h, len(event_types), types, len(prim_events), prim, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().cmx_kmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def event_states(self, unitcell, prim_event, replica=None) -> np.ndarray:
        """EventState records (structured array) of the listed events."""
        uc = np.ascontiguousarray(unitcell, dtype=np.int64)
        pe = np.ascontiguousarray(prim_event, dtype=np.int32)
        rp = np.zeros(len(uc), dtype=np.int32) if replica is None else np.ascontiguousarray(replica, dtype=np.int32)
        if not (len(uc) == len(pe) == len(rp)):
            raise CmxError(CMX_ERR_INVALID, "event_states: array lengths differ")
        out = np.zeros(len(uc), dtype=EVENT_STATE_DTYPE)
        assert out.itemsize == C.sizeof(EventState)
        check(lib().cmx_kmc_event_states(self._h, len(uc), _p(rp), _p(uc), _p(pe), _p(out)))
        return out

    # ---- rejection-free KMC on the device (cmx_kmc_run_*)
    def impact_table(self):
        """Relative impact table (events/ImpactTable.cc:150-185) from the exported tables."""
        from . import kmc as K
        st = self.state
        nbh = []
        for ev in self.prim_events:
            et = self.event_types[ev["event_type"]]
            coef = sorted(set(int(x) for x in et["kra"][0]) | set(int(x) for x in et["freq"][0]))
            nbh.append(K.required_update_neighborhood(st.tables.host, [int(x) for x in st.eci_index],
                                                      et["local_tables"][ev["equivalent_index"]].host, coef,
                                                      ev["sites"]))
        return K.make_relative_impact_table(self.prim_events, nbh)

    def set_impact_table(self, beg, entries) -> None:
        beg = np.ascontiguousarray(beg, dtype=np.int32)
        entries = np.ascontiguousarray(entries, dtype=np.int32).reshape(-1, 4)
        if len(beg) != self.n_prim + 1:
            raise CmxError(CMX_ERR_INVALID, "impact table: beg must have n_prim + 1 entries")
        check(lib().cmx_kmc_set_impact_table(self._h, len(entries), _p(beg), _p(entries)))
        self._impact = (beg, entries)

    def run_begin(self, seeds) -> None:
        """All rates from the current occupation, sum trees, std::mt19937_64(seeds[r]), time 0."""
        if getattr(self, "_impact", None) is None:
            self.set_impact_table(*self.impact_table())
        seeds = np.ascontiguousarray(np.broadcast_to(np.asarray(seeds, dtype=np.uint64), (self.state.n_replicas,)))
        check(lib().cmx_kmc_run_begin(self._h, _p(seeds)))

    def run(self, n_steps: int, log_cap: int = 0) -> dict:
        """n_steps rejection-free events of every trajectory."""
        R = self.state.n_replicas
        log = np.zeros((R, log_cap), dtype=KMC_STEP_DTYPE) if log_cap else None
        time, total = np.zeros(R), np.zeros(R)
        done = np.zeros(R, dtype=np.int64)
        check(lib().cmx_kmc_run(self._h, int(n_steps), _p(log), int(log_cap), _p(time), _p(total), _p(done)))
        return dict(time=time, total_rate=total, n_steps=done, log=log)

    def current_rates(self):
        """(rates[replica][unitcell][prim_event], total[replica]) as the selector holds them."""
        st = self.state
        r = np.zeros((st.n_replicas, st.n_cells, self.n_prim), dtype=np.float64)
        tot = np.zeros(st.n_replicas, dtype=np.float64)
        check(lib().cmx_kmc_current_rates(self._h, _p(r), _p(tot)))
        return r, tot

    def all_rates(self, rates: bool = True):
        """(rates[replica][unitcell][prim_event] or None, total[replica])"""
        st = self.state
        n_cells = int(np.prod(st.N))
        r = np.zeros((st.n_replicas, n_cells, self.n_prim), dtype=np.float64) if rates else None
        tot = np.zeros(st.n_replicas, dtype=np.float64)
        check(lib().cmx_kmc_all_rates(self._h, _p(r) if rates else None, _p(tot), None))
        return r, tot
