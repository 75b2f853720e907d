"""ctypes binding of include/spicey_native.h — the Python twin of the reference-side
`lib/native` bun:ffi binding.  Loads the in-tree CUDA library and fails
loudly if it is missing or no device is usable: there is no CPU fallback in the product.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

from . import build as _build

ELEM_R, ELEM_C, ELEM_L, ELEM_V, ELEM_S, ELEM_D, ELEM_I = range(7)
VALUE_SLOTS = {ELEM_R: 1, ELEM_C: 1, ELEM_L: 1, ELEM_V: 3, ELEM_S: 4, ELEM_D: 2, ELEM_I: 3}
ST_OK, ST_SINGULAR, ST_CDIV, ST_R_NONPOS = 0, 1, 2, 3
FLAG_STRICT, FLAG_FORCE_GMEM, FLAG_FORCE_CTA, FLAG_DENSE, FLAG_SPARSE, FLAG_GENERIC_THREAD = 1, 2, 4, 8, 16, 32
FLAG_SERIES_MAJOR, FLAG_JIT, FLAG_NO_JIT, FLAG_WARP, FLAG_NO_WARP = 64, 128, 256, 512, 1024
FLAG_BAND, FLAG_NO_BAND = 2048, 4096
FLAG_TILE, FLAG_NO_TILE, FLAG_TILE_GENERIC = 8192, 16384, 32768
TIER_THREAD, TIER_CTA_SMEM, TIER_CTA_GMEM, TIER_SPARSE, TIER_SPARSE_JIT, TIER_TRAN_JIT, TIER_SPARSE_WARP = 1, 2, 3, 4, 5, 6, 7
TIER_BAND = 8
TIER_TILE = 9
SUCCESS, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3, 4

EXPORTS = [
    "spicey_native_abi_version", "spicey_device_count", "spicey_last_error", "spicey_create",
    "spicey_destroy", "spicey_get_stats", "spicey_host_alloc", "spicey_host_free", "spicey_ac_solve",
    "spicey_ac_solve_device", "spicey_tran_solve", "spicey_tran_solve_device", "spicey_measure_fp64_peak",
    "spicey_debug_sparse_source", "spicey_series_ld", "spicey_debug_tran_source", "spicey_debug_warp_stats",
    "spicey_tran_solve_waves", "spicey_tran_solve_waves_device", "spicey_debug_tran_source_waves",
    "spicey_debug_band_stats", "spicey_debug_band_source", "spicey_debug_tile_source", "spicey_debug_band_order_deviation",
    "spicey_debug_warp_lu_source", "spicey_tran_solve_probes",
]
WAVE_DC, WAVE_TABLE, WAVE_PULSE, WAVE_PWL = 0, 1, 2, 3

_ip = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)


class ElemTableStruct(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("n_elem", C.c_int32), ("n_values", C.c_int32), ("reserved", C.c_int32),
                ("type", _ip), ("n1", _ip), ("n2", _ip), ("nc1", _ip), ("nc2", _ip), ("value_idx", _ip),
                ("values", _dp)]


class SweepStruct(C.Structure):
    _fields_ = [("n_inst", C.c_int64), ("n_var", C.c_int32), ("reserved", C.c_int32), ("var_slot", _ip),
                ("var_values", C.c_void_p)]


class WavesStruct(C.Structure):
    _fields_ = [("n_vsrc", C.c_int32), ("reserved", C.c_int32), ("kind", _ip), ("value_idx", _ip), ("n_pairs", _ip)]


class StatsStruct(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_ms", C.c_double), ("kernel_launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("solves", C.c_int64), ("tier", C.c_int32),
                ("n_devices", C.c_int32), ("fallback_solves", C.c_int64), ("program_cfma", C.c_int64)]


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("spicey_native error %d: %s" % (code, msg))
        self.code = code


_LIB = None


def load_library(path: Optional[str] = None):
    """dlopen the CUDA library (building it in-tree with nvcc if it is stale or absent)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    if path is None:
        path = _build.LIB_PATH
        if _build.is_stale():
            try:
                _build.build()
            except Exception:
                if not os.path.exists(path):
                    raise
    lib = C.CDLL(path)
    lib.spicey_native_abi_version.restype = C.c_int32
    lib.spicey_device_count.restype = C.c_int32
    lib.spicey_last_error.restype = C.c_char_p
    lib.spicey_create.restype = C.c_int32
    lib.spicey_create.argtypes = [_ip, C.c_int32, C.POINTER(C.c_void_p)]
    lib.spicey_destroy.argtypes = [C.c_void_p]
    lib.spicey_destroy.restype = None
    lib.spicey_get_stats.argtypes = [C.c_void_p, C.POINTER(StatsStruct)]
    lib.spicey_host_alloc.restype = C.c_void_p
    lib.spicey_host_alloc.argtypes = [C.c_int64]
    lib.spicey_host_free.argtypes = [C.c_void_p]
    lib.spicey_host_free.restype = None
    tb, sw, vp = C.POINTER(ElemTableStruct), C.POINTER(SweepStruct), C.c_void_p
    lib.spicey_ac_solve.restype = C.c_int32
    lib.spicey_ac_solve.argtypes = [vp, tb, sw, vp, C.c_int64, vp, vp, vp, C.c_uint32]
    lib.spicey_ac_solve_device.restype = C.c_int32
    lib.spicey_ac_solve_device.argtypes = [vp, C.c_int32, tb, sw, vp, C.c_int64, vp, vp, vp, C.c_int64, C.c_uint32, vp]
    lib.spicey_series_ld.restype = C.c_int64
    lib.spicey_series_ld.argtypes = [C.c_int64]
    lib.spicey_tran_solve.restype = C.c_int32
    lib.spicey_tran_solve.argtypes = [vp, tb, sw, C.c_double, C.c_int64, vp, vp, vp, vp, vp, vp, vp, vp, C.c_uint32]
    lib.spicey_tran_solve_device.restype = C.c_int32
    lib.spicey_tran_solve_device.argtypes = [vp, C.c_int32, tb, sw, C.c_double, C.c_int64, vp, vp, vp, vp, vp, vp,
                                             vp, vp, C.c_uint32, vp]
    wv = C.POINTER(WavesStruct)
    lib.spicey_tran_solve_waves.restype = C.c_int32
    lib.spicey_tran_solve_waves.argtypes = [vp, tb, sw, C.c_double, C.c_int64, wv, vp, vp, vp, vp, vp, vp, vp, C.c_uint32]
    lib.spicey_tran_solve_probes.restype = C.c_int32
    lib.spicey_tran_solve_probes.argtypes = [vp, tb, sw, C.c_double, C.c_int64, wv, vp, vp, vp, vp, C.c_int32, vp, vp, vp, vp, vp,
                                             C.c_uint32]
    lib.spicey_tran_solve_waves_device.restype = C.c_int32
    lib.spicey_tran_solve_waves_device.argtypes = [vp, C.c_int32, tb, sw, C.c_double, C.c_int64, wv, vp, vp, vp, vp,
                                                   vp, vp, vp, C.c_uint32, vp]
    lib.spicey_debug_tran_source_waves.restype = C.c_int64
    lib.spicey_debug_tran_source_waves.argtypes = [tb, sw, wv, C.c_int32, C.c_char_p, C.c_int64]
    lib.spicey_measure_fp64_peak.restype = C.c_int32
    lib.spicey_measure_fp64_peak.argtypes = [vp, C.c_int32, _dp]
    lib.spicey_debug_sparse_source.restype = C.c_int64
    lib.spicey_debug_sparse_source.argtypes = [tb, sw, C.c_double, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_char_p,
                                               C.c_int64, _ip]
    lib.spicey_debug_warp_stats.restype = C.c_int32
    lib.spicey_debug_warp_stats.argtypes = [tb, C.c_double, _ip]
    lib.spicey_debug_tran_source.restype = C.c_int64
    lib.spicey_debug_tran_source.argtypes = [tb, sw, C.c_int32, C.c_char_p, C.c_int64]
    lib.spicey_debug_band_stats.restype = C.c_int32
    lib.spicey_debug_band_stats.argtypes = [tb, C.c_double, _ip]
    lib.spicey_debug_band_source.restype = C.c_int64
    lib.spicey_debug_band_source.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_uint32, C.c_int32, C.c_int32, C.c_int32,
                                             C.c_char_p, C.c_int64]
    lib.spicey_debug_warp_lu_source.restype = C.c_int64
    lib.spicey_debug_warp_lu_source.argtypes = [C.c_int32, C.c_int32, _ip, C.c_char_p, C.c_int64]
    lib.spicey_debug_tile_source.restype = C.c_int64
    lib.spicey_debug_tile_source.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _ip,
                                             C.c_char_p, C.c_int64]
    if path == _build.LIB_PATH:
        _LIB = lib
    return lib


def warp_program_stats(table: "ElemTable", pilot_f: float = 1000.0) -> dict:
    """Sizes of the warp-per-system sparse program (tier 7) of a circuit.  Host-only tooling."""
    lib = load_library()
    st = (C.c_int32 * 8)()
    ts = table.struct()
    _check(lib, lib.spicey_debug_warp_stats(C.byref(ts), pilot_f, st))
    keys = ("nvar", "pool_slots", "global_slots", "max_rows_per_step", "updates", "update_rows", "backsub_entries",
            "thread_tier_slots")
    return dict(zip(keys, list(st)))


def band_plan_stats(table: "ElemTable", pilot_f: float = 1000.0) -> dict:
    """Banded + bordered plan (tier 8) of a circuit: window, lanes per system, rows per lane, measured
    half-bandwidth, whether the nodes were renumbered, border rows.  Host-only tooling; raises NativeError
    (ERR_UNSUPPORTED) when the circuit does not qualify."""
    lib = load_library()
    st = (C.c_int32 * 8)()
    ts = table.struct()
    _check(lib, lib.spicey_debug_band_stats(C.byref(ts), pilot_f, st))
    keys = ("window", "lanes", "rows_per_lane", "bandwidth", "renumbered", "border_rows", "border_col_mask",
            "workspace_values")
    return dict(zip(keys, list(st)))


def band_order_deviation(table: "ElemTable", pilot_f: float, f: float) -> float:
    """Per-entry distance between the solutions in the banded plan's order and in the netlist order at frequency f
    (host only; band_plan.h: band_order_deviation)."""
    lib = load_library()
    lib.spicey_debug_band_order_deviation.restype = C.c_double
    lib.spicey_debug_band_order_deviation.argtypes = [C.POINTER(type(table.struct())), C.c_double, C.c_double]
    ts = table.struct()
    return float(lib.spicey_debug_band_order_deviation(C.byref(ts), pilot_f, f))


def band_kernel_source(lanes=8, rows_per_lane=2, border_rows=1, border_col_mask=0, with_ielem=True, warps=4,
                       min_blocks=2, rc_only=False, umode=0) -> str:
    """CUDA source NVRTC compiles for one band shape (tier 8); umode 2: the pivot rows leave through TMA tensor
    stores, 0: plain stores.  Host-only tooling."""
    lib = load_library()
    mask = (border_col_mask & 0xFFFF) | (int(bool(rc_only)) << 16) | ((umode & 3) << 17)
    args = (lanes, rows_per_lane, border_rows, mask, int(with_ielem), warps, min_blocks)
    need = lib.spicey_debug_band_source(*args, None, 0)
    buf = C.create_string_buffer(need)
    lib.spicey_debug_band_source(*args, buf, need)
    return buf.value.decode()


def tile_kernel_source(nvar: int, n_elem: int = 0, n_src: int = 1, tr: int = 0, tc: int = 0, with_ielem=True,
                       const_tables=False, rc_only=False):
    """(CUDA source, shape dict) of the dense register-tile kernel (tier 9) for an Nvar-unknown circuit; tr, tc > 0
    force the thread grid.  None when no shape fits an SM.  Host-only tooling."""
    lib = load_library()
    shp = (C.c_int32 * 8)()
    args = (nvar, n_elem, n_src, tr, tc, (1 if with_ielem else 0) | (2 if const_tables else 0) | (4 if rc_only else 0), shp)
    need = lib.spicey_debug_tile_source(*args, None, 0)
    if need < 0:
        return None
    buf = C.create_string_buffer(need)
    lib.spicey_debug_tile_source(*args, buf, need)
    keys = ("tr", "tc", "mr", "mc", "warps", "ctas_per_sm", "regs", "smem_bytes")
    return buf.value.decode(), dict(zip(keys, list(shp)))


def warp_lu_kernel_source(nvar: int, with_ielem=True, rc_only=False, const_tables=True):
    """(CUDA source, shape dict) of the one-warp-per-system dense LU (tier 9, Nvar <= 32); None above 32.  Host-only tooling."""
    lib = load_library()
    shp = (C.c_int32 * 3)()
    args = (nvar, (1 if with_ielem else 0) | (2 if const_tables else 0) | (4 if rc_only else 0), shp)
    need = lib.spicey_debug_warp_lu_source(*args, None, 0)
    if need < 0:
        return None
    buf = C.create_string_buffer(need)
    lib.spicey_debug_warp_lu_source(*args, buf, need)
    return buf.value.decode(), dict(zip(("warps", "ctas_per_sm", "smem_bytes"), list(shp)))


def tran_kernel_source(table: "ElemTable", sweep: Optional["Sweep"] = None, with_ielem=True,
                       waves: Optional["Waves"] = None) -> str:
    """CUDA source of the compiled transient kernel (tier 6) for a circuit.  Host-only tooling."""
    lib = load_library()
    ts = table.struct()
    ss = sweep.struct() if sweep else None
    sp_ = C.byref(ss) if ss else None
    ws = waves.struct() if waves else None
    wp_ = C.byref(ws) if ws else None
    need = lib.spicey_debug_tran_source_waves(C.byref(ts), sp_, wp_, int(with_ielem), None, 0)
    if need < 0:
        raise NativeError(-1, (lib.spicey_last_error() or b"").decode())
    buf = C.create_string_buffer(need)
    lib.spicey_debug_tran_source_waves(C.byref(ts), sp_, wp_, int(with_ielem), buf, need)
    return buf.value.decode()


def sparse_kernel_source(table: "ElemTable", pilot_f: float, block=192, min_blocks=1, smem_slots=75, with_ielem=True,
                         sync=4, sweep: Optional["Sweep"] = None, reg_values: Optional[int] = None):
    """CUDA source of the compiled straight-line sparse kernel (tier 5) for a circuit, plus the generator's
    statistics.  Host-only tooling: lets the generated code be inspected / compiled offline with nvcc.
    reg_values (0..127): cross-phase values beyond shared memory + that many registers go to the kernel's global
    column, as the library does (64); None: the round-1 generator (everything else is left to the registers); the last
    statistic is then the interpreter's slot count instead of the global column's."""
    lib = load_library()
    st = (C.c_int32 * 8)()
    ts = table.struct()
    ss = sweep.struct() if sweep else None
    sp_ = C.byref(ss) if ss else None
    mode = int(with_ielem) | (sync << 16)
    if reg_values is not None:
        mode |= 2 | ((int(reg_values) & 0x7f) << 24)
    need = lib.spicey_debug_sparse_source(C.byref(ts), sp_, pilot_f, block, min_blocks, smem_slots, mode, None, 0, st)
    if need < 0:
        raise NativeError(-1, (lib.spicey_last_error() or b"").decode())
    buf = C.create_string_buffer(need)
    lib.spicey_debug_sparse_source(C.byref(ts), sp_, pilot_f, block, min_blocks, smem_slots, mode, buf, need, st)
    keys = ("saved_values", "smem_slots", "classes", "micro_ops", "cfma", "reciprocals", "virtual_values",
            "interp_slots" if reg_values is None else "gmem_slots")
    return buf.value.decode(), dict(zip(keys, list(st)))


def _check(lib, rc):
    if rc != SUCCESS:
        raise NativeError(rc, (lib.spicey_last_error() or b"").decode())


class ElemTable:
    """Flat element table (struct of arrays) — see include/spicey_native.h."""

    def __init__(self, n_nodes, type_, n1, n2, nc1, nc2, value_idx, values, names=None, node_names=None):
        self.n_nodes = int(n_nodes)
        self.type = np.ascontiguousarray(type_, dtype=np.int32)
        self.n1 = np.ascontiguousarray(n1, dtype=np.int32)
        self.n2 = np.ascontiguousarray(n2, dtype=np.int32)
        self.nc1 = np.ascontiguousarray(nc1, dtype=np.int32)
        self.nc2 = np.ascontiguousarray(nc2, dtype=np.int32)
        self.value_idx = np.ascontiguousarray(value_idx, dtype=np.int32)
        self.values = np.ascontiguousarray(values, dtype=np.float64)
        self.names = list(names) if names is not None else None
        self.node_names = list(node_names) if node_names is not None else None
        self.n_elem = int(self.type.shape[0])
        self.n_vsrc = int((self.type == ELEM_V).sum())
        self.n_ac_elem = int((self.type <= ELEM_V).sum())
        self.n_state = int(np.isin(self.type, (ELEM_C, ELEM_L, ELEM_S, ELEM_D)).sum())
        self.nvar = self.n_nodes + self.n_vsrc
        self.waves = None        # native.Waves when packed with device_waves (packing.pack_circuit)
        self.wave_params = {}

    def struct(self) -> ElemTableStruct:
        s = ElemTableStruct()
        s.n_nodes, s.n_elem, s.n_values = self.n_nodes, self.n_elem, int(self.values.shape[0])
        for name in ("type", "n1", "n2", "nc1", "nc2", "value_idx"):
            setattr(s, name, getattr(self, name).ctypes.data_as(_ip))
        s.values = self.values.ctypes.data_as(_dp)
        return s


class Sweep:
    def __init__(self, n_inst: int, var_slot: Sequence[int] = (), var_values=None):
        self.n_inst = int(n_inst)
        self.var_slot = np.ascontiguousarray(var_slot, dtype=np.int32)
        n_var = int(self.var_slot.shape[0])
        if n_var:
            self.var_values = np.ascontiguousarray(var_values, dtype=np.float64).reshape(n_var, self.n_inst)
        else:
            self.var_values = np.zeros((0, self.n_inst), dtype=np.float64)

    def struct(self, device_ptr: Optional[int] = None) -> SweepStruct:
        s = SweepStruct()
        s.n_inst, s.n_var = self.n_inst, int(self.var_slot.shape[0])
        s.var_slot = self.var_slot.ctypes.data_as(_ip)
        s.var_values = device_ptr if device_ptr is not None else self.var_values.ctypes.data
        return s


class Waves:
    """Per-source waveform descriptors (spicey_waves): kind, first parameter slot, PWL pair count."""

    def __init__(self, kind, value_idx, n_pairs):
        self.kind = np.ascontiguousarray(kind, dtype=np.int32)
        self.value_idx = np.ascontiguousarray(value_idx, dtype=np.int32)
        self.n_pairs = np.ascontiguousarray(n_pairs, dtype=np.int32)

    def struct(self) -> WavesStruct:
        s = WavesStruct()
        s.n_vsrc = int(self.kind.shape[0])
        s.kind = self.kind.ctypes.data_as(_ip)
        s.value_idx = self.value_idx.ctypes.data_as(_ip)
        s.n_pairs = self.n_pairs.ctypes.data_as(_ip)
        return s


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Engine:
    """One spicey_handle.  `devices=None` -> device 0."""

    def __init__(self, devices: Optional[Sequence[int]] = None, lib_path: Optional[str] = None):
        self.lib = load_library(lib_path)
        if self.lib.spicey_native_abi_version() != 4:
            raise NativeError(ERR_INVALID, "ABI version mismatch")
        self._h = C.c_void_p()
        arr = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32)
        _check(self.lib, self.lib.spicey_create(None if arr is None else arr.ctypes.data_as(_ip),
                                                0 if arr is None else len(arr), C.byref(self._h)))
        self.n_devices = 1 if arr is None else len(arr)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.spicey_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stats(self) -> dict:
        s = StatsStruct()
        _check(self.lib, self.lib.spicey_get_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in StatsStruct._fields_}

    def fp64_peak_gflops(self, dev_index: int = 0) -> float:
        v = C.c_double()
        _check(self.lib, self.lib.spicey_measure_fp64_peak(self._h, dev_index, C.byref(v)))
        return v.value

    # -- host-buffer entry points ------------------------------------------------
    def ac_solve(self, table: ElemTable, freqs, sweep: Optional[Sweep] = None, want_currents=True, flags=0,
                 out=None):
        """Returns (x[P,nvar] c128, ielem[P,nAc] c128 | None, status[P] i32); P = n_inst*n_freq.
        With FLAG_SERIES_MAJOR in flags the arrays are x[nvar,P], ielem[nAc,P]."""
        freqs = np.ascontiguousarray(freqs, dtype=np.float64)
        F = int(freqs.shape[0])
        P = F * (sweep.n_inst if sweep else 1)
        if out is None:
            sm = bool(flags & FLAG_SERIES_MAJOR)
            x = np.empty((table.nvar, P) if sm else (P, table.nvar), dtype=np.complex128)
            ie = None
            if want_currents:
                ie = np.empty((table.n_ac_elem, P) if sm else (P, table.n_ac_elem), dtype=np.complex128)
            st = np.empty(P, dtype=np.int32)
        else:
            x, ie, st = out
        ts = table.struct()
        ss = sweep.struct() if sweep else None
        _check(self.lib, self.lib.spicey_ac_solve(self._h, C.byref(ts), C.byref(ss) if ss else None, _ptr(freqs), F,
                                                  _ptr(x), _ptr(ie), _ptr(st), flags))
        return x, ie, st

    def tran_solve(self, table: ElemTable, dt: float, steps: int, vsrc=None, vsrc_mask=None,
                   sweep: Optional[Sweep] = None, state0=None, want_currents=True, want_iters=False, flags=0,
                   out=None, waves: Optional[Waves] = None, node_sel=None):
        """Returns dict(v[S1,nn,n_inst], ielem[S1,n_elem,n_inst]|None, state[n_state,n_inst], iters, status).
        waves: per-source descriptors (PULSE / PWL evaluated on the device); vsrc_mask is then ignored.
        node_sel: node ids (1-based) whose voltages are wanted, e.g. the .PRINT TRAN probes — v is then [S1,len(node_sel),n_inst]
        and only those rows cross the bus (spicey_tran_solve_probes)."""
        n_inst = sweep.n_inst if sweep else 1
        S1 = steps + 1
        nV = table.n_vsrc
        if vsrc is not None:
            vsrc = np.ascontiguousarray(vsrc, dtype=np.float64).reshape(nV, S1)
        mask = np.zeros(max(1, nV), dtype=np.int32)
        if vsrc_mask is not None:
            mask[:nV] = np.asarray(vsrc_mask, dtype=np.int32)
        if state0 is not None:
            state0 = np.ascontiguousarray(state0, dtype=np.float64).reshape(table.n_state, n_inst)
        sel = None if node_sel is None else np.ascontiguousarray(node_sel, dtype=np.int32)
        if out is None:
            v = np.empty((S1, table.n_nodes if sel is None else len(sel), n_inst), dtype=np.float64)
            ie = np.empty((S1, table.n_elem, n_inst), dtype=np.float64) if want_currents else None
        else:
            v, ie = out
        state = np.zeros((table.n_state, n_inst), dtype=np.float64)
        iters = np.zeros((S1, n_inst), dtype=np.int32) if want_iters else None
        status = np.empty(n_inst, dtype=np.int32)
        ts = table.struct()
        ss = sweep.struct() if sweep else None
        if sel is not None:
            ws = waves.struct() if waves is not None else None
            _check(self.lib, self.lib.spicey_tran_solve_probes(
                self._h, C.byref(ts), C.byref(ss) if ss else None, float(dt), int(steps), C.byref(ws) if ws else None,
                _ptr(vsrc), _ptr(mask), _ptr(state0), _ptr(sel), int(len(sel)), _ptr(v) if len(sel) else None, _ptr(ie),
                _ptr(state), _ptr(iters), _ptr(status), flags))
        elif waves is not None:
            ws = waves.struct()
            _check(self.lib, self.lib.spicey_tran_solve_waves(
                self._h, C.byref(ts), C.byref(ss) if ss else None, float(dt), int(steps), C.byref(ws), _ptr(vsrc),
                _ptr(state0), _ptr(v), _ptr(ie), _ptr(state), _ptr(iters), _ptr(status), flags))
        else:
            _check(self.lib, self.lib.spicey_tran_solve(
                self._h, C.byref(ts), C.byref(ss) if ss else None, float(dt), int(steps), _ptr(vsrc), _ptr(mask),
                _ptr(state0), _ptr(v), _ptr(ie), _ptr(state), _ptr(iters), _ptr(status), flags))
        return {"v": v, "ielem": ie, "state": state, "iters": iters, "status": status}

    # -- device-resident entry points (raw device pointers as ints) ----------------
    def ac_solve_device(self, table: ElemTable, d_freqs: int, n_freq: int, d_x: int, d_ielem: Optional[int],
                        d_status: int, sweep: Optional[Sweep] = None, d_var_values: Optional[int] = None,
                        flags=0, stream: int = 0, dev_index: int = 0, series_ld: int = 0):
        """series_ld != 0: series-major results x[Nvar][series_ld], ielem[nAc][series_ld] (see series_ld())."""
        ts = table.struct()
        ss = sweep.struct(d_var_values) if sweep else None
        _check(self.lib, self.lib.spicey_ac_solve_device(
            self._h, dev_index, C.byref(ts), C.byref(ss) if ss else None, d_freqs, n_freq, d_x, d_ielem, d_status,
            series_ld, flags, stream))

    def series_ld(self, n_points: int) -> int:
        """Recommended leading dimension of a series-major device result (512-byte aligned rows)."""
        return int(self.lib.spicey_series_ld(n_points))

    def tran_solve_device(self, table: ElemTable, dt: float, steps: int, d_vsrc: Optional[int], vsrc_mask,
                          d_state0: Optional[int], d_v: int, d_ielem: Optional[int], d_state_out: Optional[int],
                          d_iters: Optional[int], d_status: int, sweep: Optional[Sweep] = None,
                          d_var_values: Optional[int] = None, flags=0, stream: int = 0, dev_index: int = 0,
                          waves: Optional[Waves] = None):
        ts = table.struct()
        ss = sweep.struct(d_var_values) if sweep else None
        if waves is not None:
            ws = waves.struct()
            _check(self.lib, self.lib.spicey_tran_solve_waves_device(
                self._h, dev_index, C.byref(ts), C.byref(ss) if ss else None, float(dt), int(steps), C.byref(ws), d_vsrc,
                d_state0, d_v, d_ielem, d_state_out, d_iters, d_status, flags, stream))
            return
        mask = np.zeros(max(1, table.n_vsrc), dtype=np.int32)
        if vsrc_mask is not None:
            mask[:table.n_vsrc] = np.asarray(vsrc_mask, dtype=np.int32)
        _check(self.lib, self.lib.spicey_tran_solve_device(
            self._h, dev_index, C.byref(ts), C.byref(ss) if ss else None, float(dt), int(steps), d_vsrc, _ptr(mask),
            d_state0, d_v, d_ielem, d_state_out, d_iters, d_status, flags, stream))


def pinned_empty(lib, shape, dtype):
    """numpy array over page-locked memory from spicey_host_alloc (caller frees p with lib.spicey_host_free)."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = lib.spicey_host_alloc(max(nbytes, 1))
    if not p:
        raise MemoryError("spicey_host_alloc(%d) failed" % nbytes)
    buf = (C.c_char * max(nbytes, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p
