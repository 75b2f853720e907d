"""TEST INFRASTRUCTURE — ctypes wrapper over oracle/liboracle.so (oracle.c).

Builds the flat per-type arrays from a ParsedCircuit in the reference's list
order and runs the C restatement, optionally over a batch of instances with
per-instance component values and with several host threads (the CPU baseline
of bench.py).  Only tests/, __graft_entry__.smoke() and bench.py import this.

Per-instance overrides: dict name -> array[n_inst]; key is the element name for
R/C/L ("R1"), or "name.param" with param in dc/acmag/acphase (V),
ron/roff/von/voff (S), is/n (D).  Keys are matched case-insensitively.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ST_OK, ST_SINGULAR, ST_CDIV, ST_RNONPOS = 0, 1, 2, 3
STATUS_MESSAGES = {
    ST_SINGULAR: "Singular matrix",
    ST_CDIV: "Complex divide by ~0",
    ST_RNONPOS: "R must be > 0",
}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class _OCircuit(C.Structure):
    _ip, _dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    _fields_ = [(n, C.c_int32) for n in ("nn", "nR", "nC", "nL", "nV", "nS", "nD")] + [
        ("r_n1", _ip), ("r_n2", _ip), ("r_val", _dp),
        ("c_n1", _ip), ("c_n2", _ip), ("c_val", _dp),
        ("l_n1", _ip), ("l_n2", _ip), ("l_val", _dp),
        ("v_n1", _ip), ("v_n2", _ip), ("v_dc", _dp), ("v_acmag", _dp), ("v_acphase", _dp),
        ("s_n1", _ip), ("s_n2", _ip), ("s_cp", _ip), ("s_cn", _ip),
        ("s_ron", _dp), ("s_roff", _dp), ("s_von", _dp), ("s_voff", _dp),
        ("d_np", _ip), ("d_nm", _ip), ("d_is", _dp), ("d_n", _dp),
    ]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_ac.restype = C.c_int
        _LIB.oracle_tran.restype = C.c_int
        _LIB.oracle_solve_complex_flat.restype = C.c_int
        _LIB.oracle_solve_real_flat.restype = C.c_int
    return _LIB


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class FlatCircuit:
    """Keeps the numpy arrays alive for the lifetime of the ctypes struct."""

    def __init__(self, ckt, n_inst=1, overrides=None):
        ov = {k.lower(): np.asarray(v, dtype=np.float64) for k, v in (overrides or {}).items()}
        self.n_inst = n_inst
        self.keep = []

        def ints(vals):
            a = np.ascontiguousarray(np.array(list(vals), dtype=np.int32).reshape(-1))
            self.keep.append(a)
            return a

        def vals(items, nominal, suffix=None):
            cols = []
            for it, nom in zip(items, nominal):
                key = it.name.lower() + ("." + suffix if suffix else "")
                col = ov.get(key)
                cols.append(np.full(n_inst, float(nom)) if col is None else col.reshape(n_inst))
            a = np.ascontiguousarray(np.stack(cols, axis=1)) if cols else np.zeros((n_inst, 0))
            self.keep.append(a)
            return a

        s = _OCircuit()
        s.nn, s.nR, s.nC, s.nL = ckt.nodes.count() - 1, len(ckt.R), len(ckt.C), len(ckt.L)
        S = [sw for sw in ckt.S if sw.model is not None]
        D = [d for d in ckt.D if d.model is not None]
        s.nV, s.nS, s.nD = len(ckt.V), len(S), len(D)
        s.r_n1, s.r_n2 = _ip(ints(r.n1 for r in ckt.R)), _ip(ints(r.n2 for r in ckt.R))
        s.r_val = _dp(vals(ckt.R, [r.R for r in ckt.R]))
        s.c_n1, s.c_n2 = _ip(ints(c.n1 for c in ckt.C)), _ip(ints(c.n2 for c in ckt.C))
        s.c_val = _dp(vals(ckt.C, [c.C for c in ckt.C]))
        s.l_n1, s.l_n2 = _ip(ints(l.n1 for l in ckt.L)), _ip(ints(l.n2 for l in ckt.L))
        s.l_val = _dp(vals(ckt.L, [l.L for l in ckt.L]))
        s.v_n1, s.v_n2 = _ip(ints(v.n1 for v in ckt.V)), _ip(ints(v.n2 for v in ckt.V))
        s.v_dc = _dp(vals(ckt.V, [v.dc or 0 for v in ckt.V], "dc"))
        s.v_acmag = _dp(vals(ckt.V, [v.acMag or 0 for v in ckt.V], "acmag"))
        s.v_acphase = _dp(vals(ckt.V, [v.acPhaseDeg or 0 for v in ckt.V], "acphase"))
        s.s_n1, s.s_n2 = _ip(ints(x.n1 for x in S)), _ip(ints(x.n2 for x in S))
        s.s_cp, s.s_cn = _ip(ints(x.ncPos for x in S)), _ip(ints(x.ncNeg for x in S))
        s.s_ron = _dp(vals(S, [x.model.Ron for x in S], "ron"))
        s.s_roff = _dp(vals(S, [x.model.Roff for x in S], "roff"))
        s.s_von = _dp(vals(S, [x.model.Von for x in S], "von"))
        s.s_voff = _dp(vals(S, [x.model.Voff for x in S], "voff"))
        s.d_np, s.d_nm = _ip(ints(x.nPlus for x in D)), _ip(ints(x.nMinus for x in D))
        s.d_is = _dp(vals(D, [x.model.Is for x in D], "is"))
        s.d_n = _dp(vals(D, [x.model.N for x in D], "n"))
        self.struct = s
        self.ckt, self.S, self.D = ckt, S, D
        self.nvar = s.nn + s.nV
        self.ac_names = [e.name for e in ckt.R] + [e.name for e in ckt.C] + [e.name for e in ckt.L] + \
                        [e.name for e in ckt.V]
        self.tran_names = self.ac_names + [e.name for e in S] + [e.name for e in D]


def ac_solve(ckt, freqs, n_inst=1, overrides=None, want_currents=True, nthreads=1):
    """Returns x[n_inst*F, nvar] complex128, ielem[n_inst*F, nElem] complex128, status[n_inst*F]."""
    fc = FlatCircuit(ckt, n_inst, overrides)
    freqs = np.ascontiguousarray(freqs, dtype=np.float64)
    F = freqs.shape[0]
    P = F * n_inst
    x = np.zeros((P, fc.nvar), dtype=np.complex128)
    ie = np.zeros((P, len(fc.ac_names)), dtype=np.complex128) if want_currents else None
    st = np.zeros(P, dtype=np.int32)
    lib().oracle_ac(C.byref(fc.struct), _dp(freqs), C.c_int64(F), C.c_int64(n_inst),
                    x.ctypes.data_as(C.POINTER(C.c_double)),
                    ie.ctypes.data_as(C.POINTER(C.c_double)) if ie is not None else None,
                    st.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int(nthreads))
    return x, ie, st


def sample_sources(ckt, dt, steps):
    """[nV][steps+1] table of vs.waveform(step*dt) (simulateTRAN.ts:66-69,:147), and has-waveform flags."""
    nV = len(ckt.V)
    tab = np.zeros((nV, steps + 1), dtype=np.float64)
    flags = np.zeros(nV, dtype=np.int32)
    for i, vs in enumerate(ckt.V):
        if vs.waveform:
            flags[i] = 1
            for k in range(steps + 1):
                tab[i, k] = vs.waveform(k * dt)
    return tab, flags


def initial_state(ckt, n_inst):
    S = [sw for sw in ckt.S if sw.model is not None]
    D = [d for d in ckt.D if d.model is not None]
    row = [c.vPrev for c in ckt.C] + [l.iPrev for l in ckt.L] + [d.vdPrev for d in D] + \
          [1.0 if s.isOn else 0.0 for s in S]
    return np.ascontiguousarray(np.tile(np.array(row, dtype=np.float64), (n_inst, 1)))


def tran_solve(ckt, dt, steps, n_inst=1, overrides=None, state=None, want_currents=True, nthreads=1):
    """Returns v[n_inst, steps+1, nn], i[n_inst, steps+1, nElem], iters, status, state_out."""
    fc = FlatCircuit(ckt, n_inst, overrides)
    tab, flags = sample_sources(ckt, dt, steps)
    st8 = initial_state(ckt, n_inst) if state is None else np.ascontiguousarray(state, dtype=np.float64).copy()
    nn = fc.struct.nn
    v = np.zeros((n_inst, steps + 1, nn), dtype=np.float64)
    ie = np.zeros((n_inst, steps + 1, len(fc.tran_names)), dtype=np.float64) if want_currents else None
    iters = np.zeros((n_inst, steps + 1), dtype=np.int32)
    status = np.zeros(n_inst, dtype=np.int32)
    lib().oracle_tran(C.byref(fc.struct), C.c_double(dt), C.c_int64(steps), _dp(tab), _ip(flags),
                      C.c_int64(n_inst), _dp(st8), _dp(v), _dp(ie) if ie is not None else None,
                      _ip(iters), _ip(status), C.c_int(nthreads))
    return v, ie, iters, status, st8


def solve_complex(A, b):
    A = np.ascontiguousarray(A, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    n = A.shape[0]
    x = np.zeros(n, dtype=np.complex128)
    st = lib().oracle_solve_complex_flat(C.c_int(n), A.ctypes.data_as(C.POINTER(C.c_double)),
                                         b.ctypes.data_as(C.POINTER(C.c_double)),
                                         x.ctypes.data_as(C.POINTER(C.c_double)))
    return x, st


def solve_real(A, b):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = A.shape[0]
    x = np.zeros(n, dtype=np.float64)
    st = lib().oracle_solve_real_flat(C.c_int(n), _dp(A), _dp(b), _dp(x))
    return x, st
