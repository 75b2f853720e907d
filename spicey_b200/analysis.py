"""Drop-in analysis drivers: simulate / simulateAC / simulateTRAN with the reference's
result shapes, backed by the CUDA engine through the C ABI (pack -> FFI -> unpack).

Mirrors lib/analysis/simulate.ts:5-10, simulateAC.ts:62-130 and simulateTRAN.ts:130-252
at the boundary (names, argument meaning, `None` without the directive, error messages,
circuit state mutated by simulateTRAN), while the loop bodies run on the GPU.  Plus the
batch entry points the reference lacks: simulate_ac_batch /
simulate_tran_batch over Monte-Carlo / sweep instances.

Results are array-backed: `nodeVoltages[name]` is a lazy sequence over
a complex128 / float64 slab, so a 10^6-point sweep does not materialise 10^8 objects;
elements expose .re/.im/.abs()/.phaseDeg() like the reference's Complex.
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Optional

import numpy as np

from . import native
from .packing import has_wave_override, initial_state, make_sweep, pack_circuit, sample_sources, write_back_state
from .parsing import ParsedCircuit, build_frequency_array, compute_effective_time_step, parse_netlist

_ENGINE: Optional[native.Engine] = None


def get_engine() -> native.Engine:
    """Process-wide engine on device 0 (created on first use; raises without a GPU)."""
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = native.Engine()
    return _ENGINE


def set_engine(engine: Optional[native.Engine]) -> None:
    global _ENGINE
    _ENGINE = engine


class Complex:
    """Value type with the reference's accessors (lib/math/Complex.ts:3-62)."""

    __slots__ = ("re", "im")

    def __init__(self, re=0.0, im=0.0):
        self.re = float(re)
        self.im = float(im)

    def abs(self):
        return math.hypot(self.re, self.im)

    def phaseDeg(self):
        return (math.atan2(self.im, self.re) * 180) / math.pi

    def __complex__(self):
        return complex(self.re, self.im)

    def __repr__(self):
        return "Complex(%r, %r)" % (self.re, self.im)


class ComplexSeries:
    """Lazy Complex[] view over a complex128 array (strided views allowed)."""

    def __init__(self, arr: np.ndarray):
        self.array = arr

    def __len__(self):
        return int(self.array.shape[0])

    def __getitem__(self, k):
        if isinstance(k, slice):
            return ComplexSeries(self.array[k])
        z = self.array[k]
        return Complex(z.real, z.imag)

    def __iter__(self):
        for z in self.array:
            yield Complex(z.real, z.imag)


class LazyCurrentSeries(ComplexSeries):
    """Element-current series computed on access from the node-voltage slabs, exactly as the reference computes it per
    point (simulateAC.ts:94-126): i = Y.mul(v1.sub(v2)) with Y = 1/R, j 2 pi f C, or 1 / (j 2 pi f L) (0 when
    |2 pi f L| < EPS); a V element's current is its branch unknown.  Nothing is transferred from the device for it: the
    host call returns the solution vector only (16 Nvar bytes per point instead of 16 (Nvar + nAc))."""

    def __init__(self, kind, value, freqs, v1, v2):
        self.kind, self.value, self.freqs, self.v1, self.v2 = kind, value, freqs, v1, v2
        self._array = None

    def _admittance(self, f):
        two_pi = 2 * math.pi
        if self.kind == native.ELEM_R:
            return 1 / self.value, np.zeros_like(f)
        if self.kind == native.ELEM_C:
            return np.zeros_like(f), two_pi * f * self.value
        d = two_pi * f * self.value                      # denom = Complex(0, d); Complex.div: (0*0 + 0*d)/d^2, (0*0 - 1*d)/d^2
        with np.errstate(divide="ignore", invalid="ignore"):
            im = np.where(np.abs(d) < 1e-15, 0.0, (0.0 - d) / (d * d))
        return np.zeros_like(f), im

    def _compute(self, sel):
        f = np.asarray(self.freqs, dtype=np.float64)[sel]
        if self.kind == native.ELEM_V:
            return np.asarray(self.v1[sel])
        z = lambda a: 0.0 if a is None else a[sel]
        d = z(self.v1) - z(self.v2)
        yr, yi = self._admittance(f)
        dr, di = np.real(d), np.imag(d)
        return (yr * dr - yi * di) + 1j * (yr * di + yi * dr)   # Complex.mul, Complex.ts:33-38

    @property
    def array(self):
        if self._array is None:
            self._array = np.asarray(self._compute(slice(None)), dtype=np.complex128)
        return self._array

    def __len__(self):
        return len(self.freqs)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return ComplexSeries(self.array[k])
        if self._array is not None:
            zz = self._array[k]
        else:
            zz = complex(np.asarray(self._compute(slice(k, k + 1 if k != -1 else None)))[0])
        return Complex(zz.real, zz.imag)

    def __iter__(self):
        for zz in self.array:
            yield Complex(zz.real, zz.imag)


_INT_KEY = re.compile(r"^(0|[1-9]\d*)$")


def _js_key_order(names: List[str]) -> List[str]:
    """Own-property order of a JS object: canonical array indices ascending, then insertion order."""
    seen, ints, strs = set(), [], []
    for n in names:
        if n in seen:
            continue
        seen.add(n)
        (ints if _INT_KEY.match(n) and int(n) < 2 ** 32 - 1 else strs).append(n)
    return sorted(ints, key=int) + strs


def _series_by_name(names: List[str], cols: np.ndarray) -> Dict[str, np.ndarray]:
    """cols: [K, len(names)].  Duplicate names interleave per sample, as `(obj[name] ||= []).push` does."""
    out = {}
    for n in _js_key_order(names):
        idx = [i for i, m in enumerate(names) if m == n]
        out[n] = cols[:, idx[0]] if len(idx) == 1 else cols[:, idx].reshape(-1)
    return out


_AC_ERRORS = {native.ST_SINGULAR: "Singular matrix (complex)", native.ST_CDIV: "Complex divide by ~0"}
_TRAN_ERRORS = {native.ST_SINGULAR: "Singular matrix (real)"}


def _raise_first_failure(status: np.ndarray, table, ckt, messages) -> None:
    bad = np.nonzero(status)[0]
    if bad.size == 0:
        return
    code = int(status[bad[0]])
    if code == native.ST_R_NONPOS:
        name = next((r.name for r in ckt.R if r.R <= 0), "?")
        raise ValueError("R %s must be > 0" % name)  # simulateAC.ts:37
    raise ArithmeticError(messages.get(code, "solver failure %d" % code))


def ac_frequencies(ckt: ParsedCircuit) -> List[float]:
    a = ckt.analyses.ac
    return build_frequency_array(a.mode, a.N, a.f1, a.f2)


def simulateAC(ckt: ParsedCircuit, engine: Optional[native.Engine] = None, flags: int = 0, lazy_currents: bool = False):
    """simulateAC.ts:62-130.  Returns None without `.ac`; raises the reference's errors.
    lazy_currents: the device returns the solution vector only and `elementCurrents[name]` computes Y (v1 - v2) on
    access with the reference's own formula (:94-126) — a third of the bytes over PCIe for a ladder, identical keys."""
    if ckt.analyses.ac is None:
        return None
    eng = engine or get_engine()
    freqs = ac_frequencies(ckt)
    table = pack_circuit(ckt)
    # series-major results: every node / element series is one contiguous slab
    x, ie, st = eng.ac_solve(table, freqs, flags=flags | native.FLAG_SERIES_MAJOR, want_currents=not lazy_currents)
    _raise_first_failure(st, table, ckt, _AC_ERRORS)
    node_names = ckt.nodes.rev[1:]
    volt = _series_by_name(node_names, x[:table.n_nodes].T)
    if lazy_currents:
        names = table.names[:table.n_ac_elem]
        f = np.asarray(freqs, dtype=np.float64)
        series = []
        for e in range(table.n_ac_elem):
            kind = int(table.type[e])
            n1, n2 = int(table.n1[e]), int(table.n2[e])
            if kind == native.ELEM_V:
                k = e - int(np.argmax(table.type == native.ELEM_V))
                series.append(LazyCurrentSeries(kind, 0.0, f, x[table.n_nodes + k], None))
            else:
                series.append(LazyCurrentSeries(kind, float(table.values[int(table.value_idx[e])]), f,
                                                None if n1 == 0 else x[n1 - 1], None if n2 == 0 else x[n2 - 1]))
        lazy = {}
        for n in _js_key_order(names):
            idx = [i for i, m in enumerate(names) if m == n]
            # duplicate element names interleave per sample, as `(obj[name] ||= []).push` does: materialised
            lazy[n] = series[idx[0]] if len(idx) == 1 else ComplexSeries(np.stack([series[i].array for i in idx], axis=1).reshape(-1))
        return {"freqs": freqs, "nodeVoltages": {k: ComplexSeries(v) for k, v in volt.items()}, "elementCurrents": lazy}
    cur = _series_by_name(table.names[:table.n_ac_elem], ie.T)
    return {
        "freqs": freqs,
        "nodeVoltages": {k: ComplexSeries(v) for k, v in volt.items()},
        "elementCurrents": {k: ComplexSeries(v) for k, v in cur.items()},
    }


def simulateTRAN(ckt: ParsedCircuit, engine: Optional[native.Engine] = None, flags: int = 0):
    """simulateTRAN.ts:130-252.  Mutates the circuit's vPrev/iPrev/vdPrev/isOn like the reference."""
    if ckt.analyses.tran is None:
        return None
    eng = engine or get_engine()
    dt, steps = compute_effective_time_step(ckt.analyses.tran.dt, ckt.analyses.tran.tstop)
    table = pack_circuit(ckt)
    vsrc, mask = sample_sources(ckt, dt, steps)
    names = ckt.nodes.rev[1:]
    sel = None
    if len(ckt.probes.tran) > 0:  # :240-249 keeps the probed node voltages only (elementCurrents are not filtered):
        upper = [p.upper() for p in ckt.probes.tran]   # the others stay on the device (spicey_tran_solve_probes)
        sel = [i + 1 for i, k in enumerate(names) if k.upper() in upper]
        names = [names[i - 1] for i in sel]
    res = eng.tran_solve(table, dt, steps, vsrc=vsrc, vsrc_mask=mask, state0=initial_state(ckt, table), flags=flags, node_sel=sel)
    _raise_first_failure(res["status"], table, ckt, _TRAN_ERRORS)
    write_back_state(ckt, res["state"][:, 0])
    times = [k * dt for k in range(steps + 1)]  # t = step*dt, never accumulated (:147)
    volt = _series_by_name(names, res["v"][:, :, 0])
    cur = _series_by_name(table.names, res["ielem"][:, :, 0])
    return {"times": times, "nodeVoltages": volt, "elementCurrents": cur}


def simulate(netlist_text: str, engine: Optional[native.Engine] = None):
    """lib/analysis/simulate.ts:5-10."""
    circuit = parse_netlist(netlist_text)
    ac = simulateAC(circuit, engine)
    tran = simulateTRAN(circuit, engine)
    return {"circuit": circuit, "ac": ac, "tran": tran}


# ---- batch entry points (no reference equivalent; the caller-side loop of SURVEY §3.3) ----

def simulate_ac_batch(ckt: ParsedCircuit, freqs=None, n_inst: int = 1, overrides=None, want_currents=True,
                      engine: Optional[native.Engine] = None, flags: int = 0):
    """x[n_inst, F, nvar], ielem[n_inst, F, nAc], status[n_inst, F] for a sweep of instances."""
    eng = engine or get_engine()
    table = pack_circuit(ckt)
    if freqs is None:
        freqs = ac_frequencies(ckt)
    F = len(freqs)
    x, ie, st = eng.ac_solve(table, freqs, sweep=make_sweep(table, n_inst, overrides),
                             want_currents=want_currents, flags=flags)
    if flags & native.FLAG_SERIES_MAJOR:  # [rows, P] slabs -> logical [n_inst, F, rows] views (no copy)
        x = x.T
        ie = None if ie is None else ie.T
    return {"freqs": np.asarray(freqs), "x": x.reshape(n_inst, F, table.nvar),
            "ielem": None if ie is None else ie.reshape(n_inst, F, table.n_ac_elem),
            "status": st.reshape(n_inst, F), "node_names": ckt.nodes.rev[1:],
            "element_names": table.names[:table.n_ac_elem]}


def simulate_tran_batch(ckt: ParsedCircuit, n_inst: int = 1, overrides=None, want_currents=True,
                        want_iters=False, engine: Optional[native.Engine] = None, flags: int = 0,
                        device_waves: Optional[bool] = None):
    """v[S1, nn, n_inst], ielem[S1, n_elem, n_inst] for a Monte-Carlo / sweep batch (state not written back).
    device_waves: PULSE / PWL sources are evaluated by the kernels from their parameters instead of a row
    pre-sampled here (bit-identical values); the default does so exactly when an override names a waveform
    parameter ("v1.pulse.v2", "v1.pwl.t1", ...), which a shared pre-sampled row cannot express."""
    eng = engine or get_engine()
    dt, steps = compute_effective_time_step(ckt.analyses.tran.dt, ckt.analyses.tran.tstop)
    if device_waves is None:
        device_waves = has_wave_override(overrides)
    table = pack_circuit(ckt, device_waves=device_waves)
    vsrc, mask = (None, None) if device_waves else sample_sources(ckt, dt, steps)
    res = eng.tran_solve(table, dt, steps, vsrc=vsrc, vsrc_mask=mask, sweep=make_sweep(table, n_inst, overrides),
                         state0=initial_state(ckt, table, n_inst), want_currents=want_currents,
                         want_iters=want_iters, flags=flags, waves=table.waves)
    res.update({"dt": dt, "steps": steps, "times": np.arange(steps + 1) * dt, "node_names": ckt.nodes.rev[1:],
                "element_names": table.names})
    return res
