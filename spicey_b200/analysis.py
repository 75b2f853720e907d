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


def simulateAC(ckt: ParsedCircuit, engine: Optional[native.Engine] = None, flags: int = 0):
    """simulateAC.ts:62-130.  Returns None without `.ac`; raises the reference's errors."""
    if ckt.analyses.ac is None:
        return None
    eng = engine or get_engine()
    freqs = ac_frequencies(ckt)
    table = pack_circuit(ckt)
    # series-major results: every node / element series is one contiguous slab
    x, ie, st = eng.ac_solve(table, freqs, flags=flags | native.FLAG_SERIES_MAJOR)
    _raise_first_failure(st, table, ckt, _AC_ERRORS)
    node_names = ckt.nodes.rev[1:]
    volt = _series_by_name(node_names, x[:table.n_nodes].T)
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
    res = eng.tran_solve(table, dt, steps, vsrc=vsrc, vsrc_mask=mask, state0=initial_state(ckt, table), flags=flags)
    _raise_first_failure(res["status"], table, ckt, _TRAN_ERRORS)
    write_back_state(ckt, res["state"][:, 0])
    times = [k * dt for k in range(steps + 1)]  # t = step*dt, never accumulated (:147)
    volt = _series_by_name(ckt.nodes.rev[1:], res["v"][:, :, 0])
    cur = _series_by_name(table.names, res["ielem"][:, :, 0])
    if len(ckt.probes.tran) > 0:  # :240-249 (elementCurrents are not filtered)
        upper = [p.upper() for p in ckt.probes.tran]
        volt = {k: v for k, v in volt.items() if k.upper() in upper}
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
