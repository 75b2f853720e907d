"""ParsedCircuit -> flat element table (the `lib/analysis` packer of the north star).

Elements are grouped R, C, L, V, S, D (then I, an extension) in netlist order inside a group — the order in
which the reference pushes element currents (simulateAC.ts:94-126,
simulateTRAN.ts:173-219).  Values are taken from the parsed circuit, i.e. produced by
the reference parser's own arithmetic (hazard H1).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .native import (ELEM_C, ELEM_D, ELEM_I, ELEM_L, ELEM_R, ELEM_S, ELEM_V, WAVE_DC, WAVE_PULSE, WAVE_PWL, ElemTable, Sweep,
                     Waves)

_PARAM_OFFSETS = {
    ELEM_V: {"dc": 0, "acmag": 1, "acphase": 2},
    ELEM_I: {"dc": 0, "acmag": 1, "acphase": 2},
    ELEM_S: {"ron": 0, "roff": 1, "von": 2, "voff": 3},
    ELEM_D: {"is": 0, "n": 1},
}


_PULSE_PARAMS = ("v1", "v2", "td", "tr", "tf", "ton", "period", "ncycles")  # PulseSpec, lib/types/simulation.ts:1-10


def pack_circuit(ckt, device_waves: bool = False) -> ElemTable:
    """device_waves: also append the PULSE / PWL parameters of the V elements as value slots (after the element
    values) and attach `table.waves` (native.Waves) so the kernels evaluate pulseValue / pwlValue themselves
     — the parameters can then be swept per instance like any R or C."""
    types, n1, n2, c1, c2, vidx, values, names = [], [], [], [], [], [], [], []

    def add(t, a, b, vals, name, ca=0, cb=0):
        types.append(t); n1.append(a); n2.append(b); c1.append(ca); c2.append(cb)
        vidx.append(len(values)); values.extend(float(v) for v in vals); names.append(name)

    for r in ckt.R:
        add(ELEM_R, r.n1, r.n2, [r.R], r.name)
    for c in ckt.C:
        add(ELEM_C, c.n1, c.n2, [c.C], c.name)
    for l in ckt.L:
        add(ELEM_L, l.n1, l.n2, [l.L], l.name)
    for v in ckt.V:
        add(ELEM_V, v.n1, v.n2, [v.dc or 0, v.acMag or 0, v.acPhaseDeg or 0], v.name)
    for s in ckt.S:
        if s.model is None:
            continue
        add(ELEM_S, s.n1, s.n2, [s.model.Ron, s.model.Roff, s.model.Von, s.model.Voff], s.name, s.ncPos, s.ncNeg)
    for d in ckt.D:
        if d.model is None:
            continue
        add(ELEM_D, d.nPlus, d.nMinus, [d.model.Is, d.model.N], d.name)
    for cs in getattr(ckt, "I", []):   # extension (parse_netlist(current_sources=True)): last group of the table
        add(ELEM_I, cs.n1, cs.n2, [cs.dc or 0, cs.acMag or 0, cs.acPhaseDeg or 0], cs.name)
    kinds, widx, npairs, wparams = [], [], [], {}
    if device_waves:
        for v in ckt.V:
            kinds.append(WAVE_DC); widx.append(0); npairs.append(0)
            if getattr(v, "pulse", None) is not None:
                kinds[-1], widx[-1] = WAVE_PULSE, len(values)
                for q, nm in enumerate(_PULSE_PARAMS):
                    wparams["%s.pulse.%s" % (v.name.lower(), nm)] = len(values) + q
                values.extend(float(getattr(v.pulse, nm)) for nm in _PULSE_PARAMS)
            elif getattr(v, "pwl", None) is not None:
                kinds[-1], widx[-1], npairs[-1] = WAVE_PWL, len(values), len(v.pwl)
                for q, (t, val) in enumerate(v.pwl):
                    wparams["%s.pwl.t%d" % (v.name.lower(), q)] = len(values)
                    wparams["%s.pwl.v%d" % (v.name.lower(), q)] = len(values) + 1
                    values.extend((float(t), float(val)))
    table = ElemTable(ckt.nodes.count() - 1, types, n1, n2, c1, c2, vidx, values, names=names,
                      node_names=ckt.nodes.rev[1:])
    table.waves = Waves(kinds, widx, npairs) if device_waves else None
    table.wave_params = wparams   # "v1.pulse.v2" / "v1.pwl.t3" -> value slot, for make_sweep
    return table


def has_wave_override(overrides) -> bool:
    return any(".pulse." in k.lower() or ".pwl." in k.lower() for k in (overrides or {}))


def make_sweep(table: ElemTable, n_inst: int, overrides: Optional[Dict[str, np.ndarray]]) -> Optional[Sweep]:
    """overrides: element name -> per-instance values (R/C/L), or "name.param" with param in
    dc/acmag/acphase (V), ron/roff/von/voff (S), is/n (D); for a table packed with device_waves also
    "name.pulse.v1|v2|td|tr|tf|ton|period|ncycles" and "name.pwl.t<k>|v<k>".  Case-insensitive."""
    if n_inst == 1 and not overrides:
        return None
    slots, rows = [], []
    lower = {n.lower(): i for i, n in reversed(list(enumerate(table.names)))}
    for key, vals in (overrides or {}).items():
        k = key.lower()
        wave_slots = getattr(table, "wave_params", None) or {}
        if ".pulse." in k or ".pwl." in k:
            if k not in wave_slots:
                raise KeyError("no waveform parameter %r (table packed with device_waves=True?)" % key)
            slots.append(wave_slots[k])
            arr = np.asarray(vals, dtype=np.float64).reshape(-1)
            if arr.shape[0] != n_inst:
                raise ValueError("override %r has %d values, expected %d" % (key, arr.shape[0], n_inst))
            rows.append(arr)
            continue
        name, _, param = k.partition(".")
        if name not in lower and k in lower:
            name, param = k, ""
        if name not in lower:
            raise KeyError("no element named %r" % key)
        e = lower[name]
        t = int(table.type[e])
        off = 0
        if param:
            off = _PARAM_OFFSETS.get(t, {}).get(param)
            if off is None:
                raise KeyError("element %r has no parameter %r" % (name, param))
        slots.append(int(table.value_idx[e]) + off)
        arr = np.asarray(vals, dtype=np.float64).reshape(-1)
        if arr.shape[0] != n_inst:
            raise ValueError("override %r has %d values, expected %d" % (key, arr.shape[0], n_inst))
        rows.append(arr)
    return Sweep(n_inst, slots, np.stack(rows) if rows else None)


def initial_state(ckt, table: ElemTable, n_inst: int = 1) -> np.ndarray:
    """[n_state][n_inst] from the circuit's vPrev / iPrev / isOn / vdPrev (simulateTRAN mutates these)."""
    row = [c.vPrev for c in ckt.C] + [l.iPrev for l in ckt.L] + \
          [1.0 if s.isOn else 0.0 for s in ckt.S if s.model is not None] + \
          [d.vdPrev for d in ckt.D if d.model is not None]
    return np.repeat(np.array(row, dtype=np.float64).reshape(-1, 1), n_inst, axis=1)


def write_back_state(ckt, state_col) -> None:
    """Final state of instance 0 back into the ParsedCircuit, as the reference leaves it."""
    it = iter(state_col)
    for c in ckt.C:
        c.vPrev = float(next(it))
    for l in ckt.L:
        l.iPrev = float(next(it))
    for s in ckt.S:
        if s.model is not None:
            s.isOn = next(it) != 0.0
    for d in ckt.D:
        if d.model is not None:
            d.vdPrev = float(next(it))


def sample_sources(ckt, dt: float, steps: int):
    """[nV][steps+1] of vs.waveform(k*dt) and the has-waveform mask (pre-sampled on the
    host so the values are the reference's own, simulateTRAN.ts:66-69,:147)."""
    nV = len(ckt.V)
    tab = np.zeros((nV, steps + 1), dtype=np.float64)
    mask = np.zeros(nV, dtype=np.int32)
    for i, vs in enumerate(ckt.V):
        if vs.waveform:
            mask[i] = 1
            tab[i] = [vs.waveform(k * dt) for k in range(steps + 1)]
    return tab, mask
