"""Result rendering — host side, consumes the result shapes of analysis.py.

Mirrors lib/formatting/formatAcResult.ts:3-25, formatTranResult.ts:1-23 and
formatToVGraph.ts:11-66.  Out of the GPU hot path; kept so that
a caller switching from the reference finds the same functions and so the
golden snapshot of tests/basics/basics01.test.ts can be checked end to end.
"""
from __future__ import annotations

import math
from decimal import Decimal, ROUND_HALF_UP


def to_precision(x: float, p: int = 6) -> str:
    """ECMAScript Number.prototype.toPrecision(p)."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "0" if p == 1 else "0." + "0" * (p - 1)
    sign = "-" if x < 0 else ""
    d = Decimal(abs(x))  # exact binary value
    e = d.adjusted()
    q = Decimal(1).scaleb(e - p + 1)
    n = (d / q).to_integral_value(rounding=ROUND_HALF_UP)
    if n >= Decimal(10) ** p:  # rounding carried into a new digit
        e += 1
        q = Decimal(1).scaleb(e - p + 1)
        n = (d / q).to_integral_value(rounding=ROUND_HALF_UP)
    digits = str(int(n)).rjust(p, "0")
    if e < -6 or e >= p:
        mant = digits[0] + ("." + digits[1:] if p > 1 else "")
        return "%s%se%s%d" % (sign, mant, "+" if e >= 0 else "-", abs(e))
    if e >= 0:
        ip, fp = digits[: e + 1], digits[e + 1:]
        return sign + ip + ("." + fp if fp else "")
    return sign + "0." + "0" * (-e - 1) + digits


def format_ac_result(ac) -> str:
    """formatAcResult.ts:3-25."""
    if not ac:
        return "No AC analysis.\n"
    nodes = list(ac["nodeVoltages"].keys())
    lines = ["f(Hz), " + ", ".join("%s:|V|,∠V(deg)" % n for n in nodes)]
    freqs = ac["freqs"]
    for k in range(len(freqs)):
        parts = [to_precision(freqs[k])]
        for n in nodes:
            z = ac["nodeVoltages"][n][k]
            parts.append("%s,%s" % (to_precision(z.abs()), to_precision(z.phaseDeg())))
        lines.append(", ".join(parts))
    return "\n".join(lines)


def format_tran_result(tran) -> str:
    """formatTranResult.ts:1-23."""
    if not tran:
        return "No TRAN analysis.\n"
    nodes = list(tran["nodeVoltages"].keys())
    lines = [", ".join(["t(s)"] + ["%s:V" % n for n in nodes])]
    times = tran["times"]
    for k in range(len(times)):
        row = [to_precision(times[k])]
        for n in nodes:
            row.append(to_precision(float(tran["nodeVoltages"][n][k])))
        lines.append(", ".join(row))
    return "\n".join(lines)


def spicey_tran_to_vgraphs(tran, ckt, simulation_experiment_id: str):
    """formatToVGraph.ts:11-40."""
    if not tran or ckt.analyses.tran is None:
        return []
    dt, tstop = ckt.analyses.tran.dt, ckt.analyses.tran.tstop
    graphs = []
    for name, levels in tran["nodeVoltages"].items():
        graphs.append({
            "type": "simulation_transient_voltage_graph",
            "simulation_transient_voltage_graph_id": "stvg_%s_%s" % (simulation_experiment_id, name),
            "simulation_experiment_id": simulation_experiment_id,
            "timestamps_ms": [t * 1000 for t in tran["times"]],
            "voltage_levels": levels,
            "time_per_step": dt * 1000,
            "start_time_ms": 0,
            "end_time_ms": tstop * 1000,
            "name": "V(%s)" % name,
        })
    return graphs


def eec_engine_tran_to_vgraphs(tran_result, ckt, simulation_experiment_id: str):
    """formatToVGraph.ts:42-66."""
    if ckt.analyses.tran is None:
        return []
    dt, tstop = ckt.analyses.tran.dt, ckt.analyses.tran.tstop
    graphs = []
    for name, levels in tran_result["voltages"].items():
        graphs.append({
            "type": "simulation_transient_voltage_graph",
            "simulation_transient_voltage_graph_id": "stvg_%s_%s_eec" % (simulation_experiment_id, name),
            "simulation_experiment_id": simulation_experiment_id,
            "timestamps_ms": [t * 1000 for t in tran_result["time_s"]],
            "voltage_levels": levels,
            "time_per_step": dt * 1000,
            "start_time_ms": 0,
            "end_time_ms": tstop * 1000,
            "name": "V(%s) (ngspice)" % name,
        })
    return graphs
