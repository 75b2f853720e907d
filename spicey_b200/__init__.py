"""spicey_b200 — B200-native batched MNA solve engine behind spicey's simulate() API.

Public surface mirrors lib/index.ts:1-12 of the reference (camelCase names kept), plus
the batch entry points and the native binding.  Importing the package never needs a GPU;
calling an analysis does (there is no CPU fallback).
"""
from .parsing import parse_netlist, parse_netlist as parseNetlist  # noqa: F401
from .analysis import (  # noqa: F401
    Complex, simulate, simulateAC, simulateTRAN, simulate_ac_batch, simulate_tran_batch, get_engine, set_engine,
)
from .formatting import (  # noqa: F401
    format_ac_result as formatAcResult, format_tran_result as formatTranResult,
    spicey_tran_to_vgraphs as spiceyTranToVGraphs, eec_engine_tran_to_vgraphs as eecEngineTranToVGraphs,
)
from . import native, packing, workloads  # noqa: F401

__all__ = [
    "parseNetlist", "parse_netlist", "simulate", "simulateAC", "simulateTRAN", "formatAcResult",
    "formatTranResult", "spiceyTranToVGraphs", "eecEngineTranToVGraphs", "Complex",
    "simulate_ac_batch", "simulate_tran_batch", "native", "packing", "workloads",
]
