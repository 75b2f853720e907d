#!/usr/bin/env python3
"""Dump (and compile with nvcc for sm_100a) the generated transient kernel of a workload (host only).
   usage: tran_jit_dump.py <cfg3|cfg5|netlist-file> <out.cu>"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def main():
    wl, out = sys.argv[1], sys.argv[2]
    if wl == "cfg3":
        text, ov = workloads.RLC_TANK, workloads.rlc_tank_overrides(64)
    elif wl == "cfg5":
        text, ov = workloads.RECTIFIER, workloads.rectifier_overrides(64)
    else:
        text, ov = open(wl).read(), None
    table = packing.pack_circuit(parsing.parse_netlist(text))
    sweep = packing.make_sweep(table, 64, ov) if ov else None
    src = native.tran_kernel_source(table, sweep)
    open(out, "w").write(src)
    cubin = os.path.splitext(out)[0] + ".cubin"
    r = subprocess.run(["nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
                        "-o", cubin, out], capture_output=True, text=True)
    print(r.stderr.strip()[-2500:])


if __name__ == "__main__":
    main()
