#!/usr/bin/env python3
"""Dump (and optionally compile with nvcc for sm_100a) the generated straight-line sparse kernel of a workload.
   usage: jit_dump.py <cfg2|cfg4|ladderN> <out.cu> [block minb slots sync [noielem]]   (host only, no GPU needed)"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def main():
    wl, out = sys.argv[1], sys.argv[2]
    block, minb, slots = (int(v) for v in sys.argv[3:6]) if len(sys.argv) >= 6 else (192, 1, 75)
    sync = int(sys.argv[6]) if len(sys.argv) >= 7 and sys.argv[6].isdigit() else 4
    ie = "noielem" not in sys.argv
    if wl == "cfg2":
        text = workloads.rc_ladder()
    elif wl == "cfg4":
        text = workloads.rc_mesh()
    elif wl.startswith("ladder"):
        text = workloads.rc_ladder(int(wl[6:]))
    else:
        text = open(wl).read()
    table = packing.pack_circuit(parsing.parse_netlist(text))
    src, st = native.sparse_kernel_source(table, 1000.0, block, minb, slots, ie, sync)
    open(out, "w").write(src)
    print(st, "lines:", src.count("\n"))
    cubin = os.path.splitext(out)[0] + ".cubin"
    r = subprocess.run(["nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-lineinfo", "-Xptxas", "-v",
                        "-o", cubin, out], capture_output=True, text=True)
    print(r.stderr.strip()[-1500:])


if __name__ == "__main__":
    main()
