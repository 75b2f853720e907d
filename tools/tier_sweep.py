#!/usr/bin/env python3
"""GPU: time one AC workload under several flag sets (kernel tiers) and check each against the strict row kernel.
   usage: tier_sweep.py "SPARSE|NO_JIT" "SPARSE|JIT" "SPARSE|BAND" "DENSE" ...   (names of native.FLAG_*, SERIES_MAJOR is added)
   SWEEP_WL=ladder<n> | mesh<side> | dense<n> | rlc<nodes>x<elements> (default ladder400), SWEEP_P points (default 200000), SWEEP_NO_IELEM=1"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spicey_b200 as sp  # noqa: E402
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def main():
    wl = os.environ.get("SWEEP_WL", "ladder400")
    if wl.startswith("ladder"):
        text = workloads.rc_ladder(int(wl[6:]))
    elif wl.startswith("mesh"):
        text = workloads.rc_mesh(int(wl[4:]))
    elif wl.startswith("rlc"):   # rlc<nodes>x<extra elements>: a seeded random RLC network (resistor tree + random R / C / L)
        n_nodes, n_elem = (int(v) for v in wl[3:].split("x"))
        rng = np.random.default_rng(n_nodes * 1000 + n_elem)
        nodes = ["0"] + ["n%d" % i for i in range(1, n_nodes + 1)]
        lines = ["* random RLC", "v1 n1 0 dc 1 ac 1 0", "v2 n2 0 dc 1 ac 0.5 45"]
        for i in range(1, n_nodes + 1):
            lines.append("r%d n%d %s %g" % (i, i, nodes[rng.integers(0, i)], rng.uniform(10, 1e4)))
        for k in range(n_elem):
            a, b = rng.choice(len(nodes), 2, replace=False)
            kind = "rcl"[rng.integers(0, 3)]
            val = {"r": rng.uniform(10, 1e4), "c": rng.uniform(1e-9, 1e-6), "l": rng.uniform(1e-4, 1e-2)}[kind]
            lines.append("%s%d %s %s %g" % (kind, 100 + k, nodes[a], nodes[b], val))
        lines.append(".ac dec 40000 10 1meg")
        text = "\n".join(lines) + "\n"
    else:
        text = workloads.rc_dense(int(wl[5:]))
    ck = parsing.parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck), dtype=np.float64)
    table = packing.pack_circuit(ck)
    P = int(os.environ.get("SWEEP_P", "200000"))
    freqs = np.ascontiguousarray(freqs[:: max(1, freqs.shape[0] // P)][:P])
    P = freqs.shape[0]
    dev = torch.device("cuda", 0)
    d_f = torch.from_numpy(freqs).to(dev)
    no_i = bool(os.environ.get("SWEEP_NO_IELEM"))
    ld = (P + 31) // 32 * 32
    d_x = torch.empty((table.nvar, ld), dtype=torch.complex128, device=dev)
    d_i = None if no_i else torch.empty((table.n_ac_elem, ld), dtype=torch.complex128, device=dev)
    d_s = torch.empty(P, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sub = np.arange(0, P, max(1, P // 48))
    eng0 = native.Engine([0])
    x0, i0, _ = eng0.ac_solve(table, freqs[sub], flags=native.FLAG_STRICT)
    ref = x0.reshape(len(sub), -1)
    eng0.close()
    bytes_per = 8 + 16 * table.nvar + (0 if no_i else 16 * table.n_ac_elem)
    for arg in sys.argv[1:] or ["SPARSE"]:
        flags = native.FLAG_SERIES_MAJOR
        for nm in arg.split("|"):
            if nm:
                flags |= getattr(native, "FLAG_" + nm)
        eng = native.Engine([0])

        def step():
            eng.ac_solve_device(table, d_f.data_ptr(), P, d_x.data_ptr(), None if no_i else d_i.data_ptr(), d_s.data_ptr(),
                                flags=flags, stream=stream.cuda_stream, series_ld=ld)
        d_x.zero_()
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); step(); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        st = eng.stats()
        sel = torch.from_numpy(sub).to(dev)
        x = d_x[:, sel].T.cpu().numpy()
        err = float(np.max(np.abs(x - ref) / np.max(np.abs(ref), axis=1, keepdims=True)))
        rate = P / min(ts) / 1e3
        print("%-22s %s tier=%d fb=%d status_max=%d  ms min = %.3f   %.2f M solves/s  %.0f GB/s of results  relerr(row max)=%.2e" % (
            arg, wl, st["tier"], st["fallback_solves"], int(d_s.max().item()), min(ts), rate, rate * 1e6 * bytes_per / 1e9, err), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
