#!/usr/bin/env python3
"""GPU: RC ladders past cfg 2's size (and random sparse RC networks: a resistor tree + N/4 chords, every node with a
capacitor) through the compiled straight-line tier (tier 5, with its global column) against the interpreted program
(tier 4) of the same topology: M solves/s and the error against the strict dense kernel on a subsample.
   usage: ladder_probe.py 100 150 tree80 tree120 ...     LADDER_P points (default 1,000,000)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spicey_b200 as sp  # noqa: E402
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def main():
    P = int(os.environ.get("LADDER_P", "1000000"))
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    for arg in sys.argv[1:] or ["100", "150", "200"]:
        if arg.startswith("tree"):
            n = int(arg[4:])
            rng = np.random.default_rng(n)
            lines = ["* random sparse RC network", "v1 n1 0 dc 1 ac 1"]
            for i in range(2, n + 1):
                lines.append("r%d n%d n%d %g" % (i, i, rng.integers(1, i), rng.uniform(100, 1e4)))
            for i in range(1, n + 1):
                lines.append("c%d n%d 0 %g" % (i, i, rng.uniform(1e-9, 1e-7)))
            for k in range(n // 4):
                a, b = rng.choice(np.arange(1, n + 1), 2, replace=False)
                lines.append("r%d n%d n%d %g" % (1000 + k, a, b, rng.uniform(100, 1e4)))
            lines += [".ac dec %d 1 100k" % max(1, P // 5), ".end"]
            ck = parsing.parse_netlist("\n".join(lines) + "\n")
        else:
            n = int(arg)
            ck = parsing.parse_netlist(workloads.rc_ladder(n, ppd=max(1, P // 5)))
        freqs = np.ascontiguousarray(np.array(sp.analysis.ac_frequencies(ck), dtype=np.float64)[:P])
        table = packing.pack_circuit(ck)
        p = freqs.shape[0]
        ld = (p + 31) // 32 * 32
        d_f = torch.from_numpy(freqs).to(dev)
        d_x = torch.empty((table.nvar, ld), dtype=torch.complex128, device=dev)
        d_i = torch.empty((table.n_ac_elem, ld), dtype=torch.complex128, device=dev)
        d_s = torch.empty(p, dtype=torch.int32, device=dev)
        sub = np.arange(0, p, max(1, p // 64))
        e0 = native.Engine([0])
        x0, i0, _ = e0.ac_solve(table, freqs[sub], flags=native.FLAG_STRICT)
        e0.close()
        for name, flags in (("jit", native.FLAG_SPARSE | native.FLAG_JIT), ("interpreted", native.FLAG_SPARSE | native.FLAG_NO_JIT)):
            eng = native.Engine([0])

            def step():
                eng.ac_solve_device(table, d_f.data_ptr(), p, d_x.data_ptr(), d_i.data_ptr(), d_s.data_ptr(),
                                    flags=flags | native.FLAG_SERIES_MAJOR, stream=stream.cuda_stream, series_ld=ld)
            t0 = time.time()
            step()
            torch.cuda.synchronize()
            first = time.time() - t0
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
            ie = d_i[:, sel].T.cpu().numpy()
            err = max(float(np.max(np.abs(x - x0.reshape(len(sub), -1)) / np.maximum(np.abs(x0.reshape(len(sub), -1)), 1e-300))),
                      float(np.max(np.abs(ie - i0.reshape(len(sub), -1)) / np.maximum(np.abs(i0.reshape(len(sub), -1)), 1e-300))))
            bytes_per = 8 + 16 * (table.nvar + table.n_ac_elem)
            print("%-10s %-12s tier=%d fb=%d  first call %.1f s  ms min = %.3f  %.1f M solves/s  %.0f GB/s of results  relerr=%.2e" % (
                arg if arg.startswith("tree") else "ladder" + arg, name, st["tier"], st["fallback_solves"], first, min(ts), p / min(ts) / 1e3, p / min(ts) / 1e3 * bytes_per / 1e3, err), flush=True)
            eng.close()


if __name__ == "__main__":
    main()
