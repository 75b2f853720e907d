#!/usr/bin/env python3
"""GPU: time the compiled sparse kernel (tier 5) on cfg2 for several launch shapes (SPICEY_JIT_CFG=block,minb,slots[,sync])
and check each against the strict dense pivoting kernel (reference-order arithmetic) on a subsample.
   usage: jit_sweep.py "192,1,75,4" "128,1,113,4" ...     (tests/ hold the parity checks against the oracle)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spicey_b200 as sp  # noqa: E402
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def main():
    wl = os.environ.get("SWEEP_WL", "cfg2")
    text = workloads.rc_ladder(64) if wl == "cfg2" else workloads.rc_mesh(int(wl[4:]) if wl.startswith("mesh") else 16)
    ck = parsing.parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck), dtype=np.float64)
    table = packing.pack_circuit(ck)
    if os.environ.get("SWEEP_P"):
        freqs = freqs[:int(os.environ["SWEEP_P"])]
    P = freqs.shape[0]
    dev = torch.device("cuda", 0)
    d_f = torch.from_numpy(freqs).to(dev)
    ld = (P + 31) // 32 * 32 if not os.environ.get("SWEEP_UNALIGNED") else P
    d_x = torch.empty((table.nvar, ld), dtype=torch.complex128, device=dev)
    d_i = torch.empty((table.n_ac_elem, ld), dtype=torch.complex128, device=dev)
    d_s = torch.empty(P, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sub = np.arange(0, P, 4999)
    eng0 = native.Engine([0])
    x0, i0, _ = eng0.ac_solve(table, freqs[sub], flags=native.FLAG_STRICT)
    ref = (x0.reshape(len(sub), -1), i0.reshape(len(sub), -1))
    eng0.close()
    for cfg in sys.argv[1:]:
        os.environ["SPICEY_JIT_CFG"] = cfg
        eng = native.Engine([0])
        flags = native.FLAG_SERIES_MAJOR | native.FLAG_JIT

        def step():
            eng.ac_solve_device(table, d_f.data_ptr(), P, d_x.data_ptr(), None if os.environ.get("JIT_SWEEP_NO_IELEM") else d_i.data_ptr(), d_s.data_ptr(), flags=flags,
                                stream=stream.cuda_stream, series_ld=ld)
        d_x.zero_(); d_i.zero_()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); step(); b.record(stream)
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        st = eng.stats()
        err = -1.0
        if ref is not None:
            x = d_x[:, torch.from_numpy(sub).to(dev)].cpu().numpy().T
            ie = d_i[:, torch.from_numpy(sub).to(dev)].cpu().numpy().T
            err = max(float(np.max(np.abs(x - ref[0]) / np.maximum(np.abs(ref[0]), 1e-300))),
                      float(np.max(np.abs(ie - ref[1]) / np.maximum(np.abs(ref[1]), 1e-300))))
        print("cfg=%-12s tier=%d fb=%d status_max=%d  ms min/med = %.3f / %.3f   %.1f M solves/s  relerr=%.2e" % (
            cfg, st["tier"], st["fallback_solves"], int(d_s.max().item()), min(ts), sorted(ts)[2], P / min(ts) / 1e3, err), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
