#!/usr/bin/env python3
"""GPU: time the dense register-tile tier (tier 9) on a dense circuit (SWEEP_WL=dense<n> complete RC graph, default
dense64; ladder<n>, mesh<side>) for several thread grids, and check each against the strict one-thread-per-row kernel on
a subsample.
   usage: tile_sweep.py "13,11" "16,12,2" "old" ...   each argument = SPICEY_TILE_SHAPE (TR,TC[,CTAs per SM]), "" = the
   library's choice, "old" = the shared-memory kernel (SPICEY_FLAG_NO_TILE).
   SWEEP_P points (default 100000), SWEEP_NO_IELEM=1 without element currents, SWEEP_PM=1 point-major results"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spicey_b200 as sp  # noqa: E402
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def main():
    wl = os.environ.get("SWEEP_WL", "dense64")
    if wl.startswith("ladder"):
        text = workloads.rc_ladder(int(wl[6:]))
    elif wl.startswith("mesh"):
        text = workloads.rc_mesh(int(wl[4:]))
    else:
        text = workloads.rc_dense(int(wl[5:]))
    ck = parsing.parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck), dtype=np.float64)
    table = packing.pack_circuit(ck)
    P = int(os.environ.get("SWEEP_P", "100000"))
    freqs = np.ascontiguousarray(freqs[:: max(1, freqs.shape[0] // P)][:P])
    P = freqs.shape[0]
    dev = torch.device("cuda", 0)
    d_f = torch.from_numpy(freqs).to(dev)
    pm = bool(os.environ.get("SWEEP_PM"))
    no_i = bool(os.environ.get("SWEEP_NO_IELEM"))
    ld = 0 if pm else (P + 31) // 32 * 32
    d_x = torch.empty((P, table.nvar) if pm else (table.nvar, ld), dtype=torch.complex128, device=dev)
    d_i = None if no_i else torch.empty((P, table.n_ac_elem) if pm else (table.n_ac_elem, ld), dtype=torch.complex128, device=dev)
    d_s = torch.empty(P, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sub = np.arange(0, P, max(1, P // 48))
    eng0 = native.Engine([0])
    x0, i0, _ = eng0.ac_solve(table, freqs[sub], flags=native.FLAG_STRICT)
    ref = (x0.reshape(len(sub), -1), i0.reshape(len(sub), -1))
    peak = eng0.fp64_peak_gflops()
    eng0.close()
    n = table.nvar
    dense_flops = 8.0 * (n * (n - 1) * (2 * n + 5) / 6 + n * (n - 1) / 2)   # elimination + back-substitution, complex FMAs x 8
    for arg in sys.argv[1:] or [""]:
        old = arg == "old"
        if arg and not old:
            os.environ["SPICEY_TILE_SHAPE"] = arg
        else:
            os.environ.pop("SPICEY_TILE_SHAPE", None)
        eng = native.Engine([0])
        flags = (0 if pm else native.FLAG_SERIES_MAJOR) | native.FLAG_DENSE | (native.FLAG_NO_TILE if old else native.FLAG_TILE)

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
        x = (d_x[sel] if pm else d_x[:, sel].T).cpu().numpy()
        scale = np.max(np.abs(ref[0]), axis=1, keepdims=True)
        err = float(np.max(np.abs(x - ref[0]) / scale))
        if not no_i:
            ie = (d_i[sel] if pm else d_i[:, sel].T).cpu().numpy()
            err = max(err, float(np.max(np.abs(ie - ref[1]) / np.max(np.abs(ref[1]), axis=1, keepdims=True))))
        rate = P / min(ts) / 1e3
        print("%-10s %s tier=%d status_max=%d  ms min/med = %.3f / %.3f   %.2f M solves/s  %.1f %% of the dense FP64 roofline (%.1f TF/s peak)  relerr(row max)=%.2e" % (
            arg or "auto", wl, st["tier"], int(d_s.max().item()), min(ts), sorted(ts)[1], rate,
            100 * rate * 1e6 * dense_flops / (peak * 1e9), peak / 1e3, err), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
