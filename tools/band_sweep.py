#!/usr/bin/env python3
"""GPU: time the banded + bordered tier (tier 8) on the cfg4 mesh (or SWEEP_WL=mesh<side> / ladder<n>) for several
band shapes and launch shapes, and check each against the strict dense pivoting kernel on a subsample.
   usage: band_sweep.py "8,2:4,2" "16,1:4,4" ...    each argument = SPICEY_BAND_SHAPE:SPICEY_BAND_CFG[:SPICEY_BAND_SYNC[:SPICEY_BAND_UMODE]] (L,RPL:warps,minb[:n[:0|1|2]])
   SWEEP_P points (default 400000), SWEEP_NO_IELEM=1 without element currents, SWEEP_PM=1 point-major results"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spicey_b200 as sp  # noqa: E402
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def main():
    wl = os.environ.get("SWEEP_WL", "mesh16")
    text = workloads.rc_ladder(int(wl[6:])) if wl.startswith("ladder") else workloads.rc_mesh(int(wl[4:]))
    ck = parsing.parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck), dtype=np.float64)
    table = packing.pack_circuit(ck)
    P = int(os.environ.get("SWEEP_P", "400000"))
    freqs = np.ascontiguousarray(freqs[:: max(1, freqs.shape[0] // P)][:P])
    P = freqs.shape[0]
    dev = torch.device("cuda", 0)
    d_f = torch.from_numpy(freqs).to(dev)
    pm = bool(os.environ.get("SWEEP_PM"))
    ld = 0 if pm else (P + 31) // 32 * 32
    d_x = torch.empty((P, table.nvar) if pm else (table.nvar, ld), dtype=torch.complex128, device=dev)
    d_i = torch.empty((P, table.n_ac_elem) if pm else (table.n_ac_elem, ld), dtype=torch.complex128, device=dev)
    d_s = torch.empty(P, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream()
    sub = np.arange(0, P, max(1, P // 64))
    eng0 = native.Engine([0])
    x0, i0, _ = eng0.ac_solve(table, freqs[sub], flags=native.FLAG_STRICT)
    ref = (x0.reshape(len(sub), -1), i0.reshape(len(sub), -1))
    eng0.close()
    no_i = bool(os.environ.get("SWEEP_NO_IELEM"))
    for arg in sys.argv[1:] or ["8,2:4,2"]:
        shape, _, cfg = arg.partition(":")
        cfg, _, sync = cfg.partition(":")
        sync, _, umode = sync.partition(":")
        if umode:
            os.environ["SPICEY_BAND_UMODE"] = umode
        else:
            os.environ.pop("SPICEY_BAND_UMODE", None)
        if sync:
            os.environ["SPICEY_BAND_SYNC"] = sync
        else:
            os.environ.pop("SPICEY_BAND_SYNC", None)
        os.environ["SPICEY_BAND_SHAPE"] = shape
        if cfg:
            os.environ["SPICEY_BAND_CFG"] = cfg
        else:
            os.environ.pop("SPICEY_BAND_CFG", None)
        eng = native.Engine([0])
        flags = (0 if pm else native.FLAG_SERIES_MAJOR) | native.FLAG_SPARSE | native.FLAG_BAND

        def step():
            eng.ac_solve_device(table, d_f.data_ptr(), P, d_x.data_ptr(), None if no_i else d_i.data_ptr(), d_s.data_ptr(),
                                flags=flags, stream=stream.cuda_stream, series_ld=ld)
        d_x.zero_(); d_i.zero_()
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
        err = float(np.max(np.abs(x - ref[0]) / np.maximum(np.abs(ref[0]), 1e-300)))
        if not no_i:
            ie = (d_i[sel] if pm else d_i[:, sel].T).cpu().numpy()
            err = max(err, float(np.max(np.abs(ie - ref[1]) / np.maximum(np.abs(ref[1]), 1e-300))))
        print("%-14s tier=%d fb=%d status_max=%d  ms min/med = %.3f / %.3f   %.2f M solves/s  relerr=%.2e" % (
            arg, st["tier"], st["fallback_solves"], int(d_s.max().item()), min(ts), sorted(ts)[1], P / min(ts) / 1e3, err), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
