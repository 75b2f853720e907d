#!/usr/bin/env python3
"""GPU box with several GPUs: what limits the end-to-end (host-buffer) path when more than one GPU delivers results.

  1. raw ceiling: D devices copy 1 GiB each, device -> pinned host, concurrently (one stream per device, one process):
     aggregate GB/s for D = 1, 2, 4, 8 — the host-side ingest rate no result path can exceed;
  2. the library's own multi-device handle (spicey_create(devices, D), one process, one pinned result buffer):
     cfg2 sweep of D x 1,000,000 + 1 points through spicey_ac_solve, with element currents and with node voltages only
     (the lazy result objects compute the currents on access), solves/s and GB/s.
The 8-rank figure (one process per GPU) is bench.py's e2e at --gpus 8.  Prints one JSON document."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import spicey_b200 as sp  # noqa: E402
from spicey_b200 import native, packing, parsing, workloads  # noqa: E402


def raw_d2h(devs, gib=1.0, reps=3):
    n = int(gib * (1 << 30)) // 8
    src = [torch.empty(n, dtype=torch.float64, device="cuda:%d" % d).fill_(1.0) for d in devs]
    dst = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in devs]
    streams = [torch.cuda.Stream(device="cuda:%d" % d) for d in devs]
    best = 0.0
    for _ in range(reps):
        for d in devs:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d, s, a, b in zip(devs, streams, src, dst):
            with torch.cuda.device(d), torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
        for d in devs:
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        best = max(best, len(devs) * n * 8 / dt / 1e9)
    return best


def library_e2e(devs, want_currents):
    D = len(devs)
    ck = parsing.parse_netlist(workloads.rc_ladder(64, ppd=200000 * D))
    table = packing.pack_circuit(ck)
    freqs = np.array(sp.analysis.ac_frequencies(ck), dtype=np.float64)
    P = freqs.shape[0]
    eng = native.Engine(devs)
    h_f, p0 = native.pinned_empty(eng.lib, (P,), np.float64)
    h_f[:] = freqs
    h_x, p1 = native.pinned_empty(eng.lib, (table.nvar, P), np.complex128)
    h_i, p2 = native.pinned_empty(eng.lib, (table.n_ac_elem, P), np.complex128) if want_currents else (None, None)
    h_s, p3 = native.pinned_empty(eng.lib, (P,), np.int32)
    flags = native.FLAG_SERIES_MAJOR
    eng.ac_solve(table, h_f, out=(h_x, h_i, h_s), flags=flags, want_currents=want_currents)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        eng.ac_solve(table, h_f, out=(h_x, h_i, h_s), flags=flags, want_currents=want_currents)
        ts.append(time.perf_counter() - t0)
    st = eng.stats()
    assert int(h_s.max()) == 0
    # spot check of every device's slab against the strict dense kernel of device 0
    e0 = native.Engine([devs[0]])
    pick = np.linspace(0, P - 1, 8 * D).astype(np.int64)
    xr, _, _ = e0.ac_solve(table, freqs[pick], flags=native.FLAG_STRICT, want_currents=False)
    err = float(np.max(np.abs(h_x[:, pick].T - xr) / np.abs(xr)))
    e0.close()
    for p in (p0, p1, p2, p3):
        if p:
            eng.lib.spicey_host_free(p)
    eng.close()
    t = min(ts)
    return {"devices": D, "points": int(P), "currents": bool(want_currents), "seconds": t, "solves_per_s": P / t,
            "d2h_gbs": st["d2h_bytes"] / t / 1e9, "kernel_ms_max_over_devices": st["kernel_ms"], "tier": st["tier"],
            "max_rel_err_vs_strict": err}


def main():
    nd = torch.cuda.device_count()
    out = {"n_devices": nd, "host_cpus": len(os.sched_getaffinity(0)), "raw_d2h_gbs": {}, "library": []}
    for D in (1, 2, 4, 8):
        if D <= nd:
            out["raw_d2h_gbs"][str(D)] = raw_d2h(list(range(D)))
    for D in (1, 2, 4, 8):
        if D <= nd:
            for cur in (True, False):
                out["library"].append(library_e2e(list(range(D)), cur))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
