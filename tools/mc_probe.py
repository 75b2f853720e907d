#!/usr/bin/env python3
"""GPU: AC Monte-Carlo on random RLC networks (every R, C, L +-5 %, 4,096 instances x 48 frequencies): kernel time of the
dense register tier with and without the one-warp-per-system kernel (SPICEY_WARP_LU) and of the default policy."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spicey_b200 as sp
from spicey_b200 import native
from spicey_b200.parsing import parse_netlist
def random_rlc_netlist(rng, n_nodes, n_elem, n_v=2):
    lines = ["* random RLC"]
    nodes = ["0"] + ["n%d" % i for i in range(1, n_nodes + 1)]
    for k in range(n_v):
        lines.append("v%d n%d 0 dc 1 ac %g %g" % (k + 1, k + 1, rng.uniform(0.5, 2), rng.uniform(-90, 90)))
    for i in range(1, n_nodes + 1):
        lines.append("r%d n%d %s %g" % (i, i, nodes[rng.integers(0, i)], rng.uniform(10, 1e4)))
    for k in range(n_elem):
        a, b = rng.choice(len(nodes), 2, replace=False)
        kind = "rcl"[rng.integers(0, 3)]
        val = {"r": rng.uniform(10, 1e4), "c": rng.uniform(1e-9, 1e-6), "l": rng.uniform(1e-4, 1e-2)}[kind]
        lines.append("%s%d %s %s %g" % (kind, 100 + k, nodes[a], nodes[b], val))
    lines.append(".ac dec 7 10 1meg")
    return "\n".join(lines) + "\n"
for n_nodes, n_elem in ((6, 14), (14, 60), (30, 120)):
    rng = np.random.default_rng(n_nodes)
    ck = parse_netlist(random_rlc_netlist(rng, n_nodes, n_elem))
    n = 4096
    ov = {}
    for el in list(ck.R) + list(ck.C) + list(ck.L):
        val = getattr(el, "R", None) or getattr(el, "C", None) or getattr(el, "L", None)
        ov[el.name] = val * (1 + 0.05 * rng.uniform(-1, 1, n))
    freqs = np.logspace(1, 6, 48)
    for wl in ("0", "1"):
        os.environ["SPICEY_WARP_LU"] = wl
        e = native.Engine()
        for flags in (native.FLAG_DENSE | native.FLAG_TILE | native.FLAG_SERIES_MAJOR, native.FLAG_SERIES_MAJOR):
            best = 1e9
            for _ in range(3):
                out = sp.simulate_ac_batch(ck, freqs, n_inst=n, overrides=ov, engine=e, flags=flags)
                st = e.stats(); best = min(best, st["kernel_ms"])
            print("nvar=%d warp_lu=%s flags=%s tier=%d fb=%d kernel %.3f ms  %.1f M solves/s" % (out["x"].shape[2], wl, "TILE" if flags & native.FLAG_TILE else "default", st["tier"], st["fallback_solves"], best, n * len(freqs) / best / 1e3), flush=True)
        e.close()
