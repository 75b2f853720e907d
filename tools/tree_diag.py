"""GPU: per-entry error of every AC tier on random RC trees with chords (strongly attenuating networks) against the strict
dense kernel: what a renumbered elimination order costs (band_plan.h: band_order_deviation)."""
import os, sys
import numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import spicey_b200 as sp
from spicey_b200 import native, packing, parsing

def tree(n, ppd):
    rng = np.random.default_rng(n)
    lines = ["* random sparse RC network", "v1 n1 0 dc 1 ac 1"]
    for i in range(2, n + 1):
        lines.append("r%d n%d n%d %g" % (i, i, rng.integers(1, i), rng.uniform(100, 1e4)))
    for i in range(1, n + 1):
        lines.append("c%d n%d 0 %g" % (i, i, rng.uniform(1e-9, 1e-7)))
    for k in range(n // 4):
        a, b = rng.choice(np.arange(1, n + 1), 2, replace=False)
        lines.append("r%d n%d n%d %g" % (1000 + k, a, b, rng.uniform(100, 1e4)))
    lines += [".ac dec %d 1 100k" % ppd, ".end"]
    return "\n".join(lines) + "\n"

for n in (60, 100):
    text = tree(n, 600)
    ck = parsing.parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    table = packing.pack_circuit(ck)
    print("tree%d" % n, "band plan", native.band_plan_stats(table, 300.0))
    eng = native.Engine([0])
    # the reference values: the dense kernel in reference-order unfused arithmetic (SPICEY_FLAG_STRICT), which the GPU
    # tests hold to the oracle at 1e-12 (in the run kept under profiles/ the two were identical to the last bit)
    ref = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=native.FLAG_STRICT)
    xr, ier = ref["x"][0].copy(), ref["ielem"][0].copy()
    for name, flags in (("band", native.FLAG_SPARSE | native.FLAG_BAND), ("interp", native.FLAG_SPARSE | native.FLAG_NO_JIT | native.FLAG_NO_BAND),
                        ("dense", native.FLAG_DENSE), ("strict", native.FLAG_STRICT)):
        out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags)
        x = out["x"][0]; ie = out["ielem"][0]
        for lab, got, ref in (("x", x, xr), ("i", ie, ier)):
            err = np.abs(got - ref)
            per = err / np.maximum(np.abs(ref), 1e-300)
            big = np.abs(ref) > 1e-6 * np.max(np.abs(ref), axis=1, keepdims=True)
            scaled = err / np.max(np.abs(ref), axis=1, keepdims=True)
            print("  %-7s tier %d %s: per-entry max %.2e   where |ref| > 1e-6 row max: %.2e   relative to the row max: %.2e   smallest |ref| / row max %.1e" % (
                name, eng.stats()["tier"], lab, per.max(), per[big].max(), scaled.max(), (np.abs(ref) / np.max(np.abs(ref), axis=1, keepdims=True)).min()))
    eng.close()
