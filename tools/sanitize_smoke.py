"""Small end-to-end pass over every kernel tier, meant to be run under compute-sanitizer (one tool per call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import spicey_b200 as sp
from spicey_b200 import native, workloads as w
from spicey_b200.parsing import parse_netlist

eng = native.Engine()
sp.set_engine(eng)
sp.simulate(w.README_RC)
ck = parse_netlist(w.rc_ladder(64))
f = np.array(sp.analysis.ac_frequencies(ck))[::20011]
for fl in (native.FLAG_DENSE, native.FLAG_STRICT, native.FLAG_FORCE_GMEM, native.FLAG_SPARSE,
           native.FLAG_SPARSE | native.FLAG_SERIES_MAJOR):
    out = sp.simulate_ac_batch(ck, f, engine=eng, flags=fl)
    assert out["status"].max() == 0
ov = {"r1": np.linspace(900, 1100, 8), "c5": np.linspace(0.9e-9, 1.1e-9, 8)}
assert sp.simulate_ac_batch(ck, f, n_inst=8, overrides=ov, engine=eng, flags=native.FLAG_SPARSE)["status"].max() == 0
ck4 = parse_netlist(w.rc_mesh(6))
f4 = np.logspace(0, 5, 9)
for fl in (native.FLAG_DENSE, native.FLAG_SPARSE):
    assert sp.simulate_ac_batch(ck4, f4, engine=eng, flags=fl)["status"].max() == 0
boost = """* b
.MODEL D D
.MODEL SWMOD SW
LL1 N1 N2 1
DD1 N2 N3 D
CC1 N3 0 10U
RR1 N3 0 1K
SM1 N2 0 N4 0 SWMOD
V0 N1 0 DC 5
V1 N4 0 PULSE(0 10 0 1n 1n 0.00068 0.001)
.tran 0.0001 0.005
"""
for fl in (0, native.FLAG_STRICT, native.FLAG_GENERIC_THREAD, native.FLAG_FORCE_CTA, native.FLAG_FORCE_GMEM):
    r = sp.simulateTRAN(parse_netlist(boost), flags=fl)
    assert len(r["times"]) == 51
ckt = parse_netlist(w.RLC_TANK)
ov = {k: v[:100] for k, v in w.rlc_tank_overrides(65536).items()}
assert sp.simulate_tran_batch(ckt, n_inst=100, overrides=ov, engine=eng)["status"].max() == 0
print("sanitize smoke ok")
