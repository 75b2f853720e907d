"""Pins the oracle (oracle/spicey_oracle.py and oracle/oracle.c) to the reference's own
known-answer vectors, and the two restatements to each other."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import spicey_oracle as o
from spicey_b200.formatting import format_ac_result
from spicey_b200.parsing import compute_effective_time_step, parse_netlist

SVG_CASES = ["transient01_rc_pulse", "two_probes", "switch_vt_vh", "vswitch_pwl", "boost_converter_probe"]


def test_basics01_inline_snapshot_201_rows(golden):
    """tests/basics/basics01.test.ts:19-220 — character-for-character."""
    g = golden("basics01_ac")
    res = o.simulate(g["netlist"])
    lines = format_ac_result(res["ac"]).split("\n")
    assert lines[0] == g["header"]
    assert len(lines) - 1 == len(g["rows"]) == 201
    for ln, row in zip(lines[1:], g["rows"]):
        assert ln == "%s, %s,%s, %s,%s" % tuple(row)


def _px(g, times, series):
    vmin, vmax = g["v_range"]
    t0, t1 = g["t_ms_range"]
    xs = [100 + (t * 1000 - t0) / (t1 - t0) * 1052 for t in times]
    ys = [520 - (v - vmin) / (vmax - vmin) * 456 for v in series]
    return xs, ys


@pytest.mark.parametrize("name", SVG_CASES)
def test_svg_snapshot_series_python_oracle(golden, name):
    """Every spicey polyline of the five SVG snapshots, to 1e-5 px (the SVG stores 6 decimals)."""
    g = golden(name)
    tran = o.simulate(g["netlist"])["tran"]
    assert len(g["series_px"]) == len(tran["nodeVoltages"])
    for sname, pts in g["series_px"].items():
        node = sname[2:-1]
        key = [k for k in tran["nodeVoltages"] if k.upper() == node.upper()][0]
        xs, ys = _px(g, tran["times"], tran["nodeVoltages"][key])
        assert len(pts) == len(ys)
        assert max(abs(a - p[0]) for a, p in zip(xs, pts)) <= 0.0051  # x has 2 decimals
        assert max(abs(a - p[1]) for a, p in zip(ys, pts)) <= 1e-5


@pytest.mark.parametrize("name", SVG_CASES + ["diode_switch"])
def test_c_oracle_bit_identical_to_python_oracle_tran(golden, name):
    g = golden(name)
    ck = parse_netlist(g["netlist"])
    ck.probes.tran = []
    ref = o.simulate_tran(ck)
    ck2 = parse_netlist(g["netlist"])
    dt, steps = compute_effective_time_step(ck2.analyses.tran.dt, ck2.analyses.tran.tstop)
    v, ie, iters, st, state = co.tran_solve(ck2, dt, steps)
    assert st[0] == 0
    pv = np.array([ref["nodeVoltages"][n] for n in ref["nodeVoltages"]]).T
    pi = np.array([ref["elementCurrents"][n] for n in ref["elementCurrents"]]).T
    assert np.array_equal(pv, v[0])
    assert np.array_equal(pi, ie[0], equal_nan=True)
    # final state written back as the reference mutates ckt (simulateTRAN.ts:221-237)
    fin = [c.vPrev for c in ck.C] + [l.iPrev for l in ck.L] + [d.vdPrev for d in ck.D] + \
          [1.0 if s.isOn else 0.0 for s in ck.S]
    assert np.array_equal(np.array(fin), state[0])


def test_c_oracle_bit_identical_to_python_oracle_ac(golden):
    from spicey_b200 import workloads as w
    for text, stride in ((w.README_RC, 1), (w.rc_ladder(64, ppd=10), 7), (w.rc_mesh(4, ppd=5), 5)):
        ck = parse_netlist(text)
        full = o.simulate_ac(ck)
        freqs = full["freqs"][::stride]
        ref = o.simulate_ac(ck, freqs)
        x, ie, st = co.ac_solve(ck, freqs, nthreads=2)
        assert st.max() == 0
        nn = ck.nodes.count() - 1
        pv = np.array([[complex(ref["nodeVoltages"][n][k]) for n in ref["nodeVoltages"]] for k in range(len(freqs))])
        pi = np.array([[complex(ref["elementCurrents"][n][k]) for n in ref["elementCurrents"]] for k in range(len(freqs))])
        assert np.array_equal(pv, x[:, :nn])
        assert np.array_equal(pi, ie)


def test_scalar_assertions_of_reference_tests(golden):
    """vswitch-pwl.test.ts:58-76, switch-vt-vh.test.ts:33-34,61-70, two-probes.test.ts:36-37,
    boost-converter-probe.test.ts:80 (101 samples)."""
    def sampler(tran, node):
        times, v = tran["times"], tran["nodeVoltages"][node]
        return lambda target: v[min(range(len(times)), key=lambda i: abs(times[i] - target))]

    r = o.simulate(golden("vswitch_pwl")["netlist"])
    out, ctrl = sampler(r["tran"], "OUT"), sampler(r["tran"], "CTRL")
    assert ctrl(0.0005) > 2 and abs(out(0.0005)) < 0.02
    assert ctrl(0.0035) < 1 and out(0.0035) > 2
    assert ctrl(0.0045) < 2 and out(0.0045) > 4
    assert ctrl(0.0085) > 1 and abs(out(0.0085)) < 0.02
    assert abs(ctrl(0.0095)) < 5e-10 and out(0.0095) > 2

    r = o.simulate(golden("switch_vt_vh")["netlist"])
    m = r["circuit"].S[0].model
    assert abs(m.Von - 2.55) < 5e-3 and abs(m.Voff - 2.45) < 5e-3
    assert r["circuit"].probes.tran == ["n2", "nctrl_sw1"]
    n2 = sampler(r["tran"], "N2")
    assert n2(0.0002) > 4.9 and n2(0.0007) < 0.1 and n2(0.0012) > 4.9 and n2(0.0017) < 0.1

    r = o.simulate(golden("two_probes")["netlist"])
    assert sorted(r["tran"]["nodeVoltages"]) == ["1", "2"]
    assert abs(r["tran"]["nodeVoltages"]["1"][0]) < 5e-3 and abs(r["tran"]["nodeVoltages"]["2"][0]) < 5e-3

    r = o.simulate(golden("boost_converter_probe")["netlist"])
    assert len(r["tran"]["times"]) == 101
    assert all(v == 5 for v in r["tran"]["nodeVoltages"]["N1"])


def test_solver_guards():
    """solveReal.ts:28 / solveComplex.ts:29 singular throw; Complex.ts:41-42 divide guard."""
    with pytest.raises(o.SingularMatrixError):
        o.solve_real([[1.0, 2.0], [2.0, 4.0]], [1.0, 1.0])
    x, st = co.solve_real([[1.0, 2.0], [2.0, 4.0]], [1.0, 1.0])
    assert st == co.ST_SINGULAR
    with pytest.raises(o.ComplexDivideError):  # |pivot| = 1e-8 passes the 1e-15 test, |pivot|^2 = 1e-16 does not
        o.solve_complex([[o.Complex(1e-8, 0), o.Complex(0, 0)], [o.Complex(0, 0), o.Complex(1, 0)]],
                        [o.Complex(1, 0), o.Complex(1, 0)])
    x, st = co.solve_complex([[1e-8, 0], [0, 1]], [1, 1])
    assert st == co.ST_CDIV
    rng = np.random.default_rng(1)
    for n in (1, 2, 5, 17):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        x, st = co.solve_complex(A, b)
        assert st == 0 and np.allclose(A @ x, b, rtol=0, atol=1e-10)
        xr, st = co.solve_real(A.real, b.real)
        assert st == 0 and np.allclose(A.real @ xr, b.real, rtol=0, atol=1e-10)
