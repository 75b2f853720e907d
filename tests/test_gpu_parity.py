"""Parity of the CUDA path (through the C ABI) against the oracle on the same inputs.

Bars (BASELINE.json north_star): max relative error 1e-9 for AC (node voltages and branch
currents, complex: magnitude and phase), 1e-6 for transient waveforms.  Strict mode
(SPICEY_FLAG_STRICT: reference-order unfused arithmetic) is held to 1e-12.
"""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import spicey_oracle as o
from spicey_b200 import native, workloads as w
from spicey_b200.parsing import compute_effective_time_step, parse_netlist

pytestmark = pytest.mark.gpu

AC_TOL = 1e-9
TRAN_TOL = 1e-6


@pytest.fixture(scope="module")
def eng():
    import spicey_b200 as sp
    e = native.Engine()
    sp.set_engine(e)
    yield e
    sp.set_engine(None)
    e.close()


def rel_err(a, b):
    """max |a-b| / |b| over entries, entries of b below 1e-300 compared absolutely."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b)
    den = np.where(den < 1e-300, 1.0, den)
    return float(np.max(np.abs(a - b) / den)) if a.size else 0.0


def ac_case(eng, text, freqs, flags=0, n_inst=1, overrides=None):
    import spicey_b200 as sp
    ck = parse_netlist(text)
    out = sp.simulate_ac_batch(ck, freqs, n_inst=n_inst, overrides=overrides, engine=eng, flags=flags)
    x, ie, st = co.ac_solve(ck, freqs, n_inst=n_inst, overrides=overrides, nthreads=4)
    F = len(freqs)
    return out, x.reshape(n_inst, F, -1), ie.reshape(n_inst, F, -1), st.reshape(n_inst, F), ck


def test_library_loaded_and_device_present(eng):
    assert eng.lib.spicey_device_count() >= 1
    assert eng.fp64_peak_gflops() > 1000.0


def test_simulate_readme_golden_text(eng, golden):
    """tests/basics/basics01.test.ts through the drop-in simulate(): 201 rows, character for character."""
    import spicey_b200 as sp
    g = golden("basics01_ac")
    res = sp.simulate(g["netlist"])
    assert res["tran"] is None
    lines = sp.formatAcResult(res["ac"]).split("\n")
    assert lines[0] == g["header"] and len(lines) == 202
    for ln, row in zip(lines[1:], g["rows"]):
        assert ln == "%s, %s,%s, %s,%s" % tuple(row)
    assert list(res["ac"]["elementCurrents"].keys()) == ["r1", "c1", "v1"]


@pytest.mark.parametrize("flags,tol", [(0, AC_TOL), (native.FLAG_STRICT, 1e-12), (native.FLAG_FORCE_GMEM, AC_TOL)])
def test_ac_readme_rc(eng, flags, tol):
    ck = parse_netlist(w.README_RC)
    import spicey_b200 as sp
    freqs = sp.analysis.ac_frequencies(ck)
    out, x, ie, st, _ = ac_case(eng, w.README_RC, freqs, flags)
    assert out["status"].max() == 0 and st.max() == 0
    assert rel_err(out["x"], x) <= tol
    assert rel_err(out["ielem"], ie) <= tol


SM = native.FLAG_SERIES_MAJOR


@pytest.mark.parametrize("flags,tol", [(native.FLAG_DENSE, AC_TOL), (native.FLAG_STRICT, 1e-12),
                                       (native.FLAG_FORCE_GMEM, AC_TOL), (native.FLAG_SPARSE, AC_TOL),
                                       (native.FLAG_DENSE | SM, AC_TOL), (native.FLAG_SPARSE | SM, AC_TOL),
                                       (native.FLAG_SPARSE | native.FLAG_JIT, AC_TOL),
                                       (native.FLAG_SPARSE | native.FLAG_JIT | SM, AC_TOL),
                                       (native.FLAG_SPARSE | native.FLAG_WARP, AC_TOL),
                                       (native.FLAG_SPARSE | native.FLAG_WARP | SM, AC_TOL),
                                       (native.FLAG_FORCE_GMEM | native.FLAG_STRICT | SM, 1e-12)])
def test_ac_ladder64_slice(eng, flags, tol):
    """cfg 2 topology (Nvar = 65), every 997th of the 1,000,001 frequencies."""
    import spicey_b200 as sp
    text = w.rc_ladder(64)
    freqs = np.array(sp.analysis.ac_frequencies(parse_netlist(text)))
    assert freqs.shape[0] == 1000001
    sub = freqs[::997]
    out, x, ie, st, _ = ac_case(eng, text, sub, flags)
    if flags & native.FLAG_SPARSE:
        stt = eng.stats()
        want = (native.TIER_SPARSE_JIT if flags & native.FLAG_JIT else
                native.TIER_SPARSE_WARP if flags & native.FLAG_WARP else native.TIER_SPARSE)
        assert stt["tier"] == want and stt["fallback_solves"] == 0 and stt["program_cfma"] > 0
    assert out["status"].max() == 0 and st.max() == 0
    assert rel_err(out["x"], x) <= tol, rel_err(out["x"], x)
    assert rel_err(out["ielem"], ie) <= tol
    # magnitude and phase separately, as the north star words it
    assert rel_err(np.abs(out["x"]), np.abs(x)) <= tol
    dphi = np.angle(out["x"] * np.conj(x))
    assert np.max(np.abs(dphi)) <= tol


@pytest.mark.parametrize("flags,tier", [(native.FLAG_DENSE, native.TIER_CTA_GMEM),
                                        (native.FLAG_SPARSE, native.TIER_SPARSE_WARP),
                                        (native.FLAG_SPARSE | native.FLAG_SERIES_MAJOR, native.TIER_SPARSE_WARP),
                                        (native.FLAG_SPARSE | native.FLAG_NO_WARP, native.TIER_SPARSE)])
def test_ac_mesh16_slice(eng, flags, tier):
    """cfg 4 topology (Nvar = 257): does not fit one SM's shared memory -> global-scratch tier (dense), the
    warp-per-system sparse tier (default for programs this large) or the thread-per-system sparse tier."""
    import spicey_b200 as sp
    text = w.rc_mesh(16)
    freqs = np.array(sp.analysis.ac_frequencies(parse_netlist(text)))
    assert freqs.shape[0] == 8000001
    sub = freqs[::200003]
    out, x, ie, st, _ = ac_case(eng, text, sub, flags)
    assert eng.stats()["tier"] == tier
    assert out["status"].max() == 0 and st.max() == 0
    assert rel_err(out["x"], x) <= AC_TOL, rel_err(out["x"], x)
    assert rel_err(out["ielem"], ie) <= AC_TOL


FALLBACKS = []


def random_rlc_netlist(rng, n_nodes, n_elem, n_v=2):
    lines = ["* random RLC"]
    nodes = ["0"] + ["n%d" % i for i in range(1, n_nodes + 1)]
    for k in range(n_v):
        lines.append("v%d n%d 0 dc 1 ac %g %g" % (k + 1, k + 1, rng.uniform(0.5, 2), rng.uniform(-90, 90)))
    for i in range(1, n_nodes + 1):  # a resistor tree keeps every node connected
        lines.append("r%d n%d %s %g" % (i, i, nodes[rng.integers(0, i)], rng.uniform(10, 1e4)))
    for k in range(n_elem):
        a, b = rng.choice(len(nodes), 2, replace=False)
        kind = "rcl"[rng.integers(0, 3)]
        val = {"r": rng.uniform(10, 1e4), "c": rng.uniform(1e-9, 1e-6), "l": rng.uniform(1e-4, 1e-2)}[kind]
        lines.append("%s%d %s %s %g" % (kind, 100 + k, nodes[a], nodes[b], val))
    lines.append(".ac dec 7 10 1meg")
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("n_nodes,n_elem", [(1, 0), (2, 3), (5, 12), (14, 60), (30, 200), (31, 40), (62, 300), (100, 700)])
def test_ac_random_rlc_networks(eng, n_nodes, n_elem):
    """Dense-ish random networks: exercises pivoting away from the diagonal, fill-in and ties."""
    import spicey_b200 as sp
    rng = np.random.default_rng(n_nodes * 1000 + n_elem)
    text = random_rlc_netlist(rng, n_nodes, n_elem, n_v=min(2, n_nodes))
    freqs = sp.analysis.ac_frequencies(parse_netlist(text))
    for flags, tol in ((native.FLAG_DENSE, AC_TOL), (native.FLAG_STRICT, 1e-11), (native.FLAG_SPARSE, AC_TOL),
                       (native.FLAG_SPARSE | SM, AC_TOL), (native.FLAG_SPARSE | native.FLAG_JIT | SM, AC_TOL),
                       (native.FLAG_SPARSE | native.FLAG_WARP, AC_TOL), (native.FLAG_SPARSE | native.FLAG_WARP | SM, AC_TOL)):
        out, x, ie, st, _ = ac_case(eng, text, freqs, flags)
        if flags == native.FLAG_SPARSE:
            FALLBACKS.append(eng.stats()["fallback_solves"])
        if flags & native.FLAG_WARP:
            assert eng.stats()["tier"] in (native.TIER_SPARSE_WARP, native.TIER_SPARSE, native.TIER_CTA_SMEM, native.TIER_CTA_GMEM, native.TIER_THREAD)
        assert np.array_equal(out["status"], st)
        assert st.max() == 0
        scale = np.max(np.abs(x), axis=2, keepdims=True)  # mixed-magnitude solutions: error relative to the row's max
        assert np.max(np.abs(out["x"] - x) / scale) <= tol
        iscale = np.max(np.abs(ie), axis=2, keepdims=True)
        assert np.max(np.abs(out["ielem"] - ie) / iscale) <= tol


def test_sparse_fallback_path_was_exercised():
    """Across the random networks the pivot order changes along the 5-decade sweep for some points:
    those must have gone through the dense fallback (and matched the oracle above)."""
    assert len(FALLBACKS) >= 8 and sum(FALLBACKS) > 0, FALLBACKS


def test_ac_sweep_instances(eng):
    """Monte-Carlo axis on AC: 37 instances x 11 frequencies with swept R/C/source."""
    rng = np.random.default_rng(5)
    n = 37
    ov = {"r1": rng.uniform(10, 100, n), "c1": rng.uniform(1e-5, 1e-3, n), "v1.acmag": rng.uniform(0.5, 2, n),
          "v1.acphase": rng.uniform(-180, 180, n)}
    freqs = np.logspace(0, 3, 11)
    for flags, tier in ((0, native.TIER_CTA_SMEM), (native.FLAG_SPARSE, native.TIER_SPARSE),
                        (native.FLAG_SPARSE | native.FLAG_JIT, native.TIER_SPARSE_JIT),
                        (native.FLAG_SPARSE | native.FLAG_JIT | SM, native.TIER_SPARSE_JIT)):
        out, x, ie, st, _ = ac_case(eng, w.README_RC, freqs, flags, n_inst=n, overrides=ov)
        assert eng.stats()["tier"] == tier
        assert out["status"].max() == 0
        assert rel_err(out["x"], x) <= AC_TOL
        assert rel_err(out["ielem"], ie) <= AC_TOL


@pytest.mark.parametrize("flags", [native.FLAG_SPARSE, native.FLAG_SPARSE | SM, 0, native.FLAG_SPARSE | native.FLAG_JIT,
                                   native.FLAG_SPARSE | native.FLAG_JIT | SM])
def test_ac_monte_carlo_ladder_sparse_eager(eng, flags):
    """Component-tolerance Monte-Carlo on the AC axis (64-node ladder, 40 instances x 53 frequencies, every
    R and C swept): the sparse program with per-instance (eager) stamping, interpreted or compiled; flags=0
    checks the automatic tier choice (2120 points >= 2048 -> sparse)."""
    text = w.rc_ladder(64)
    n = 40
    u = w.splitmix_uniform_pm1(n, 126)
    ov = {}
    for k in range(1, 64):
        ov["r%d" % k] = 1000.0 * (1 + 0.05 * u[:, k - 1])
        ov["c%d" % k] = 1e-9 * (1 + 0.05 * u[:, 62 + k])
    freqs = np.logspace(0, 5, 53)
    out, x, ie, st, _ = ac_case(eng, text, freqs, flags, n_inst=n, overrides=ov)
    assert eng.stats()["tier"] == (native.TIER_SPARSE_JIT if flags & native.FLAG_JIT else native.TIER_SPARSE)
    assert out["status"].max() == 0 and st.max() == 0
    assert rel_err(out["x"], x) <= AC_TOL
    assert rel_err(out["ielem"], ie) <= AC_TOL


def test_ac_sweep_random_rlc_compiled_per_instance_stamping(eng):
    """Random RLC networks with every R, C and L value and one source phasor swept over 12 instances: the compiled
    kernel with per-instance stamping (tier 5) — inductor admittances and their divide guards, two sources,
    pivoting away from the diagonal — against the oracle, point-major and series-major."""
    for n_nodes, n_elem in ((5, 12), (14, 40), (24, 70)):
        rng = np.random.default_rng(77 + n_nodes)
        text = random_rlc_netlist(rng, n_nodes, n_elem, n_v=2)
        ck = parse_netlist(text)
        n = 12
        ov = {}
        for el in list(ck.R) + list(ck.C) + list(ck.L):
            val = getattr(el, "R", None) or getattr(el, "C", None) or getattr(el, "L", None)
            ov[el.name] = val * (1 + 0.2 * rng.uniform(-1, 1, n))
        ov["v1.acmag"] = rng.uniform(0.5, 2, n)
        ov["v1.acphase"] = rng.uniform(-90, 90, n)
        freqs = np.logspace(1, 6, 9)
        for flags in (native.FLAG_SPARSE | native.FLAG_JIT, native.FLAG_SPARSE | native.FLAG_JIT | SM):
            out, x, ie, st, _ = ac_case(eng, text, freqs, flags, n_inst=n, overrides=ov)
            # (the largest network may exceed what one thread can hold: the interpreted program then runs)
            assert eng.stats()["tier"] == native.TIER_SPARSE_JIT or n_nodes == 24, eng.stats()
            assert np.array_equal(out["status"], st) and st.max() == 0
            scale = np.max(np.abs(x), axis=2, keepdims=True)
            assert np.max(np.abs(out["x"] - x) / scale) <= AC_TOL
            iscale = np.max(np.abs(ie), axis=2, keepdims=True)
            assert np.max(np.abs(out["ielem"] - ie) / iscale) <= AC_TOL


def test_ac_sweep_sparse_bad_instances_fall_back(eng):
    """Sweep with one R<=0 instance and random RLC values: failures and pivot changes go through the dense
    fallback with exact statuses."""
    n = 16
    rng = np.random.default_rng(11)
    r = rng.uniform(10, 100, n)
    r[5] = 0.0
    ov = {"r1": r, "c1": rng.uniform(1e-5, 1e-3, n)}
    for flags in (native.FLAG_SPARSE, native.FLAG_SPARSE | native.FLAG_JIT):
        out, x, ie, st, _ = ac_case(eng, w.README_RC, np.logspace(0, 3, 7), flags, n_inst=n, overrides=ov)
        assert np.array_equal(out["status"], st) and (st[5] == native.ST_R_NONPOS).all()
        ok = [i for i in range(n) if i != 5]
        assert rel_err(out["x"][ok], x[ok]) <= AC_TOL and rel_err(out["ielem"][ok], ie[ok]) <= AC_TOL
        assert eng.stats()["fallback_solves"] == 7


def test_ac_error_statuses_do_not_poison_batch(eng):
    """R<=0 (simulateAC.ts:37), singular matrix (solveComplex.ts:29), Complex.div guard (Complex.ts:42)."""
    n = 8
    r = np.full(n, 30.0)
    r[3] = -1.0
    out, x, ie, st, _ = ac_case(eng, w.README_RC, [1.0, 10.0], n_inst=n, overrides={"r1": r})
    assert np.array_equal(out["status"], st)
    assert (st[3] == native.ST_R_NONPOS).all() and st[[0, 1, 2, 4, 5, 6, 7]].max() == 0
    ok = [0, 1, 2, 4, 5, 6, 7]
    assert rel_err(out["x"][ok], x[ok]) <= AC_TOL
    assert np.isnan(out["x"][3]).all()
    # two ideal sources in parallel -> singular
    text = "* sing\nv1 a 0 ac 1\nv2 a 0 ac 1\nr1 a 0 1k\n.ac lin 2 1 2\n"
    out, x, ie, st, _ = ac_case(eng, text, [1.0, 2.0])
    assert (st == native.ST_SINGULAR).all() and np.array_equal(out["status"], st)
    # inductor with 1e-15 <= 2*pi*f*L < 3.2e-8 -> "Complex divide by ~0" (hazard H5)
    text = "* cdiv\nv1 a 0 ac 1\nl1 a b 1e-10\nr1 b 0 1k\n.ac lin 2 1 2\n"
    for flags in (0, native.FLAG_SPARSE, native.FLAG_SPARSE | SM, native.FLAG_SPARSE | native.FLAG_JIT):
        out, x, ie, st, _ = ac_case(eng, text, [1.0, 1e6], flags)
        assert st[0, 0] == native.ST_CDIV and st[0, 1] == 0 and np.array_equal(out["status"], st)
        assert rel_err(out["x"][0, 1], x[0, 1]) <= AC_TOL
    for text in ("* sing\nv1 a 0 ac 1\nv2 a 0 ac 1\nr1 a 0 1k\n.ac lin 2 1 2\n",
                 "* bad\nv1 1 0 ac 1\nr1 1 2 -5\nc1 2 0 1u\n.ac dec 2 1 10\n"):
        out, x, ie, st, _ = ac_case(eng, text, [1.0, 2.0], native.FLAG_SPARSE)
        assert st.min() > 0 and np.array_equal(out["status"], st)
    import spicey_b200 as sp
    with pytest.raises(ValueError, match="R r1 must be > 0"):
        sp.simulate("* bad\nv1 1 0 ac 1\nr1 1 2 -5\nc1 2 0 1u\n.ac dec 2 1 10\n")
    with pytest.raises(ArithmeticError, match=r"Singular matrix \(complex\)"):
        sp.simulate("* sing\nv1 a 0 ac 1\nv2 a 0 ac 1\nr1 a 0 1k\n.ac lin 2 1 2\n")


# ---- transient --------------------------------------------------------------------

GOLDEN_TRAN = ["transient01_rc_pulse", "two_probes", "switch_vt_vh", "vswitch_pwl", "boost_converter_probe",
               "diode_switch", "case_insensitive_nodes"]


@pytest.mark.parametrize("name", GOLDEN_TRAN)
@pytest.mark.parametrize("flags", [0, native.FLAG_STRICT, native.FLAG_GENERIC_THREAD,
                                   native.FLAG_GENERIC_THREAD | native.FLAG_STRICT, native.FLAG_FORCE_CTA,
                                   native.FLAG_FORCE_GMEM, native.FLAG_JIT])
def test_tran_reference_netlists(eng, golden, name, flags):
    """Every transient netlist of the reference's tests through the drop-in simulateTRAN, all tiers
    (FLAG_JIT: the kernel compiled for the netlist, tier 6)."""
    import spicey_b200 as sp
    text = golden(name)["netlist"]
    ref = o.simulate(text)
    ck = parse_netlist(text)
    got = sp.simulateTRAN(ck, flags=flags)
    if flags == native.FLAG_JIT:
        assert eng.stats()["tier"] == native.TIER_TRAN_JIT
    assert got["times"] == ref["tran"]["times"]
    assert list(got["nodeVoltages"].keys()) == list(ref["tran"]["nodeVoltages"].keys())
    assert list(got["elementCurrents"].keys()) == list(ref["tran"]["elementCurrents"].keys())
    for kind in ("nodeVoltages", "elementCurrents"):
        for k, b in ref["tran"][kind].items():
            a, b = np.asarray(got[kind][k]), np.asarray(b)
            scale = max(1e-30, float(np.max(np.abs(b))))
            assert np.max(np.abs(a - b)) <= TRAN_TOL * scale, (kind, k, np.max(np.abs(a - b)) / scale)
    # circuit state is left as the reference leaves it (simulateTRAN.ts:221-237)
    rc = ref["circuit"]
    for a, b in zip(ck.C, rc.C):
        assert abs(a.vPrev - b.vPrev) <= TRAN_TOL * max(1.0, abs(b.vPrev))
    for a, b in zip(ck.L, rc.L):
        assert abs(a.iPrev - b.iPrev) <= TRAN_TOL * max(1.0, abs(b.iPrev))
    for a, b in zip(ck.S, rc.S):
        assert a.isOn == b.isOn


def test_tran_svg_golden_pixels(eng, golden):
    """GPU waveforms against the reference's SVG snapshots directly (1e-4 px at 6 decimals)."""
    import spicey_b200 as sp
    for name in ["transient01_rc_pulse", "switch_vt_vh", "vswitch_pwl", "boost_converter_probe"]:
        g = golden(name)
        tran = sp.simulate(g["netlist"])["tran"]
        vmin, vmax = g["v_range"]
        for sname, pts in g["series_px"].items():
            key = [k for k in tran["nodeVoltages"] if k.upper() == sname[2:-1].upper()][0]
            ys = 520 - (np.asarray(tran["nodeVoltages"][key]) - vmin) / (vmax - vmin) * 456
            assert np.max(np.abs(ys - np.array(pts)[:, 1])) <= 1e-4


def tran_batch_case(eng, text, n_inst, overrides, flags=0):
    import spicey_b200 as sp
    ck = parse_netlist(text)
    got = sp.simulate_tran_batch(ck, n_inst=n_inst, overrides=overrides, engine=eng, flags=flags, want_iters=True)
    ck2 = parse_netlist(text)
    dt, steps = compute_effective_time_step(ck2.analyses.tran.dt, ck2.analyses.tran.tstop)
    v, ie, iters, st, state = co.tran_solve(ck2, dt, steps, n_inst=n_inst, overrides=overrides, nthreads=8)
    return got, v, ie, iters, st


@pytest.mark.parametrize("flags", [0, native.FLAG_GENERIC_THREAD, native.FLAG_FORCE_CTA, native.FLAG_JIT])
def test_tran_rlc_tank_monte_carlo_slice(eng, flags):
    """cfg 3 on its first 512 instances: steps = 1001 (hazard H2), +-5 % R/L/C."""
    n = 64 if flags == native.FLAG_FORCE_CTA else 512
    ov = {k: v[:n] for k, v in w.rlc_tank_overrides(65536).items()}
    got, v, ie, iters, st = tran_batch_case(eng, w.RLC_TANK, n, ov, flags)
    assert eng.stats()["tier"] == (native.TIER_TRAN_JIT if flags == native.FLAG_JIT else
                                   native.TIER_CTA_SMEM if flags == native.FLAG_FORCE_CTA else native.TIER_THREAD)
    assert got["steps"] == 1001 and got["v"].shape == (1002, 2, n)
    assert got["status"].max() == 0 and st.max() == 0
    assert np.array_equal(got["iters"].T, iters)
    ref_v = np.transpose(v, (1, 2, 0))
    ref_i = np.transpose(ie, (1, 2, 0))
    assert np.max(np.abs(got["v"] - ref_v)) <= TRAN_TOL * np.max(np.abs(ref_v))
    assert np.max(np.abs(got["ielem"] - ref_i)) <= TRAN_TOL * np.max(np.abs(ref_i))


@pytest.mark.parametrize("flags", [0, native.FLAG_GENERIC_THREAD, native.FLAG_FORCE_CTA, native.FLAG_JIT])
def test_tran_rectifier_sweep_slice(eng, flags):
    """cfg 5 (diode, single linearisation per step) on 400 instances spread over the sweep."""
    n = 48 if flags == native.FLAG_FORCE_CTA else 400
    full = w.rectifier_overrides(100000)
    pick = np.linspace(0, 99999, n).astype(int)
    ov = {k: v[pick] for k, v in full.items()}
    got, v, ie, iters, st = tran_batch_case(eng, w.RECTIFIER, n, ov, flags)
    if flags == native.FLAG_JIT:
        assert eng.stats()["tier"] == native.TIER_TRAN_JIT
    assert got["steps"] == 3000
    assert got["status"].max() == 0 and st.max() == 0
    ref_v = np.transpose(v, (1, 2, 0))
    ref_i = np.transpose(ie, (1, 2, 0))
    vs = np.max(np.abs(ref_v), axis=0, keepdims=True)
    assert np.max(np.abs(got["v"] - ref_v) / vs) <= TRAN_TOL
    cs = np.maximum(np.max(np.abs(ref_i), axis=0, keepdims=True), 1e-30)
    assert np.max(np.abs(got["ielem"] - ref_i) / cs) <= TRAN_TOL


# ---- source waveforms evaluated on the device ---------------------------------

WAVE_FLAGS = [0, native.FLAG_STRICT, native.FLAG_GENERIC_THREAD, native.FLAG_FORCE_CTA, native.FLAG_JIT]


@pytest.mark.parametrize("name", ["transient01_rc_pulse", "vswitch_pwl", "boost_converter_probe", "switch_vt_vh"])
@pytest.mark.parametrize("flags", WAVE_FLAGS)
def test_tran_device_waveforms_equal_presampled_rows(eng, golden, name, flags):
    """pulseValue / pwlValue evaluated by the kernels (t = step*dt, separately rounded operations) against the
    rows the host samples with the reference's own functions: the waveforms of the reference's tests, every tier.
    The source values are bit-identical, so the recorded source node voltage is too."""
    import spicey_b200 as sp
    text = golden(name)["netlist"]
    a = sp.simulate_tran_batch(parse_netlist(text), engine=eng, flags=flags, want_iters=True, device_waves=False)
    tier_rows = eng.stats()["tier"]
    b = sp.simulate_tran_batch(parse_netlist(text), engine=eng, flags=flags, want_iters=True, device_waves=True)
    assert eng.stats()["tier"] == tier_rows
    if flags == native.FLAG_JIT:
        assert tier_rows == native.TIER_TRAN_JIT
    assert b["status"].max() == 0 and np.array_equal(a["iters"], b["iters"])
    assert np.array_equal(a["v"], b["v"]) and np.array_equal(a["ielem"], b["ielem"])
    assert np.array_equal(a["state"], b["state"])


def _per_instance_oracle(text, n, mutate, overrides=None):
    """The reference has no sweep API: one oracle run per instance on a circuit whose parsed PULSE / PWL (the
    object the waveform closure reads) is changed in place."""
    vs, ies = [], []
    for i in range(n):
        ck = parse_netlist(text)
        mutate(ck, i)
        dt, steps = compute_effective_time_step(ck.analyses.tran.dt, ck.analyses.tran.tstop)
        ov = {k: np.asarray(v)[i:i + 1] for k, v in (overrides or {}).items()}
        v, ie, iters, st, _ = co.tran_solve(ck, dt, steps, n_inst=1, overrides=ov or None)
        assert st.max() == 0
        vs.append(v[0]); ies.append(ie[0])
    return np.stack(vs, axis=2), np.stack(ies, axis=2)   # [S1, rows, n]


@pytest.mark.parametrize("flags", WAVE_FLAGS)
def test_tran_pulse_parameter_sweep_per_instance(eng, flags):
    """A sweep over the SOURCE: amplitude, delay, rise time, width and period of a PULSE differ per instance (plus R),
    evaluated on the device; each instance against its own oracle run."""
    import spicey_b200 as sp
    text = "* pulse sweep\nV1 in 0 PULSE(0 5 1u 0.5u 0.5u 2u 8u)\nR1 in out 1k\nC1 out 0 1n\nL1 out x 10u\nR2 x 0 50\n.tran 0.1u 30u\n"
    n = 24 if flags == native.FLAG_FORCE_CTA else 96
    rng = np.random.default_rng(7)
    ov = {"v1.pulse.v1": rng.uniform(-1, 1, n), "v1.pulse.v2": rng.uniform(2, 6, n), "v1.pulse.td": rng.uniform(0, 3e-6, n),
          "v1.pulse.tr": rng.choice([0.0, 0.1e-6, 0.5e-6, 1e-6], n), "v1.pulse.tf": rng.choice([0.0, 0.3e-6, 0.5e-6], n),
          "v1.pulse.ton": rng.uniform(0.5e-6, 3e-6, n), "v1.pulse.period": rng.uniform(5e-6, 12e-6, n),
          "v1.pulse.ncycles": rng.choice([1.0, 2.0, np.inf], n), "R1": rng.uniform(500, 2000, n)}
    ov["v1.pulse.period"][0] = 0.0     # JavaScript semantics of tt / 0 (Infinity / NaN): the pulse never fires
    got = sp.simulate_tran_batch(parse_netlist(text), n_inst=n, overrides=ov, engine=eng, flags=flags)
    if flags == native.FLAG_JIT:
        assert eng.stats()["tier"] == native.TIER_TRAN_JIT
    assert got["status"].max() == 0

    def mutate(ck, i):
        p = ck.V[0].pulse
        for nm in ("v1", "v2", "td", "tr", "tf", "ton", "period", "ncycles"):
            setattr(p, nm, float(ov["v1.pulse." + nm][i]))

    ref_v, ref_i = _per_instance_oracle(text, n, mutate, {"R1": ov["R1"]})
    in_row = [k.lower() for k in got["node_names"]].index("in")
    assert np.array_equal(got["v"][:, in_row, :], ref_v[:, in_row, :])     # the source itself: bit for bit
    assert np.max(np.abs(got["v"][:, in_row, 0])) == abs(ov["v1.pulse.v1"][0])
    tol = 1e-12 if flags & native.FLAG_STRICT else TRAN_TOL
    assert np.max(np.abs(got["v"] - ref_v)) <= tol * np.max(np.abs(ref_v))
    assert np.max(np.abs(got["ielem"] - ref_i)) <= tol * np.max(np.abs(ref_i))


@pytest.mark.parametrize("flags", WAVE_FLAGS)
def test_tran_pwl_parameter_sweep_with_switch(eng, golden, flags):
    """The reference's PWL-driven voltage switch with the PWL breakpoints and levels varied per instance: the
    switch toggles at different steps in different instances (per-instance re-solve masks)."""
    import spicey_b200 as sp
    text = golden("vswitch_pwl")["netlist"]
    ck0 = parse_netlist(text)
    k = [i for i, v in enumerate(ck0.V) if v.pwl is not None][0]
    name = ck0.V[k].name.lower()
    pairs = ck0.V[k].pwl
    n = 16 if flags == native.FLAG_FORCE_CTA else 64
    rng = np.random.default_rng(11)
    tscale = rng.uniform(0.7, 1.3, n)
    vscale = rng.uniform(0.8, 1.2, n)
    ov = {}
    for q, (t, v) in enumerate(pairs):
        ov["%s.pwl.t%d" % (name, q)] = t * tscale
        ov["%s.pwl.v%d" % (name, q)] = v * vscale
    got = sp.simulate_tran_batch(parse_netlist(text), n_inst=n, overrides=ov, engine=eng, flags=flags, want_iters=True)
    assert got["status"].max() == 0
    assert len(set(map(tuple, got["iters"].T))) > 1      # instances really toggle at different steps

    def mutate(ck, i):
        ck.V[k].pwl[:] = [(float(ov["%s.pwl.t%d" % (name, q)][i]), float(ov["%s.pwl.v%d" % (name, q)][i]))
                          for q in range(len(pairs))]

    ref_v, ref_i = _per_instance_oracle(text, n, mutate)
    tol = 1e-12 if flags & native.FLAG_STRICT else TRAN_TOL
    assert np.max(np.abs(got["v"] - ref_v)) <= tol * np.max(np.abs(ref_v))
    assert np.max(np.abs(got["ielem"] - ref_i)) <= tol * np.max(np.abs(ref_i))


def random_linear_tran_netlist(rng, n_nodes, n_elem):
    """Random connected R / C / L network driven by one PULSE and one PWL source (parsed by the reference grammar)."""
    nodes = ["n%d" % i for i in range(1, n_nodes + 1)]
    lines = ["* random transient network", "V1 n1 0 PULSE(0 %g 1u 0.5u 0.7u 3u 9u)" % rng.uniform(1, 5)]
    if n_nodes >= 2:
        lines.append("V2 n2 0 PWL(0 0 2u %g 5u %g 11u 0 30u %g)" % (rng.uniform(0.5, 2), rng.uniform(-1, 1), rng.uniform(0.5, 2)))
    for i in range(n_nodes):   # spanning chain of resistors, every node has a resistive path to a source
        lines.append("r%d %s %s %g" % (i, nodes[i], nodes[i + 1] if i + 1 < n_nodes else "0", rng.uniform(100, 5e3)))
    for k in range(n_elem):
        a = rng.integers(0, n_nodes)
        b = rng.integers(-1, n_nodes)
        if a == b:
            b = -1
        kind = "rcl"[rng.integers(0, 3)]
        val = {"r": rng.uniform(100, 1e4), "c": rng.uniform(1e-10, 1e-8), "l": rng.uniform(1e-5, 1e-3)}[kind]
        lines.append("%s%d %s %s %g" % (kind, 100 + k, nodes[a], "0" if b < 0 else nodes[b], val))
    lines.append(".tran 0.25u 40u")
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("n_nodes,n_elem", [(1, 2), (3, 5), (5, 9), (10, 20), (22, 50)])
def test_tran_random_linear_networks_with_swept_sources(eng, n_nodes, n_elem):
    """Random R / C / L networks with a PULSE and a PWL source whose parameters (and two resistors) differ per
    instance: every transient tier the system size admits (register-resident, compiled, generic thread, CTA with
    the matrix in shared memory / in the global scratch), device-evaluated sources, one oracle run per instance."""
    import spicey_b200 as sp
    rng = np.random.default_rng(500 + n_nodes)
    text = random_linear_tran_netlist(rng, n_nodes, n_elem)
    n = 12
    ov = {"v1.pulse.v2": rng.uniform(1, 5, n), "v1.pulse.td": rng.uniform(0, 4e-6, n), "v1.pulse.period": rng.uniform(6e-6, 15e-6, n),
          "r0": rng.uniform(100, 5e3, n)}
    if n_nodes >= 2:
        ov.update({"v2.pwl.t1": rng.uniform(1e-6, 4e-6, n), "v2.pwl.v2": rng.uniform(-1, 1, n), "r1": rng.uniform(100, 5e3, n)})

    def mutate(ck, i):
        p = ck.V[0].pulse
        p.v2, p.td, p.period = (float(ov["v1.pulse." + k][i]) for k in ("v2", "td", "period"))
        if n_nodes >= 2:
            q = ck.V[1].pwl
            q[1] = (float(ov["v2.pwl.t1"][i]), q[1][1])
            q[2] = (q[2][0], float(ov["v2.pwl.v2"][i]))

    ref_v, ref_i = _per_instance_oracle(text, n, mutate, {k: v for k, v in ov.items() if k.startswith("r")})
    nvar = n_nodes + min(2, n_nodes)
    seen = set()
    for flags in (0, native.FLAG_JIT, native.FLAG_GENERIC_THREAD, native.FLAG_FORCE_CTA, native.FLAG_FORCE_GMEM, native.FLAG_STRICT):
        got = sp.simulate_tran_batch(parse_netlist(text), n_inst=n, overrides=ov, engine=eng, flags=flags)
        seen.add(eng.stats()["tier"])
        assert got["status"].max() == 0
        tol = 1e-11 if flags & native.FLAG_STRICT else TRAN_TOL
        assert np.array_equal(got["v"][:, 0, :], ref_v[:, 0, :])            # node n1 is the PULSE source itself
        assert np.max(np.abs(got["v"] - ref_v)) <= tol * np.max(np.abs(ref_v)), flags
        assert np.max(np.abs(got["ielem"] - ref_i)) <= tol * np.max(np.abs(ref_i)), flags
    assert native.TIER_CTA_SMEM in seen and native.TIER_CTA_GMEM in seen
    assert (native.TIER_TRAN_JIT in seen) == (nvar <= 8) and (native.TIER_THREAD in seen) == (nvar <= 16)


def test_tran_waves_argument_errors(eng):
    from spicey_b200 import packing
    ck = parse_netlist("* t\nV1 1 0 PULSE(0 5 0 1n 1n 5u 10u)\nR1 1 2 1k\nC1 2 0 1u\n.tran 0.1u 2u\n")
    table = packing.pack_circuit(ck, device_waves=True)
    for waves, msg in ((native.Waves([native.WAVE_PULSE, 0], [0, 0], [0, 0]), "n_vsrc differs"),
                       (native.Waves([7], [0], [0]), "unknown waveform kind"),
                       (native.Waves([native.WAVE_PULSE], [len(table.values) - 7], [0]), "slots out of range"),
                       (native.Waves([native.WAVE_TABLE], [0], [0]), "needs vsrc")):
        with pytest.raises(native.NativeError, match=msg):
            eng.tran_solve(table, 1e-7, 20, waves=waves)


def test_tran_singular_instance_is_isolated(eng):
    """A singular instance (two sources fighting) reports status 1 and NaNs; neighbours are untouched."""
    import spicey_b200 as sp
    text = "* t\nV1 1 0 PULSE(0 5 0 1n 1n 5u 10u)\nR1 1 2 1k\nC1 2 0 1u\n.tran 0.1u 2u\n"
    r = np.array([1000.0, 1000.0, 1000.0])
    got, v, ie, iters, st = tran_batch_case(eng, text, 3, {"R1": r})
    assert got["status"].max() == 0
    with pytest.raises(ArithmeticError, match=r"Singular matrix \(real\)"):
        sp.simulate("* sing\nv1 a 0 dc 1\nv2 a 0 dc 2\nr1 a 0 1k\n.tran 1u 5u\n")


def test_full_size_properties_cfg2(eng):
    """BASELINE cfg 2 at full size (1,000,001 points), size-independent properties:
    every status 0; V(n1) equals the source phasor exactly; KCL at the source node
    (i_v1 = -i_r1); |V| decreases monotonically along the ladder; a 1/997 subsample matches
    the oracle at 1e-9."""
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_ladder(64))
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=SM)
    stt = eng.stats()
    assert stt["tier"] == native.TIER_SPARSE_JIT and stt["fallback_solves"] == 0   # 1e6 points: compiled program
    out_i = sp.simulate_ac_batch(ck, freqs[::7], engine=eng, flags=SM | native.FLAG_NO_JIT)
    assert eng.stats()["tier"] == native.TIER_SPARSE
    assert rel_err(out_i["x"][0], out["x"][0][::7]) <= 1e-12   # interpreter and compiled program agree
    assert out["status"].max() == 0
    x, ie = out["x"][0], out["ielem"][0]
    assert np.max(np.abs(x[:, 0] - 1.0)) <= 1e-15
    assert np.max(np.abs(ie[:, -1] + ie[:, 0])) <= 1e-12 * np.max(np.abs(ie[:, 0]))
    mags = np.abs(x[:, :64])
    assert np.all(np.diff(mags, axis=1) <= 1e-12 * mags[:, :-1])
    sub = slice(0, None, 997)
    xr, ier, st = co.ac_solve(ck, freqs[sub], nthreads=8)
    assert rel_err(x[sub], xr) <= AC_TOL


def test_full_size_properties_cfg3(eng):
    """BASELINE cfg 3 at full size (65,536 instances x 1,002 recorded steps) with the default tier policy: the
    compiled transient kernel (tier 6) takes it; every status 0; the source node follows the pulse exactly;
    KCL at node 2 (i_R = i_L + i_C) at every step; instances spread over the batch match the oracle at 1e-6;
    and the generic kernel (NO_JIT) agrees on a slice."""
    import spicey_b200 as sp
    n = 65536
    ov = w.rlc_tank_overrides(n)
    ck = parse_netlist(w.RLC_TANK)
    got = sp.simulate_tran_batch(ck, n_inst=n, overrides=ov, engine=eng)
    assert eng.stats()["tier"] == native.TIER_TRAN_JIT
    assert got["steps"] == 1001 and got["v"].shape == (1002, 2, n) and got["status"].max() == 0
    names = list(got["element_names"])
    iR, iL, iC = (got["ielem"][:, names.index(k), :] for k in ("R1", "L1", "C1"))
    scale = np.max(np.abs(iR))
    assert np.max(np.abs(iR - iL - iC)) <= 1e-9 * scale
    vsrc = got["v"][:, 0, :]
    assert np.all(vsrc[0] == 0.0) and np.max(np.abs(vsrc[1:] - 1.0)) == 0.0    # PULSE(0 1 0 1n 1n 1 2), dt ~ 1 us
    pick = np.linspace(0, n - 1, 48).astype(int)
    ck2 = parse_netlist(w.RLC_TANK)
    dt, steps = compute_effective_time_step(ck2.analyses.tran.dt, ck2.analyses.tran.tstop)
    v, ie, iters, st, state = co.tran_solve(ck2, dt, steps, n_inst=len(pick), overrides={k: a[pick] for k, a in ov.items()},
                                            nthreads=8)
    ref_v, ref_i = np.transpose(v, (1, 2, 0)), np.transpose(ie, (1, 2, 0))
    assert np.max(np.abs(got["v"][:, :, pick] - ref_v)) <= TRAN_TOL * np.max(np.abs(ref_v))
    assert np.max(np.abs(got["ielem"][:, :, pick] - ref_i)) <= TRAN_TOL * np.max(np.abs(ref_i))
    m = 2048
    gen = sp.simulate_tran_batch(parse_netlist(w.RLC_TANK), n_inst=m, overrides={k: a[:m] for k, a in ov.items()},
                                 engine=eng, flags=native.FLAG_NO_JIT)
    assert eng.stats()["tier"] == native.TIER_THREAD
    assert np.max(np.abs(gen["v"] - got["v"][:, :, :m])) <= 1e-9 * np.max(np.abs(gen["v"]))


def test_large_slice_properties_cfg5(eng):
    """BASELINE cfg 5 (diode rectifier) on 16,384 instances spread over the 100,000-instance sweep: 49 M
    instance-steps, so the default policy compiles (tier 6); status 0; the diode current is never below -Is (no
    reverse conduction), the recorded R and C currents follow from the recorded voltages; a subsample matches
    the oracle at 1e-6."""
    import spicey_b200 as sp
    n = 16384
    full = w.rectifier_overrides(100000)
    pick = np.linspace(0, 99999, n).astype(int)
    ov = {k: v[pick] for k, v in full.items()}
    got = sp.simulate_tran_batch(parse_netlist(w.RECTIFIER), n_inst=n, overrides=ov, engine=eng)
    assert eng.stats()["tier"] == native.TIER_TRAN_JIT
    assert got["steps"] == 3000 and got["status"].max() == 0
    names = list(got["element_names"])
    iD, iR, iC = (got["ielem"][:, names.index(k), :] for k in ("D1", "R1", "C1"))
    assert np.all(iD >= -1.0001 * ov["D1.is"][None, :])
    # (no KCL check at the output node: the reference records the diode current from the unclamped exponential
    #  at the new solution while the step was solved with the model linearised about vdPrev — hazard H6)
    vout = got["v"][:, list(got["node_names"]).index("out"), :]
    assert np.max(np.abs(iR - vout / ov["R1"][None, :])) <= 1e-12 * np.max(np.abs(iR))
    assert np.max(np.abs(iC[1:] - ov["C1"][None, :] * (vout[1:] - vout[:-1]) / got["dt"])) <= 1e-9 * np.max(np.abs(iC))
    sub = np.linspace(0, n - 1, 40).astype(int)
    ck2 = parse_netlist(w.RECTIFIER)
    dt, steps = compute_effective_time_step(ck2.analyses.tran.dt, ck2.analyses.tran.tstop)
    v, ie, iters, st, state = co.tran_solve(ck2, dt, steps, n_inst=len(sub), overrides={k: a[sub] for k, a in ov.items()},
                                            nthreads=8)
    ref_v, ref_i = np.transpose(v, (1, 2, 0)), np.transpose(ie, (1, 2, 0))
    vs = np.max(np.abs(ref_v), axis=0, keepdims=True)
    assert np.max(np.abs(got["v"][:, :, sub] - ref_v) / vs) <= TRAN_TOL
    cs = np.maximum(np.max(np.abs(ref_i), axis=0, keepdims=True), 1e-30)
    assert np.max(np.abs(got["ielem"][:, :, sub] - ref_i) / cs) <= TRAN_TOL


def test_full_size_properties_cfg5(eng):
    """BASELINE cfg 5 at full size: 100,000 instances x 3,001 recorded steps (14.4 GB of results), default tier
    policy (compiled, tier 6): every status 0, no reverse conduction beyond -Is, the recorded R and C currents follow
    from the recorded voltages for EVERY instance, and 256 instances spread over the sweep match the oracle at 1e-6."""
    import spicey_b200 as sp
    n = 100000
    ov = w.rectifier_overrides(n)
    got = sp.simulate_tran_batch(parse_netlist(w.RECTIFIER), n_inst=n, overrides=ov, engine=eng)
    assert eng.stats()["tier"] == native.TIER_TRAN_JIT
    assert got["steps"] == 3000 and got["v"].shape == (3001, 2, n) and got["status"].max() == 0
    names = list(got["element_names"])
    iD, iR, iC = (got["ielem"][:, names.index(k), :] for k in ("D1", "R1", "C1"))
    assert np.all(iD >= -1.0001 * ov["D1.is"][None, :])
    vout = got["v"][:, list(got["node_names"]).index("out"), :]
    assert np.max(np.abs(iR - vout / ov["R1"][None, :])) <= 1e-12 * np.max(np.abs(iR))
    assert np.max(np.abs(iC[1:] - ov["C1"][None, :] * (vout[1:] - vout[:-1]) / got["dt"])) <= 1e-9 * np.max(np.abs(iC))
    sub = np.linspace(0, n - 1, 256).astype(int)
    ck2 = parse_netlist(w.RECTIFIER)
    dt, steps = compute_effective_time_step(ck2.analyses.tran.dt, ck2.analyses.tran.tstop)
    v, ie, iters, st, state = co.tran_solve(ck2, dt, steps, n_inst=len(sub), overrides={k: a[sub] for k, a in ov.items()},
                                            nthreads=8)
    assert st.max() == 0
    ref_v, ref_i = np.transpose(v, (1, 2, 0)), np.transpose(ie, (1, 2, 0))
    vs = np.max(np.abs(ref_v), axis=0, keepdims=True)
    assert np.max(np.abs(got["v"][:, :, sub] - ref_v) / vs) <= TRAN_TOL
    cs = np.maximum(np.max(np.abs(ref_i), axis=0, keepdims=True), 1e-30)
    assert np.max(np.abs(got["ielem"][:, :, sub] - ref_i) / cs) <= TRAN_TOL


def test_full_size_properties_cfg4(eng):
    """BASELINE cfg 4 at full size: the 8,000,001-point sweep of the 16x16 mesh through the device entry point in
    chunks of 500,000 points (the results of a chunk are 8 GB; the whole sweep would be 127 GB), default tier policy
    (>= 200,000 points: the banded tier).  Per chunk: every status 0, no point handed to the fallback, V(n0_0) equals
    the source phasor exactly, KCL at the source node (i_v1 = -(i_r1 + i_r2)); 512 points spread over the whole sweep
    match the oracle at 1e-9 per entry."""
    import torch
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_mesh(16))
    table = sp.packing.pack_circuit(ck)
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    P, chunk = freqs.shape[0], 500000
    assert P == 8000001
    dev = torch.device("cuda", 0)
    ld = eng.series_ld(chunk)
    d_f = torch.from_numpy(freqs).to(dev)
    d_x = torch.empty((table.nvar, ld), dtype=torch.complex128, device=dev)
    d_i = torch.empty((table.n_ac_elem, ld), dtype=torch.complex128, device=dev)
    d_s = torch.empty(chunk, dtype=torch.int32, device=dev)
    names = table.names[:table.n_ac_elem]
    src_node = list(ck.nodes.rev[1:]).index("n0_0")
    at_src = [j for j, nm in enumerate(names) if nm[0] == "r" and src_node + 1 in (int(table.n1[j]), int(table.n2[j]))]
    assert len(at_src) == 2
    sgn = torch.tensor([1.0 if int(table.n1[j]) == src_node + 1 else -1.0 for j in at_src], dtype=torch.complex128, device=dev)
    pick_all = np.linspace(0, P - 1, 512).astype(np.int64)
    got_x, got_i = [], []
    stream = torch.cuda.current_stream()
    for c0 in range(0, P, chunk):
        n = min(chunk, P - c0)
        eng.ac_solve_device(table, d_f.data_ptr() + 8 * c0, n, d_x.data_ptr(), d_i.data_ptr(), d_s.data_ptr(),
                            flags=SM | native.FLAG_JIT, stream=stream.cuda_stream, series_ld=ld)
        torch.cuda.synchronize()
        stt = eng.stats()
        # (the last chunk is the sweep's final single point: batches below 2,048 points take the dense kernel)
        assert stt["tier"] == (native.TIER_BAND if n >= 2048 else native.TIER_CTA_GMEM) and stt["fallback_solves"] == 0, stt
        assert int(d_s[:n].max().item()) == 0
        assert float((d_x[src_node, :n] - 1.0).abs().max().item()) <= 1e-15
        iv = d_i[names.index("v1"), :n]
        ir = (d_i[at_src, :n] * sgn[:, None]).sum(dim=0)
        # (the currents are differences of node voltages near 1 V: 1e-6 A over 1 kOhm is a 1e-3 V difference, so a
        #  relative 1e-16 on the voltages is 1e-13 on a current, and the residual is a sum of three of them)
        assert float(((iv + ir).abs() / ir.abs()).max().item()) <= 1e-9
        loc = pick_all[(pick_all >= c0) & (pick_all < c0 + n)] - c0
        sel = torch.from_numpy(loc).to(dev)
        got_x.append(d_x[:, sel].T.cpu().numpy())
        got_i.append(d_i[:, sel].T.cpu().numpy())
    xr, ir_, st = co.ac_solve(ck, freqs[pick_all], nthreads=8)
    assert st.max() == 0
    assert rel_err(np.concatenate(got_x), xr) <= AC_TOL and rel_err(np.concatenate(got_i), ir_) <= AC_TOL


def _solve_extended(A, b):
    """Gaussian elimination with partial pivoting in numpy's extended precision (the yardstick for per-entry errors)."""
    A = A.astype(np.clongdouble).copy()
    b = b.astype(np.clongdouble).copy()
    n = A.shape[0]
    for k in range(n):
        p = k + int(np.argmax(np.abs(A[k:, k])))
        if p != k:
            A[[k, p]] = A[[p, k]]
            b[[k, p]] = b[[p, k]]
        f = A[k + 1:, k] / A[k, k]
        A[k + 1:, k:] -= f[:, None] * A[k, k:][None, :]
        b[k + 1:] -= f * b[k]
    x = np.zeros(n, dtype=np.clongdouble)
    for i in range(n - 1, -1, -1):
        x[i] = (b[i] - A[i, i + 1:] @ x[i + 1:]) / A[i, i]
    return x


@pytest.mark.parametrize("n_nodes,n_elem", [(14, 60), (30, 200), (62, 300)])
def test_ac_random_rlc_networks_per_entry(eng, n_nodes, n_elem):
    """Per-entry relative errors on the ill-conditioned random networks (the row-maximum scale of the test above hides
    small entries): against the same system solved in extended precision, every entry above 1e-6 of its row's maximum
    is as accurate as the oracle's own double-precision answer (within 4x, or 1e-9), on every tier."""
    import spicey_b200 as sp
    rng = np.random.default_rng(n_nodes * 1000 + n_elem)
    text = random_rlc_netlist(rng, n_nodes, n_elem, n_v=2)
    ck = parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck))[::6]
    xo, _, st = co.ac_solve(ck, freqs)
    assert st.max() == 0
    nvar = xo.shape[1]
    exact = []
    for f in freqs:
        A, b = o.build_linear_system_for_ac(ck, float(f), nvar)
        exact.append(_solve_extended(np.array([[complex(z) for z in row] for row in A]), np.array([complex(z) for z in b])))
    exact = np.array(exact)
    big = np.abs(exact) > 1e-6 * np.max(np.abs(exact), axis=1, keepdims=True)
    e_oracle = float(np.max((np.abs(xo - exact) / np.abs(exact))[big]))
    for flags in (native.FLAG_DENSE, native.FLAG_SPARSE, native.FLAG_SPARSE | native.FLAG_JIT | SM, native.FLAG_SPARSE | native.FLAG_WARP, BAND):
        out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags)
        assert out["status"].max() == 0
        e_gpu = float(np.max((np.abs(out["x"][0] - exact) / np.abs(exact))[big]))
        assert e_gpu <= max(1e-9, 4 * e_oracle), (flags, e_gpu, e_oracle)


def test_mesh_default_policy_takes_the_warp_tier(eng):
    """cfg 4 topology, 4,096 frequencies, default flags: the sparse path engages (>= 2048 points), the program is
    large (thread-tier workspace >= 512 slots), so the warp-per-system tier runs; KCL at the source node and a
    subsample against the oracle."""
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_mesh(16))
    freqs = np.array(sp.analysis.ac_frequencies(ck))[::1953][:4096]
    out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=SM)
    stt = eng.stats()
    assert stt["tier"] == native.TIER_SPARSE_WARP and stt["fallback_solves"] == 0 and out["status"].max() == 0
    x = out["x"][0]
    assert np.max(np.abs(x[:, 0] - 1.0)) <= 1e-15              # n0_0 is the source node
    sub = slice(0, None, 64)
    xr, ier, st = co.ac_solve(ck, freqs[sub], nthreads=8)
    assert rel_err(x[sub], xr) <= AC_TOL and rel_err(out["ielem"][0][sub], ier) <= AC_TOL


BAND = native.FLAG_SPARSE | native.FLAG_BAND


@pytest.mark.parametrize("flags", [BAND, BAND | SM])
def test_ac_mesh16_band_tier(eng, flags):
    """cfg 4 topology through the banded + bordered tier (tier 8): the nodes are renumbered to half-bandwidth 16
    (netlist numbering: 30), 8 lanes x 2 rows per system; 4,097 points spread over the 8,000,001-point sweep,
    every node voltage and element current against the oracle at 1e-9, no system flagged for the fallback."""
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_mesh(16))
    st_plan = native.band_plan_stats(sp.packing.pack_circuit(ck), 300.0)
    assert st_plan["window"] == 16 and st_plan["lanes"] == 8 and st_plan["rows_per_lane"] == 2 and st_plan["renumbered"] == 1
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    assert freqs.shape[0] == 8000001
    sub = np.ascontiguousarray(freqs[::1953][:4097])
    out = sp.simulate_ac_batch(ck, sub, engine=eng, flags=flags)
    stt = eng.stats()
    assert stt["tier"] == native.TIER_BAND and stt["fallback_solves"] == 0 and stt["program_cfma"] > 0, stt
    assert out["status"].max() == 0
    x, ie, st = co.ac_solve(ck, sub, nthreads=8)
    assert st.max() == 0
    assert rel_err(out["x"][0], x) <= AC_TOL, rel_err(out["x"][0], x)
    assert rel_err(out["ielem"][0], ie) <= AC_TOL, rel_err(out["ielem"][0], ie)
    assert rel_err(np.abs(out["x"][0]), np.abs(x)) <= AC_TOL
    assert np.max(np.abs(np.angle(out["x"][0] * np.conj(x)))) <= AC_TOL
    # without element currents: the other compiled variant
    out2 = sp.simulate_ac_batch(ck, sub[:300], engine=eng, flags=flags, want_currents=False)
    assert eng.stats()["tier"] == native.TIER_BAND and np.array_equal(out2["x"], out["x"][:, :300])


@pytest.mark.parametrize("name,text,shape", [
    ("mesh8", w.rc_mesh(8, ppd=50), (8, 1)), ("mesh5", w.rc_mesh(5, ppd=50), (8, 1)), ("mesh3", w.rc_mesh(3, ppd=50), (4, 1)),
    ("ladder64", w.rc_ladder(64, ppd=50), (2, 1)), ("ladder400", w.rc_ladder(400, ppd=50), (2, 1)),
    ("ladder3", w.rc_ladder(3, ppd=50), (2, 1))])
def test_ac_band_tier_shapes(eng, name, text, shape):
    """Every lanes x rows-per-lane shape the plan picks: narrow bands use fewer lanes per system."""
    import spicey_b200 as sp
    ck = parse_netlist(text)
    stp = native.band_plan_stats(sp.packing.pack_circuit(ck), 300.0)
    assert (stp["lanes"], stp["rows_per_lane"]) == shape, stp
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    for flags in (BAND, BAND | SM):
        out, x, ie, st, _ = ac_case(eng, text, freqs, flags)
        assert eng.stats()["tier"] == native.TIER_BAND, eng.stats()
        assert out["status"].max() == 0 and st.max() == 0
        assert rel_err(out["x"], x) <= AC_TOL, (name, rel_err(out["x"], x))
        assert rel_err(out["ielem"], ie) <= AC_TOL


@pytest.mark.parametrize("shape", ["16,1", "32,1", "4,4", "16,2"])
def test_ac_band_tier_forced_shapes_mesh16(shape, monkeypatch):
    """The same mesh with other lane / row splits of the window (SPICEY_BAND_SHAPE): 16 x 1, 32 x 1 (window 32), 4 x 4."""
    import spicey_b200 as sp
    monkeypatch.setenv("SPICEY_BAND_SHAPE", shape)
    e = native.Engine()
    try:
        ck = parse_netlist(w.rc_mesh(16))
        freqs = np.ascontiguousarray(np.array(sp.analysis.ac_frequencies(ck))[::40001])
        out = sp.simulate_ac_batch(ck, freqs, engine=e, flags=BAND)
        stt = e.stats()
        if stt["tier"] != native.TIER_BAND:
            pytest.skip("shape %s does not fit the register file / shared memory: %s" % (shape, stt))
        x, ie, st = co.ac_solve(ck, freqs, nthreads=8)
        assert out["status"].max() == 0 and stt["fallback_solves"] == 0
        assert rel_err(out["x"][0], x) <= AC_TOL and rel_err(out["ielem"][0], ie) <= AC_TOL
    finally:
        e.close()


@pytest.mark.parametrize("name,text", [("mesh16", w.rc_mesh(16)), ("mesh5", w.rc_mesh(5, ppd=50))])
def test_ac_band_tier_tensor_store_equals_plain_stores(name, text, monkeypatch):
    """The banded tier writes its pivot rows to the workspace through TMA tensor stores by default (BAND_UMODE 2,
    band_kernel.cuh); SPICEY_BAND_UMODE=0 compiles the per-lane store form.  Same arithmetic, same order: every node
    voltage, element current and status bit-identical, no fallback in either."""
    import spicey_b200 as sp
    ck = parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    freqs = np.ascontiguousarray(freqs[:: max(1, freqs.shape[0] // 3001)])
    res = {}
    for umode in ("2", "0"):
        monkeypatch.setenv("SPICEY_BAND_UMODE", umode)
        e = native.Engine()
        try:
            out = sp.simulate_ac_batch(ck, freqs, engine=e, flags=BAND)
            stt = e.stats()
            assert stt["tier"] == native.TIER_BAND and stt["fallback_solves"] == 0, stt
            res[umode] = out
        finally:
            e.close()
    assert res["2"]["status"].max() == 0
    for key in ("x", "ielem", "status"):
        assert np.array_equal(res["2"][key], res["0"][key]), (name, key)


def test_attenuating_network_keeps_the_netlist_order(eng):
    """A random RC tree with chords (60 nodes; at 100 kHz its far nodes sit 1e-12 below the source) qualifies for the
    banded tier only after a renumbering, and in that order its small unknowns lose per-entry agreement with the
    reference (1e-8; 3e-13 relative to the largest unknown).  The host checks a renumbered plan against the netlist order
    before using it (band_plan.h: band_order_deviation) and leaves this circuit to a tier that eliminates in the
    reference's order: EVERY entry within 1e-9 of the oracle.  SPICEY_FLAG_BAND still forces the banded tier, whose
    result is then as good as backward stability makes it."""
    import spicey_b200 as sp
    from test_band_plan_host import _rc_tree
    text = _rc_tree(60, ppd=600)
    ck = parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    xr, ier, st = co.ac_solve(ck, freqs, nthreads=8)
    assert st.max() == 0
    out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=native.FLAG_SPARSE | native.FLAG_JIT)
    stt = eng.stats()
    assert stt["tier"] != native.TIER_BAND, stt
    assert out["status"].max() == 0
    assert np.max(np.abs(out["x"][0] - xr) / np.abs(xr)) <= AC_TOL
    assert np.max(np.abs(out["ielem"][0] - ier) / np.maximum(np.abs(ier), 1e-300)) <= AC_TOL
    forced = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=BAND)
    assert eng.stats()["tier"] == native.TIER_BAND and forced["status"].max() == 0
    assert np.max(np.abs(forced["x"][0] - xr) / np.max(np.abs(xr), axis=1, keepdims=True)) <= 1e-11


def test_ac_band_tier_pivot_changes_fall_back(eng):
    """An RLC ladder with two sources swept over seven decades: the pivot order of the pilot point does not hold
    everywhere, those points go to the dense kernel; statuses and values equal the oracle's either way."""
    import spicey_b200 as sp
    lines = ["* rlc ladder", "v1 n1 0 ac 1", "v2 n40 n39 ac 0.5 30"]
    for k in range(1, 60):
        lines.append("r%d n%d n%d %g" % (k, k, k + 1, 50 + 7 * (k % 5)))
        lines.append("l%d n%d n%d %g" % (k, k, k + 2 if k + 2 <= 60 else 0, 1e-3 * (1 + k % 3)))
        lines.append("c%d n%d 0 %g" % (k, k + 1, 1e-8 * (1 + k % 4)))
    lines += [".ac dec 40 1 10meg", ".end"]
    text = "\n".join(lines) + "\n"
    freqs = np.array(sp.analysis.ac_frequencies(parse_netlist(text)))
    out, x, ie, st, _ = ac_case(eng, text, freqs, BAND)
    stt = eng.stats()
    assert stt["tier"] == native.TIER_BAND, stt
    assert np.array_equal(out["status"], st) and st.max() == 0
    scale = np.max(np.abs(x), axis=2, keepdims=True)
    assert np.max(np.abs(out["x"] - x) / scale) <= AC_TOL
    iscale = np.max(np.abs(ie), axis=2, keepdims=True)
    assert np.max(np.abs(out["ielem"] - ie) / iscale) <= AC_TOL
    assert 0 < stt["fallback_solves"] < freqs.shape[0], stt


NORTON = """* current source
i1 0 n1 dc 1m ac 2m 30
r0 n1 0 500
r1 n1 n2 1k
c1 n2 0 1u
l1 n2 n3 1m
r2 n3 0 50
.ac dec 5 10 100k
.tran 10u 5m
.end
"""
THEVENIN = NORTON.replace("i1 0 n1 dc 1m ac 2m 30\nr0 n1 0 500", "v1 nx 0 dc 0.5 ac 1 30\nr0 nx n1 500")


def test_current_sources_ac_and_tran(eng):
    """SPICEY_ELEM_I (north star: "R, L, C, V and I"): the reference ships stampCurrentReal.ts / stampCurrentComplex.ts
    but parses no I line, so the element enters through parse_netlist(current_sources=True).  Checked against the
    oracle extended by the same two stamps, and against the Thevenin-equivalent V + R netlist the reference does
    parse (every node voltage of the rest of the circuit is the same)."""
    import spicey_b200 as sp
    ck = parse_netlist(NORTON, current_sources=True)
    assert len(ck.I) == 1 and ck.I[0].dc == 1e-3 and ck.I[0].acMag == 2e-3 and ck.I[0].acPhaseDeg == 30
    assert len(parse_netlist(NORTON).I) == 0 and len(parse_netlist(NORTON).skipped) == 1   # the reference's behaviour
    ref = o.simulate_ac(parse_netlist(NORTON, current_sources=True))
    ckt = parse_netlist(THEVENIN)
    freqs = np.array(ref["freqs"])
    xt, _, stt = co.ac_solve(ckt, freqs)
    names = ["n1", "n2", "n3"]
    for flags in (0, native.FLAG_STRICT, native.FLAG_SPARSE, native.FLAG_SPARSE | native.FLAG_JIT | SM, BAND, BAND | SM,
                  native.FLAG_SPARSE | native.FLAG_WARP):
        out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags)
        assert out["status"].max() == 0, (flags, eng.stats())
        for j, nm in enumerate(out["node_names"]):
            want = np.array([complex(z) for z in ref["nodeVoltages"][nm]])
            assert rel_err(out["x"][0][:, j], want) <= AC_TOL, (flags, nm)
        for j, nm in enumerate(out["element_names"]):
            want = np.array([complex(z) for z in ref["elementCurrents"][nm]])
            assert rel_err(out["ielem"][0][:, j], want) <= AC_TOL, (flags, nm)
        # Thevenin equivalent: node order there is nx, n1, n2, n3
        tn = ckt.nodes.rev[1:]
        for nm in names:
            assert rel_err(out["x"][0][:, out["node_names"].index(nm)], xt[:, tn.index(nm)]) <= AC_TOL, (flags, nm)
    assert stt.max() == 0
    # a sweep of the source magnitude (per-instance values): circuits with current sources stay with the dense kernel
    mags = np.array([1e-3, 2e-3, 4e-3])
    outs = sp.simulate_ac_batch(ck, freqs, n_inst=3, overrides={"i1.acmag": mags}, engine=eng)
    assert outs["status"].max() == 0
    base = outs["x"][1]
    assert rel_err(outs["x"][0] * 2, base) <= 1e-12 and rel_err(outs["x"][2] / 2, base) <= 1e-12
    # transient: constant 1 mA into R || C ... ; its own value is recorded as the element current
    reft = o.simulate_tran(parse_netlist(NORTON, current_sources=True))
    reftt = o.simulate_tran(parse_netlist(THEVENIN))
    for flags in (0, native.FLAG_FORCE_CTA, native.FLAG_STRICT):
        got = sp.simulate_tran_batch(parse_netlist(NORTON, current_sources=True), engine=eng, flags=flags)
        assert got["status"].max() == 0
        for j, nm in enumerate(got["node_names"]):
            want = np.asarray(reft["nodeVoltages"][nm])
            assert np.max(np.abs(got["v"][:, j, 0] - want)) <= TRAN_TOL * max(1.0, np.max(np.abs(want))), (flags, nm)
            if nm in names:
                wt = np.asarray(reftt["nodeVoltages"][nm])
                assert np.max(np.abs(got["v"][:, j, 0] - wt)) <= TRAN_TOL * max(1.0, np.max(np.abs(wt))), (flags, nm)
        for j, nm in enumerate(got["element_names"]):
            want = np.asarray(reft["elementCurrents"][nm])
            assert np.max(np.abs(got["ielem"][:, j, 0] - want)) <= TRAN_TOL * max(1e-30, np.max(np.abs(want))), (flags, nm)
        assert got["element_names"][-1] == "i1" and np.all(got["ielem"][:, -1, 0] == 1e-3)


def test_simulate_ac_lazy_currents_match_the_oracle(eng):
    """simulateAC(lazy_currents=True): the host call moves the solution vector only, element currents are computed on
    access with the reference's formula (simulateAC.ts:94-126) — same keys, same order, values within 1e-9 of the
    oracle (R, C, L and V elements, one element to ground on either side)."""
    import spicey_b200 as sp
    text = "* rlc\nv1 in 0 ac 1 15\nr1 in a 50\nl1 a b 1m\nc1 b 0 1u\nr2 0 b 2k\nc2 a b 10n\n.ac dec 20 10 1meg\n.end\n"
    ref = o.simulate(text)["ac"]
    got = sp.simulateAC(parse_netlist(text), engine=eng, lazy_currents=True)
    eager = sp.simulateAC(parse_netlist(text), engine=eng)
    assert eng.stats()["d2h_bytes"] > 0
    assert list(got["elementCurrents"].keys()) == list(ref["elementCurrents"].keys()) == list(eager["elementCurrents"].keys())
    for nm, series in ref["elementCurrents"].items():
        want = np.array([complex(z) for z in series])
        assert rel_err(got["elementCurrents"][nm].array, want) <= AC_TOL, nm
        assert rel_err(eager["elementCurrents"][nm].array, want) <= AC_TOL, nm
        z = got["elementCurrents"][nm][3]
        assert abs(complex(z) - want[3]) <= AC_TOL * abs(want[3]) and len(got["elementCurrents"][nm]) == len(want)
    assert sp.formatAcResult(got) == sp.formatAcResult(eager)


@pytest.mark.parametrize("n", [100, 150])
def test_mid_size_ladder_compiled_with_its_global_column(eng, n):
    """Ladders past cfg 2's size (Nvar 101 / 151: 201 / 301 values cross from the elimination into the
    back-substitution, 75 fit shared memory and 40 registers) still take the compiled straight-line tier: the
    longest-lived values go to the kernel's [slot][thread] column of global memory (sparse_codegen.h, JitArgs.work).
    Values within 1e-9 of the oracle, both result layouts, with and without element currents; the interpreted
    program of the same topology agrees with it to rounding."""
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_ladder(n, ppd=400))
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    sub = slice(0, None, 37)
    xr, ier, st = co.ac_solve(ck, freqs[sub], nthreads=8)
    assert st.max() == 0
    for flags in (native.FLAG_SPARSE | native.FLAG_JIT, native.FLAG_SPARSE | native.FLAG_JIT | SM):
        out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags)
        stt = eng.stats()
        assert stt["tier"] == native.TIER_SPARSE_JIT and stt["fallback_solves"] == 0, stt
        assert out["status"].max() == 0
        assert rel_err(out["x"][0][sub], xr) <= AC_TOL, rel_err(out["x"][0][sub], xr)
        assert rel_err(out["ielem"][0][sub], ier) <= AC_TOL
        out2 = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags, want_currents=False)
        assert eng.stats()["tier"] == native.TIER_SPARSE_JIT and np.array_equal(out2["x"], out["x"])
    interp = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=native.FLAG_SPARSE | native.FLAG_NO_JIT)
    assert eng.stats()["tier"] == native.TIER_SPARSE
    assert rel_err(out["x"][0], interp["x"][0]) <= 1e-12


def test_long_ladder_default_policy(eng):
    """A 400-node ladder (Nvar = 401, 3,200 points): chain-like, so the warp tier declines (2-3 updates per row
    would idle the lanes), and its program of 5,199 micro-ops is past what the straight-line compiled tier takes
    (kJitMaxOps: such a kernel compiles for more than a minute), so that tier declines.  The banded tier could take it with two
    lanes per system (half-bandwidth 1) but loses to the interpreted thread-per-system program below half-bandwidth 4
    (measured: 14 against 38 M solves/s), so the default stays with the program whatever is asked for, and the banded
    tier runs only when forced."""
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_ladder(400, ppd=640))
    freqs = np.array(sp.analysis.ac_frequencies(ck))[:3200]
    xr, ier, st = co.ac_solve(ck, freqs[::50], nthreads=8)
    for flags, tier in ((SM, native.TIER_SPARSE), (SM | native.FLAG_SPARSE | native.FLAG_JIT, native.TIER_SPARSE),
                        (SM | BAND, native.TIER_BAND)):
        out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags)
        assert eng.stats()["tier"] == tier, (flags, eng.stats())
        assert out["status"].max() == 0
        assert rel_err(out["x"][0][::50], xr) <= AC_TOL and rel_err(out["ielem"][0][::50], ier) <= AC_TOL


def test_multi_device_handle_shards_contiguous_ranges(eng):
    """spicey_create with every device of the box (2, 4 or 8): the library shards the batch axis in contiguous
    ranges (no collective), the result equals the single-device one bit for bit — compiled ladder kernel (tier 5),
    banded mesh kernel (tier 8), point-major and series-major, a transient Monte-Carlo batch — and the caller's
    current device is left alone.  Skipped on a one-GPU box."""
    ndev = eng.lib.spicey_device_count()
    if ndev < 2:
        pytest.skip("needs 2 GPUs")
    import torch
    import spicey_b200 as sp
    e2 = native.Engine(list(range(ndev)))
    try:
        ck = parse_netlist(w.rc_ladder(64))
        freqs = np.array(sp.analysis.ac_frequencies(ck))[::97]
        for flags in (SM, 0, SM | native.FLAG_SPARSE | native.FLAG_JIT):
            a = sp.simulate_ac_batch(ck, freqs, engine=e2, flags=flags)
            assert e2.stats()["n_devices"] == ndev and torch.cuda.current_device() == 0
            b = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags)
            assert e2.stats()["tier"] == eng.stats()["tier"]
            assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["ielem"], b["ielem"]) and a["status"].max() == 0
        ckm = parse_netlist(w.rc_mesh(16))
        fm = np.ascontiguousarray(np.array(sp.analysis.ac_frequencies(ckm))[::3907])
        a = sp.simulate_ac_batch(ckm, fm, engine=e2, flags=BAND | SM)
        assert e2.stats()["tier"] == native.TIER_BAND
        b = sp.simulate_ac_batch(ckm, fm, engine=eng, flags=BAND | SM)
        assert np.array_equal(a["x"], b["x"]) and np.array_equal(a["ielem"], b["ielem"]) and a["status"].max() == 0
        xr, ir, st = co.ac_solve(ckm, fm[::64], nthreads=8)
        assert rel_err(a["x"][0][::64], xr) <= AC_TOL and rel_err(a["ielem"][0][::64], ir) <= AC_TOL
        n = 4096
        ov = {k: v[:n] for k, v in w.rlc_tank_overrides(65536).items()}
        ck = parse_netlist(w.RLC_TANK)
        ta = sp.simulate_tran_batch(ck, n_inst=n, overrides=ov, engine=e2)
        tb = sp.simulate_tran_batch(ck, n_inst=n, overrides=ov, engine=eng)
        assert np.array_equal(ta["v"], tb["v"]) and np.array_equal(ta["ielem"], tb["ielem"])
        assert torch.cuda.current_device() == 0
    finally:
        e2.close()


def test_tran_second_call_continues_from_mutated_state(eng, golden):
    """The reference mutates ckt (vPrev/iPrev/vdPrev/isOn, simulateTRAN.ts:221-237): a second simulateTRAN on the
    same ParsedCircuit continues from the end state instead of restarting.  Same here."""
    import spicey_b200 as sp
    text = golden("boost_converter_probe")["netlist"]
    ref = o.simulate(text)
    ref2 = o.simulate_tran(ref["circuit"])
    ck = parse_netlist(text)
    sp.simulateTRAN(ck)
    got2 = sp.simulateTRAN(ck)
    for k, b in ref2["nodeVoltages"].items():
        a, b = np.asarray(got2["nodeVoltages"][k]), np.asarray(b)
        assert np.max(np.abs(a - b)) <= TRAN_TOL * max(1e-30, np.max(np.abs(b)))


def test_edge_cases_and_argument_errors(eng):
    """Degenerate and maximum sizes, optional outputs, call-level errors (no crash, clear message)."""
    import spicey_b200 as sp
    from spicey_b200.packing import pack_circuit
    # smallest system: one node, one source (Nvar = 2), single frequency, no element currents requested
    ck = parse_netlist("* one\nv1 a 0 ac 2 45\nr1 a 0 10\n.ac lin 1 5 5\n")
    out = sp.simulate_ac_batch(ck, [5.0], want_currents=False, engine=eng)
    assert out["ielem"] is None and out["status"][0, 0] == 0
    assert abs(out["x"][0, 0, 0] - 2 * np.exp(1j * np.pi / 4)) < 1e-15
    # a netlist without analyses: simulate() returns None for both, like the reference (:63, :131)
    res = sp.simulate("* none\nv1 a 0 dc 1\nr1 a 0 1\n")
    assert res["ac"] is None and res["tran"] is None
    # empty frequency list / zero steps are argument errors, not crashes
    table = pack_circuit(ck)
    with pytest.raises(native.NativeError) as ei:
        eng.ac_solve(table, np.zeros(0))
    assert ei.value.code == native.ERR_INVALID
    with pytest.raises(native.NativeError):
        eng.tran_solve(table, 1e-6, 0)
    # elements out of R,C,L,V,S,D order are rejected
    bad = native.ElemTable(1, [native.ELEM_V, native.ELEM_R], [1, 1], [0, 0], [0, 0], [0, 0], [0, 3], [0, 1, 0, 10.0])
    with pytest.raises(native.NativeError, match="grouped"):
        eng.ac_solve(bad, [1.0])
    # node id out of range
    bad = native.ElemTable(1, [native.ELEM_R], [2], [0], [0], [0], [0], [10.0])
    with pytest.raises(native.NativeError, match="node id"):
        eng.ac_solve(bad, [1.0])
    # largest supported system: Nvar = 1024 runs (global-scratch tier), 1025 is refused with UNSUPPORTED
    for n_nodes, ok in ((1023, True), (1024, False)):
        lines = ["* big", "v1 n1 0 ac 1"] + ["r%d n%d n%d 1k" % (k, k, k + 1) for k in range(1, n_nodes)] + \
                ["c%d n%d 0 1n" % (k, k + 1) for k in range(1, n_nodes)]
        ckb = parse_netlist("\n".join(lines) + "\n.ac lin 2 10 20\n")
        if ok:
            outb = sp.simulate_ac_batch(ckb, [10.0, 20.0], engine=eng, flags=native.FLAG_DENSE)
            xr, ier, st = co.ac_solve(ckb, [10.0, 20.0])
            assert eng.stats()["tier"] == native.TIER_CTA_GMEM and outb["status"].max() == 0
            assert rel_err(outb["x"][0], xr) <= AC_TOL
        else:
            with pytest.raises(native.NativeError) as ei:
                sp.simulate_ac_batch(ckb, [10.0], engine=eng)
            assert ei.value.code == native.ERR_UNSUPPORTED


def test_tran_mid_sized_circuit_cta_tiers(eng):
    """A 40-node RC ladder transient (Nvar = 41 > the thread tiers): CTA tier with the matrix in shared memory
    and in the global scratch, against the oracle."""
    import spicey_b200 as sp
    lines = ["* ladder tran", "V1 n1 0 PULSE(0 1 0 1u 1u 20u 50u)"] + \
            ["r%d n%d n%d 1k" % (k, k, k + 1) for k in range(1, 40)] + ["c%d n%d 0 1n" % (k, k + 1) for k in range(1, 40)]
    text = "\n".join(lines) + "\n.tran 1u 60u\n"
    ref = o.simulate(text)["tran"]
    for flags, tier in ((0, native.TIER_CTA_SMEM), (native.FLAG_FORCE_GMEM, native.TIER_CTA_GMEM)):
        got = sp.simulateTRAN(parse_netlist(text), flags=flags)
        assert eng.stats()["tier"] == tier
        for kind in ("nodeVoltages", "elementCurrents"):
            for k, b in ref[kind].items():
                a, b = np.asarray(got[kind][k]), np.asarray(b)
                assert np.max(np.abs(a - b)) <= TRAN_TOL * max(1e-30, float(np.max(np.abs(b)))), (kind, k)


# ---- dense register-tile tier (tier 9, tile_kernel.cuh) -----------------------------------

TILE = native.FLAG_DENSE | native.FLAG_TILE
TGEN = native.FLAG_TILE_GENERIC   # stamp from the element table (the per-instance path) instead of the sweep's constants


@pytest.mark.parametrize("flags", [TILE, TILE | SM, TILE | TGEN, TILE | TGEN | SM])
def test_ac_tile_tier_ladder64_slice(eng, flags):
    """cfg 2 topology through the dense register-tile kernel (2-D cyclic tiles of [A b] in registers, partial pivoting with
    physical row swaps as solveComplex.ts:15-53): every 997th frequency against the oracle, magnitude and phase."""
    import spicey_b200 as sp
    text = w.rc_ladder(64)
    freqs = np.array(sp.analysis.ac_frequencies(parse_netlist(text)))[::997]
    out, x, ie, st, _ = ac_case(eng, text, freqs, flags)
    assert eng.stats()["tier"] == native.TIER_TILE, eng.stats()
    assert out["status"].max() == 0 and st.max() == 0
    assert rel_err(out["x"], x) <= AC_TOL, rel_err(out["x"], x)
    assert rel_err(out["ielem"], ie) <= AC_TOL
    assert rel_err(np.abs(out["x"]), np.abs(x)) <= AC_TOL
    assert np.max(np.abs(np.angle(out["x"] * np.conj(x)))) <= AC_TOL


def test_ac_tile_tier_dense65_default_policy(eng):
    """A complete RC graph on 64 nodes (Nvar = 65, no structural zero, the V row forces a row swap at the first step):
    4,100 points with default flags must take the register-tile tier (the sparse program would execute the whole dense
    elimination); every 64th point against the oracle, per entry."""
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_dense(64))
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    assert freqs.shape[0] == 1000001
    freqs = np.ascontiguousarray(freqs[:: freqs.shape[0] // 4100][:4100])
    xr, ier, st = co.ac_solve(ck, freqs[::64], nthreads=8)
    iscale = np.max(np.abs(ier), axis=1, keepdims=True)   # a current is a difference of two node voltages: relative to the row's largest
    for flags in (0, SM, TGEN):
        out = sp.simulate_ac_batch(ck, freqs, engine=eng, flags=flags)
        assert eng.stats()["tier"] == native.TIER_TILE, (flags, eng.stats())
        assert out["status"].max() == 0 and st.max() == 0
        assert rel_err(out["x"][0][::64], xr) <= AC_TOL, rel_err(out["x"][0][::64], xr)
        assert np.max(np.abs(out["ielem"][0][::64] - ier) / iscale) <= AC_TOL
    # the old one-thread-per-row kernel on the same points: both are partial pivoting on the same numbers
    out2 = sp.simulate_ac_batch(ck, freqs[::64], engine=eng, flags=native.FLAG_DENSE | native.FLAG_NO_TILE)
    assert eng.stats()["tier"] == native.TIER_CTA_SMEM
    assert rel_err(out2["x"][0], xr) <= AC_TOL


@pytest.mark.parametrize("n_nodes,n_elem", [(1, 0), (2, 3), (5, 12), (14, 60), (30, 200), (31, 40), (62, 300), (90, 500)])
def test_ac_tile_tier_random_rlc_networks(eng, n_nodes, n_elem):
    """Random RLC networks (pivoting away from the diagonal, ties, inductor guards) through the register-tile tier for
    Nvar = 3 .. 92, point-major and series-major, statuses equal to the oracle's."""
    import spicey_b200 as sp
    rng = np.random.default_rng(n_nodes * 1000 + n_elem)
    text = random_rlc_netlist(rng, n_nodes, n_elem, n_v=min(2, n_nodes))
    freqs = sp.analysis.ac_frequencies(parse_netlist(text))
    for flags in (TILE, TILE | SM, TILE | TGEN):
        out, x, ie, st, _ = ac_case(eng, text, freqs, flags)
        assert eng.stats()["tier"] == native.TIER_TILE, eng.stats()
        assert np.array_equal(out["status"], st) and st.max() == 0
        scale = np.max(np.abs(x), axis=2, keepdims=True)
        assert np.max(np.abs(out["x"] - x) / scale) <= AC_TOL
        iscale = np.max(np.abs(ie), axis=2, keepdims=True)
        assert np.max(np.abs(out["ielem"] - ie) / iscale) <= AC_TOL


@pytest.mark.parametrize("shape", ["16,5", "8,9", "32,3", "4,17", "13,10,1", "7,7", "1,36", "32,1"])
def test_ac_tile_tier_forced_shapes(shape, monkeypatch):
    """Thread grids other than the chosen one (SPICEY_TILE_SHAPE=TR,TC[,CTAs per SM]): thread columns that do not fill a
    warp (mirror lanes), one thread row, one thread column, tiles that do not divide Nvar."""
    import spicey_b200 as sp
    monkeypatch.setenv("SPICEY_TILE_SHAPE", shape)
    e = native.Engine()
    try:
        for n_nodes, n_elem in ((30, 200), (9, 30)):
            rng = np.random.default_rng(n_nodes * 1000 + n_elem)
            text = random_rlc_netlist(rng, n_nodes, n_elem, n_v=2)
            ck = parse_netlist(text)
            freqs = sp.analysis.ac_frequencies(ck)
            x, ie, st = co.ac_solve(ck, freqs, nthreads=4)
            for flags in (TILE, TILE | TGEN):
                out = sp.simulate_ac_batch(ck, freqs, engine=e, flags=flags)
                assert e.stats()["tier"] == native.TIER_TILE, e.stats()
                assert out["status"].max() == 0 and st.max() == 0
                scale = np.max(np.abs(x), axis=1, keepdims=True)
                assert np.max(np.abs(out["x"][0] - x) / scale) <= AC_TOL, shape
                iscale = np.max(np.abs(ie), axis=1, keepdims=True)
                assert np.max(np.abs(out["ielem"][0] - ie) / iscale) <= AC_TOL, shape
    finally:
        e.close()


def test_ac_tile_tier_statuses_sweeps_and_current_sources(eng):
    """Error statuses (R <= 0, singular, Complex.div guard) isolated per point, per-instance value sweeps, no element
    currents requested, and an I element, all through the register-tile tier."""
    import spicey_b200 as sp
    n = 8
    r = np.full(n, 30.0)
    r[3] = -1.0
    out, x, ie, st, _ = ac_case(eng, w.README_RC, [1.0, 10.0], TILE, n_inst=n, overrides={"r1": r})
    assert eng.stats()["tier"] == native.TIER_TILE
    assert np.array_equal(out["status"], st) and (st[3] == native.ST_R_NONPOS).all()
    ok = [0, 1, 2, 4, 5, 6, 7]
    assert rel_err(out["x"][ok], x[ok]) <= AC_TOL and np.isnan(out["x"][3]).all()
    out, x, ie, st, _ = ac_case(eng, "* sing\nv1 a 0 ac 1\nv2 a 0 ac 1\nr1 a 0 1k\n.ac lin 2 1 2\n", [1.0, 2.0], TILE)
    assert (st == native.ST_SINGULAR).all() and np.array_equal(out["status"], st)
    out, x, ie, st, _ = ac_case(eng, "* cdiv\nv1 a 0 ac 1\nl1 a b 1e-10\nr1 b 0 1k\n.ac lin 2 1 2\n", [1.0, 1e6], TILE)
    assert st[0, 0] == native.ST_CDIV and st[0, 1] == 0 and np.array_equal(out["status"], st)
    assert rel_err(out["x"][0, 1], x[0, 1]) <= AC_TOL
    # Monte-Carlo axis: 37 instances x 11 frequencies with swept R / C / source phasor
    rng = np.random.default_rng(5)
    n = 37
    ov = {"r1": rng.uniform(10, 100, n), "c1": rng.uniform(1e-5, 1e-3, n), "v1.acmag": rng.uniform(0.5, 2, n),
          "v1.acphase": rng.uniform(-180, 180, n)}
    out, x, ie, st, _ = ac_case(eng, w.README_RC, np.logspace(0, 3, 11), TILE | SM, n_inst=n, overrides=ov)
    assert eng.stats()["tier"] == native.TIER_TILE and out["status"].max() == 0
    assert rel_err(out["x"], x) <= AC_TOL and rel_err(out["ielem"], ie) <= AC_TOL
    # current source (Norton) against the oracle
    ck = parse_netlist(NORTON, current_sources=True)
    ref = o.simulate_ac(parse_netlist(NORTON, current_sources=True))
    got = sp.simulate_ac_batch(ck, np.array(ref["freqs"]), engine=eng, flags=TILE)
    assert eng.stats()["tier"] == native.TIER_TILE and got["status"].max() == 0
    for j, nm in enumerate(got["node_names"]):
        assert rel_err(got["x"][0][:, j], np.array([complex(z) for z in ref["nodeVoltages"][nm]])) <= AC_TOL, nm
    for j, nm in enumerate(got["element_names"]):
        assert rel_err(got["ielem"][0][:, j], np.array([complex(z) for z in ref["elementCurrents"][nm]])) <= AC_TOL, nm


def test_tran_host_call_pipelines_instance_chunks(monkeypatch):
    """The host-buffer transient call splits a batch into instance chunks (two result buffers on the device, 2-D copies of
    chunk i overlapping the kernel of chunk i + 1): forced to 9 chunks of 32 instances, ragged last chunk, the results
    (waveforms, element currents, final state, iteration counts, statuses) are the single-launch ones bit for bit."""
    import spicey_b200 as sp
    n = 270
    ov = {k: v[:n] for k, v in w.rectifier_overrides(100000).items()}
    ck = parse_netlist(w.RECTIFIER)
    e1 = native.Engine()
    try:
        one = sp.simulate_tran_batch(ck, n_inst=n, overrides=ov, want_iters=True, engine=e1)
        l1 = e1.stats()["kernel_launches"]
    finally:
        e1.close()
    monkeypatch.setenv("SPICEY_TRAN_CHUNK_BYTES", str(32 * 3001 * 8 * (2 + 4) + 32 * 3001 * 4))
    e2 = native.Engine()
    try:
        many = sp.simulate_tran_batch(ck, n_inst=n, overrides=ov, want_iters=True, engine=e2)
        l2 = e2.stats()["kernel_launches"]
    finally:
        e2.close()
    assert l1 == 1 and l2 == 9, (l1, l2)
    for key in ("v", "ielem", "iters", "status", "state"):
        if key in one and one[key] is not None:
            assert np.array_equal(one[key], many[key]), key
    vr, ir, _, st, _ = co.tran_solve(ck, one["dt"], one["steps"], n_inst=n, overrides=ov, nthreads=8)
    assert st.max() == 0 and many["status"].max() == 0
    assert np.max(np.abs(many["v"] - np.transpose(vr, (1, 2, 0)))) <= TRAN_TOL * np.max(np.abs(vr))


def test_ac_program_tiers_hand_over_when_the_pilot_order_does_not_hold():
    """A random RLC network swept over five decades: the pivot order of the pilot point holds for almost no other point, so
    nearly every system of the first launch comes back on the fallback list — solved there by the register kernels (the
    one-warp-per-system LU in list mode), exact statuses and results — and the handle then sends later launches of this
    topology to the dense tier directly (one check per cached program)."""
    import spicey_b200 as sp
    rng = np.random.default_rng(14060)
    text = random_rlc_netlist(rng, 14, 60, n_v=2).replace(".ac dec 7 10 1meg", ".ac dec 1640 10 1meg")
    ck = parse_netlist(text)
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    assert freqs.shape[0] >= 8192
    pick = np.arange(0, freqs.shape[0], 61)
    xr, ier, st = co.ac_solve(ck, freqs[pick], nthreads=8)
    scale, iscale = np.max(np.abs(xr), axis=1, keepdims=True), np.max(np.abs(ier), axis=1, keepdims=True)
    e = native.Engine()
    try:
        for call, flags in enumerate((SM, SM, 0)):
            out = sp.simulate_ac_batch(ck, freqs, engine=e, flags=flags)
            stt = e.stats()
            if call == 0:
                assert stt["tier"] in (native.TIER_SPARSE, native.TIER_SPARSE_JIT), stt
                assert stt["fallback_solves"] * 4 >= freqs.shape[0], stt
            else:
                assert stt["tier"] == native.TIER_TILE and stt["fallback_solves"] == 0, (call, stt)
            assert out["status"].max() == 0 and st.max() == 0
            assert np.max(np.abs(out["x"][0][pick] - xr) / scale) <= AC_TOL, call
            assert np.max(np.abs(out["ielem"][0][pick] - ier) / iscale) <= AC_TOL, call
        # a topology whose pilot order does hold keeps its program tier on the same handle (cfg 2's ladder)
        ckl = parse_netlist(w.rc_ladder(64))
        fl = np.array(sp.analysis.ac_frequencies(ckl))[::97]
        for _ in range(3):
            sp.simulate_ac_batch(ckl, fl, engine=e, flags=SM)
            assert e.stats()["tier"] == native.TIER_SPARSE and e.stats()["fallback_solves"] == 0, e.stats()
    finally:
        e.close()


def _rc_system_matrices(table):
    """A(w) = A0 + j w A1 and b of an R / C / V circuit, stamped as simulateAC.ts:24-60 does (rows of node ids n - 1,
    V branch rows nn + k, b[branch] = the source phasor) — independent of the oracle and of the library's own plan."""
    n, nn = table.nvar, table.n_nodes
    A0, A1, b = np.zeros((n, n)), np.zeros((n, n)), np.zeros(n, dtype=np.complex128)
    kv = 0
    for e in range(table.n_ac_elem):
        ty, n1, n2, v = int(table.type[e]), int(table.n1[e]) - 1, int(table.n2[e]) - 1, table.values[table.value_idx[e]:]
        if ty == native.ELEM_V:
            j = nn + kv
            kv += 1
            if n1 >= 0:
                A0[n1, j] += 1; A0[j, n1] += 1
            if n2 >= 0:
                A0[n2, j] -= 1; A0[j, n2] -= 1
            b[j] += v[1] * np.exp(1j * np.pi * v[2] / 180)
            continue
        M, y = (A0, 1 / v[0]) if ty == native.ELEM_R else (A1, v[0])
        assert ty in (native.ELEM_R, native.ELEM_C)
        if n1 >= 0:
            M[n1, n1] += y
        if n2 >= 0:
            M[n2, n2] += y
        if n1 >= 0 and n2 >= 0:
            M[n1, n2] -= y; M[n2, n1] -= y
    return A0, A1, b


def test_full_size_properties_dense64(eng):
    """The dense workload (complete RC graph, Nvar 65) at the bench's size, 200,000 points through the register-tile tier:
    every status 0, and EVERY point's solution satisfies its own system — the residual |A(w) x - b| against
    |A| |x| + |b|, row by row, with A(w) built here from the element table in batches on the GPU — plus KCL at the source
    node and a 1/997 subsample against the oracle."""
    import torch
    import spicey_b200 as sp
    ck = parse_netlist(w.rc_dense(64))
    table = sp.packing.pack_circuit(ck)
    freqs = np.array(sp.analysis.ac_frequencies(ck))
    freqs = np.ascontiguousarray(freqs[:: freqs.shape[0] // 200000][:200000])
    out = sp.simulate_ac_batch(ck, freqs, engine=eng, want_currents=False)
    assert eng.stats()["tier"] == native.TIER_TILE and out["status"].max() == 0
    x = out["x"][0]
    A0, A1, b = _rc_system_matrices(table)
    dev = torch.device("cuda", 0)
    tA0, tA1 = torch.from_numpy(A0).to(dev).to(torch.complex128), torch.from_numpy(A1).to(dev).to(torch.complex128)
    tb = torch.from_numpy(b).to(dev)
    worst = 0.0
    for c0 in range(0, freqs.shape[0], 20000):
        wv = torch.from_numpy(2 * np.pi * freqs[c0:c0 + 20000]).to(dev).to(torch.complex128)
        A = tA0[None] + 1j * wv[:, None, None] * tA1[None]
        tx = torch.from_numpy(np.ascontiguousarray(x[c0:c0 + 20000])).to(dev)
        res = torch.abs(torch.einsum("pij,pj->pi", A, tx) - tb[None])
        den = torch.einsum("pij,pj->pi", torch.abs(A), torch.abs(tx)) + torch.abs(tb)[None]
        worst = max(worst, float(torch.max(res / den)))
        del A, res, den
    assert worst <= 1e-13, worst
    assert np.max(np.abs(x[:, 0] - 1.0)) <= 1e-15            # the driven node carries the source phasor
    sub = slice(0, None, 997)
    xr, _, st = co.ac_solve(ck, freqs[sub], nthreads=8)
    assert st.max() == 0 and rel_err(x[sub], xr) <= AC_TOL


def test_tran_probes_select_what_crosses_the_bus(eng, golden, monkeypatch):
    """spicey_tran_solve_probes: simulateTRAN keeps the `.PRINT TRAN` node voltages only (simulateTRAN.ts:240-249), so only
    those rows are copied to the host — same keys and bit-identical waveforms as the full call, fewer bytes; any order and
    repetition of node ids, batches in several pipeline chunks, and an empty selection."""
    import spicey_b200 as sp
    for name in ("switch_vt_vh", "boost_converter_probe", "two_probes"):
        g = golden(name)
        ck = parse_netlist(g["netlist"])
        assert len(ck.probes.tran) > 0
        got = sp.simulateTRAN(ck, engine=eng)
        d2h_sel = eng.stats()["d2h_bytes"]
        ref = o.simulate_tran(parse_netlist(g["netlist"]))
        assert list(got["nodeVoltages"].keys()) == list(ref["nodeVoltages"].keys())
        ck2 = parse_netlist(g["netlist"])
        ck2.probes.tran.clear()
        full = sp.simulateTRAN(ck2, engine=eng)
        if name == "two_probes":   # (probes both of its nodes)
            assert eng.stats()["d2h_bytes"] == d2h_sel and len(full["nodeVoltages"]) == len(got["nodeVoltages"]) == 2
        else:
            assert eng.stats()["d2h_bytes"] > d2h_sel and len(full["nodeVoltages"]) > len(got["nodeVoltages"])
        for k, series in got["nodeVoltages"].items():
            assert np.array_equal(series, full["nodeVoltages"][k]), (name, k)
            want = np.asarray(ref["nodeVoltages"][k])
            assert np.max(np.abs(series - want)) <= TRAN_TOL * max(1.0, np.max(np.abs(want))), (name, k)
        for k, series in got["elementCurrents"].items():
            assert np.array_equal(series, full["elementCurrents"][k]), (name, k)
    # batch, ragged chunks, repeated / reordered ids
    n = 150
    ov = {k: v[:n] for k, v in w.rectifier_overrides(100000).items()}
    ck = parse_netlist(w.RECTIFIER)
    table = sp.packing.pack_circuit(ck)
    dt, steps = compute_effective_time_step(ck.analyses.tran.dt, ck.analyses.tran.tstop)
    vsrc, mask = sp.packing.sample_sources(ck, dt, steps)
    sweep = sp.packing.make_sweep(table, n, ov)
    st0 = sp.packing.initial_state(ck, table, n)
    full = eng.tran_solve(table, dt, steps, vsrc=vsrc, vsrc_mask=mask, sweep=sweep, state0=st0)
    monkeypatch.setenv("SPICEY_TRAN_CHUNK_BYTES", str(64 * 3001 * 8 * 6))
    e2 = native.Engine()
    try:
        for sel in ([2], [2, 1, 2], []):
            part = e2.tran_solve(table, dt, steps, vsrc=vsrc, vsrc_mask=mask, sweep=sweep, state0=st0, node_sel=sel)
            assert part["v"].shape == (steps + 1, len(sel), n)
            for k, node in enumerate(sel):
                assert np.array_equal(part["v"][:, k, :], full["v"][:, node - 1, :]), (sel, k)
            assert np.array_equal(part["ielem"], full["ielem"]) and np.array_equal(part["state"], full["state"])
            assert np.array_equal(part["status"], full["status"])
        with pytest.raises(native.NativeError):
            e2.tran_solve(table, dt, steps, vsrc=vsrc, vsrc_mask=mask, sweep=sweep, state0=st0, node_sel=[3])
    finally:
        e2.close()
