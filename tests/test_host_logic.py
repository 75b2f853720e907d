"""CPU-only checks of the host side: parser mirror, packing, formatting, the C-ABI library
(loads and exports every declared symbol; no compute without a GPU), sharding over gloo."""
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from spicey_b200 import native, workloads as w
from spicey_b200.analysis import _js_key_order, _series_by_name
from spicey_b200.formatting import to_precision
from spicey_b200.packing import initial_state, make_sweep, pack_circuit, sample_sources
from spicey_b200.parsing import (build_frequency_array, compute_effective_time_step, parse_netlist,
                                 parse_number_with_units, pulse_value, pwl_value)
from spicey_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parse_number_with_units_reference_arithmetic():
    """Hazard H1: parseFloat(num) * multiplier, not the decimal literal."""
    assert parse_number_with_units("3m") == 3 * 1e-3
    assert parse_number_with_units("20u") == 20 * 1e-6 and parse_number_with_units("20u") != 2e-5
    assert parse_number_with_units("1meg") == 1e6 and parse_number_with_units("2.2k") == 2.2 * 1e3
    assert parse_number_with_units("100uF") == 100 * 1e-6 and parse_number_with_units("10kohm") == 10e3
    assert parse_number_with_units("1e-3") == 1e-3 and parse_number_with_units("5") == 5.0
    assert math.isnan(parse_number_with_units("abc")) and math.isnan(parse_number_with_units(""))
    assert parse_number_with_units("7x") == 7.0  # unknown suffix: bare value (parseNumberWithUnits.ts:29)


def test_step_count_hazard_h2():
    ck = parse_netlist(w.RLC_TANK)
    dt, steps = compute_effective_time_step(ck.analyses.tran.dt, ck.analyses.tran.tstop)
    assert steps == 1001 and dt == ck.analyses.tran.tstop / 1001
    ck = parse_netlist(w.RECTIFIER)
    assert compute_effective_time_step(ck.analyses.tran.dt, ck.analyses.tran.tstop)[1] == 3000
    assert compute_effective_time_step(0, 10 * 1e-3) == ((10 * 1e-3) / 1000, 1000)


def test_parser_mirrors_reference_tests(golden):
    ck = parse_netlist(golden("case_insensitive_nodes")["netlist"])
    assert ck.nodes.count() == 3 and ck.nodes.rev == ["0", "nOdE1", "nOde2"]
    assert sorted(ck.probes.tran) == sorted(["NODE2", "node1"])
    ck = parse_netlist(golden("switch_vt_vh")["netlist"])
    m = ck.S[0].model
    assert abs(m.Von - 2.55) < 1e-12 and abs(m.Voff - 2.45) < 1e-12 and m.Ron == 0.1 and m.Roff == 1e9
    assert ck.probes.tran == ["n2", "nctrl_sw1"]
    ck = parse_netlist(golden("diode_switch")["netlist"])
    assert len(ck.D) == 1 and len(ck.S) == 1 and ck.models.diode["d"].Is == 1e-14 and ck.models.vswitch["swmod"].Ron == 1
    assert [v.index for v in ck.V] == [4, 5]
    ck = parse_netlist(golden("vswitch_pwl")["netlist"])
    m = ck.S[0].model
    assert (m.Ron, m.Roff, m.Von, m.Voff) == (1, 1e9, 2, 1)
    assert ck.V[1].pwl[1] == (1 * 1e-3, 5.0)
    # README quirk: the title "Demo of ..." starts with 'd' -> skipped as a malformed diode
    ck = parse_netlist(w.README_RC)
    assert any(s.startswith("Demo") for s in ck.skipped) and len(ck.R) == 1 and ck.analyses.ac.N == 100
    with pytest.raises(ValueError, match="Unknown .model"):
        parse_netlist("* x\ns1 a 0 c 0 nomodel\n")
    with pytest.raises(ValueError, match="Resistor missing node"):
        parse_netlist("RC ladder\n")


def test_waveforms_and_frequency_lists():
    ck = parse_netlist("* p\nV1 1 0 PULSE(0 5 1u 1u 1u 2u 10u)\nR1 1 0 1\n")
    f = ck.V[0].waveform
    assert f(0) == 0 and f(1.5e-6) == pytest.approx(2.5) and f(3e-6) == 5 and f(9e-6) == 0 and f(12.5e-6) == 5
    assert pwl_value([(0, 0), (1, 10)], 0.25) == 2.5 and pwl_value([], 3) == 0 and pwl_value([(0, 1), (1, 2)], 5) == 2
    fr = build_frequency_array("dec", 100, 1.0, 100.0)
    assert len(fr) == 201 and fr[0] == 1.0 and fr[-1] == pytest.approx(100.0)
    assert build_frequency_array("lin", 5, 1.0, 3.0) == [1.0, 1.5, 2.0, 2.5, 3.0]
    assert len(build_frequency_array("dec", 200000, 1.0, 100 * 1e3)) == 1000001
    with pytest.raises(ValueError, match="frequencies must be > 0"):
        build_frequency_array("dec", 10, 0.0, 10.0)


def test_to_precision_matches_ecmascript():
    cases = {1.0: "1.00000", 0.999822: "0.999822", -1.07987: "-1.07987", 123456.7: "123457", 1234567.0: "1.23457e+6",
             0.000001234: "0.00000123400", 0.0000001234: "1.23400e-7", 0.0: "0.00000", 99999.96: "100000",
             999999.5: "1.00000e+6", 5e-324: "4.94066e-324", 1e21: "1.00000e+21"}
    for x, s in cases.items():
        assert to_precision(x, 6) == s, (x, to_precision(x, 6), s)


def test_js_key_order_and_duplicate_names():
    assert _js_key_order(["b", "2", "a", "1", "01"]) == ["1", "2", "b", "a", "01"]
    cols = np.arange(12.0).reshape(4, 3)
    s = _series_by_name(["r1", "x", "r1"], cols)
    assert list(s) == ["r1", "x"] and s["r1"].tolist() == [0, 2, 3, 5, 6, 8, 9, 11]  # duplicates interleave per sample


def test_pack_circuit_layout(golden):
    ck = parse_netlist(golden("boost_converter_probe")["netlist"])
    t = pack_circuit(ck)
    assert t.names == ["RR1", "CC1", "LL1", "Vsimulation_voltage_source_0", "Vsimulation_voltage_source_1", "SM1", "DD1"]
    assert t.type.tolist() == [0, 1, 2, 3, 3, 4, 5] and t.nvar == 6 and t.n_state == 4 and t.n_ac_elem == 5
    assert t.values[t.value_idx[5]:t.value_idx[5] + 4].tolist() == [1, 1e12, 0, 0]
    assert t.nc1[5] == ck.nodes.get("N4") and t.nc2[5] == 0
    sw = make_sweep(t, 3, {"rr1": [1, 2, 3], "dd1.is": [1e-14, 2e-14, 3e-14], "SM1.ron": [1, 1, 2]})
    assert sw.var_slot.tolist() == [t.value_idx[0], t.value_idx[6], t.value_idx[5]]
    assert sw.var_values.shape == (3, 3)
    assert initial_state(ck, t, 2).shape == (4, 2)
    tab, mask = sample_sources(ck, 1e-3, 4)
    assert mask.tolist() == [0, 1] and tab.shape == (2, 5)
    with pytest.raises(KeyError):
        make_sweep(t, 3, {"nope": [1, 2, 3]})


def test_library_loads_and_exports_every_declared_symbol():
    lib = native.load_library()
    hdr = open(os.path.join(ROOT, "include", "spicey_native.h")).read()
    declared = set(re.findall(r"\b(spicey_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(native.EXPORTS), declared ^ set(native.EXPORTS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.spicey_native_abi_version() == 4


def test_no_cpu_fallback_without_device():
    lib = native.load_library()
    if lib.spicey_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(native.NativeError) as ei:
        native.Engine()
    assert ei.value.code == native.ERR_NO_DEVICE
    import spicey_b200 as sp
    with pytest.raises(native.NativeError):
        sp.simulate(w.README_RC)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spicey_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f


def test_shard_range_partitions():
    for n in (0, 1, 7, 1000001):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[i][1] == r[i + 1][0] for i in range(world - 1))


def test_sharded_sweep_world_size_2_gloo():
    """N>1 path on CPU: two gloo ranks each solve their contiguous frequency slice (with the oracle as the
    stand-in solver, this being a CPU test) and rank 0 gathers; result equals the single-process sweep."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", PYTHONPATH=ROOT)
    procs = [subprocess.Popen([sys.executable, script, str(r), "2"], env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GATHER_OK" in outs[0]


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the reference algorithm's C port on the host cores) prints exactly one JSON line
    with the contract's keys; the native arm refuses to run without a CUDA device instead of falling back."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    import torch
    if not torch.cuda.is_available():
        r2 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "cfg1", "--steps", "1",
                             "--warmup", "1"], capture_output=True, text=True, timeout=300)
        assert r2.returncode != 0 and "no CPU path" in (r2.stderr + r2.stdout)


def test_current_source_extension_parses_and_packs():
    """`I` lines: skipped like the reference does by default, parsed into ckt.I on request, packed as the last group."""
    from spicey_b200 import native
    from spicey_b200.packing import pack_circuit
    from spicey_b200.parsing import parse_netlist
    text = "* i\nv1 a 0 dc 1\ni1 0 b dc 2m ac 3m 45\nr1 a b 1k\nc1 b 0 1u\n.ac dec 2 1 10\n.end\n"
    ref_like = parse_netlist(text)
    assert ref_like.I == [] and ref_like.skipped == ["i1 0 b dc 2m ac 3m 45"]
    ck = parse_netlist(text, current_sources=True)
    assert len(ck.I) == 1 and (ck.I[0].n1, ck.I[0].n2) == (0, ck.nodes.get_or_create("b"))
    tb = pack_circuit(ck)
    assert list(tb.type) == [native.ELEM_R, native.ELEM_C, native.ELEM_V, native.ELEM_I]
    assert tb.n_ac_elem == 3 and tb.n_elem == 4 and tb.nvar == 3 and tb.names[-1] == "i1"
    vi = int(tb.value_idx[-1])
    assert list(tb.values[vi:vi + 3]) == [2 * 1e-3, 3 * 1e-3, 45.0]


def test_lazy_current_series_reproduces_the_reference_formula():
    """LazyCurrentSeries (simulateAC(lazy_currents=True)) on the oracle's own node voltages: Y.mul(v1.sub(v2)) per
    element kind, including the inductor's Complex.div form and ground on either side (simulateAC.ts:94-126)."""
    from oracle import spicey_oracle as o
    from spicey_b200 import native
    from spicey_b200.analysis import LazyCurrentSeries
    from spicey_b200.packing import pack_circuit
    from spicey_b200.parsing import parse_netlist
    text = "* rlc\nv1 in 0 ac 1 15\nr1 in a 50\nl1 a b 1m\nc1 b 0 1u\nr2 0 b 2k\nc2 a b 10n\n.ac dec 7 10 1meg\n.end\n"
    ref = o.simulate(text)["ac"]
    ck = parse_netlist(text)
    tb = pack_circuit(ck)
    f = np.array(ref["freqs"])
    volt = {nm: np.array([complex(z) for z in s]) for nm, s in ref["nodeVoltages"].items()}
    names = ck.nodes.rev
    for e in range(tb.n_ac_elem):
        kind, n1, n2 = int(tb.type[e]), int(tb.n1[e]), int(tb.n2[e])
        want = np.array([complex(z) for z in ref["elementCurrents"][tb.names[e]]])
        if kind == native.ELEM_V:
            continue   # the branch unknown itself
        s = LazyCurrentSeries(kind, float(tb.values[int(tb.value_idx[e])]), f, None if n1 == 0 else volt[names[n1]],
                              None if n2 == 0 else volt[names[n2]])
        assert np.array_equal(s.array, want), tb.names[e]            # same operations in the same order: bit-identical
        assert complex(s[2]) == want[2] and len(s) == len(want)
