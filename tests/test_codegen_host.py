"""CPU checks of the two kernel generators (no GPU): the CUDA source they emit for the BASELINE netlists and for
the reference's own transient netlists must compile for sm_100a with nvcc (the GPU box compiles the same text
with NVRTC), without local-memory spills where the design promises none."""
import os
import re
import shutil
import subprocess

import pytest

from spicey_b200 import native, packing, parsing, workloads

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")


def _compile(src, tmp_path, name):
    cu = tmp_path / (name + ".cu")
    cu.write_text(src)
    r = subprocess.run([NVCC, "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xptxas", "-v",
                        "-o", str(tmp_path / (name + ".cubin")), str(cu)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    regs = int(re.search(r"Used (\d+) registers", r.stderr).group(1))
    spill = int(re.search(r"(\d+) bytes spill stores", r.stderr).group(1))
    return regs, spill


def test_sparse_ac_kernel_source_cfg2_compiles_without_spills(tmp_path):
    table = packing.pack_circuit(parsing.parse_netlist(workloads.rc_ladder(64)))
    src, st = native.sparse_kernel_source(table, 1000.0, block=192, min_blocks=1, smem_slots=75, sync=4)
    # 65 reciprocals + 64 eliminated right-hand sides cross into the back-substitution; 75 of them in shared memory
    assert st["saved_values"] == 129 and st["smem_slots"] == 75 and st["cfma"] == 313 and st["reciprocals"] == 65
    assert "rcp_nr(" in src and "__constant__ double KC[" in src and "st.shared.v2.f64" in src
    # result stores are unconditional (lanes past the end re-solve and re-write the last point): no predicate per row
    assert "if (valid) *(double2*)" not in src and "const long long pc = min(p, plast);" in src
    assert src.count("*(double2*)(xb + ") == 65 and src.count("*(double2*)(ib + ") == 127
    regs, spill = _compile(src, tmp_path, "ac_cfg2")
    assert regs <= 255 and spill == 0
    # without element currents (ielem == NULL): a second variant of the same program
    src2, _ = native.sparse_kernel_source(table, 1000.0, block=192, min_blocks=1, smem_slots=75, sync=4, with_ielem=False)
    assert "a.ielem" not in src2.split("spicey_sparse_jit")[1]
    _compile(src2, tmp_path, "ac_cfg2_noi")


def test_sparse_ac_kernel_source_small_and_rlc(tmp_path):
    for name, text in (("readme", workloads.README_RC),
                       ("rlc", "* rlc\nv1 a 0 ac 1 30\nr1 a b 10\nl1 b c 1m\nc1 c 0 1u\nr2 c 0 1k\n.ac dec 5 10 1meg\n")):
        table = packing.pack_circuit(parsing.parse_netlist(text))
        src, st = native.sparse_kernel_source(table, 1000.0)
        if name == "rlc":
            assert "iw" in src   # 1/(w*L) terms
        _compile(src, tmp_path, "ac_" + name)


@pytest.mark.parametrize("name", ["cfg3", "cfg5", "transient01_rc_pulse", "switch_vt_vh", "boost_converter_probe",
                                  "diode_switch"])
def test_transient_kernel_source_compiles(tmp_path, golden, name):
    if name == "cfg3":
        text, ov = workloads.RLC_TANK, workloads.rlc_tank_overrides(8)
    elif name == "cfg5":
        text, ov = workloads.RECTIFIER, workloads.rectifier_overrides(8)
    else:
        text, ov = golden(name)["netlist"], None
    table = packing.pack_circuit(parsing.parse_netlist(text))
    sweep = packing.make_sweep(table, 8, ov) if ov else None
    src = native.tran_kernel_source(table, sweep)
    assert "spicey_tran_jit" in src and "SmallLU<NV>" in src
    if name == "cfg3":
        assert src.count("lu.factor()") == 1 and "for (; it < 20" not in src      # constant matrix: factored once
    if name == "cfg5":
        assert "elo3" in src and src.count("exp(") == 4                            # one exp per step + 3 in the prologue
    if name in ("switch_vt_vh", "boost_converter_probe"):
        assert "for (; it < 20; ++it)" in src                                      # re-solve while a switch toggled
    regs, spill = _compile(src, tmp_path, "tran_" + name)
    assert spill == 0


WAVE_NETLIST = ("* device-evaluated sources\nV1 1 0 PULSE(0 5 1u 1u 1u 3u 10u)\nV2 3 0 PWL(0 0 1u 1 2u 0.5 4u 2)\n"
                "R1 1 2 1k\nR2 3 2 2k\nC1 2 0 1u\n.tran 0.1u 20u\n")


def test_transient_kernel_source_with_device_waveforms_compiles(tmp_path):
    """PULSE / PWL parameters are value slots; the generated kernel evaluates pulseValue / pwlValue
    itself with separately rounded operations, swept parameters are per-instance loads."""
    ck = parsing.parse_netlist(WAVE_NETLIST)
    table = packing.pack_circuit(ck, device_waves=True)
    assert list(table.waves.kind) == [native.WAVE_PULSE, native.WAVE_PWL] and list(table.waves.n_pairs) == [0, 4]
    p0 = int(table.waves.value_idx[0])
    assert list(table.values[p0:p0 + 8]) == [0.0, 5.0, 1e-6, 1e-6, 1e-6, 3 * 1e-6, 10 * 1e-6, float("inf")]
    assert table.wave_params["v2.pwl.t1"] == int(table.waves.value_idx[1]) + 2
    sweep = packing.make_sweep(table, 4, {"V1.pulse.v2": [1.0, 2.0, 3.0, 4.0], "v2.pwl.t1": [1e-6, 1.1e-6, 1.2e-6, 1.3e-6]})
    assert list(sweep.var_slot) == [p0 + 1, table.wave_params["v2.pwl.t1"]]
    src = native.tran_kernel_source(table, sweep, waves=table.waves)
    body = src.split("spicey_tran_jit")[-1]
    assert "pulse_value(pw3v1" in body and body.count("pwl_segment(") == 3 and "__dmul_rn((double)step, a.dt)" in body
    assert "pw3v2 = a.var_values[0ll" in body and "pl4t1 = a.var_values[1ll" in body
    assert "__longlong_as_double(0x7ff0000000000000ll)" in body        # ncycles = Infinity has no literal form
    assert "a.vsrc" not in body.split("for (; step")[1]                 # no pre-sampled row is read
    regs, spill = _compile(src, tmp_path, "tran_waves")
    assert spill == 0
    # without descriptors the same table compiles to the row-reading kernel
    assert "__ldg(a.vsrc" in native.tran_kernel_source(table, sweep)
    # descriptor validation happens on the host
    bad = native.Waves([native.WAVE_PULSE, native.WAVE_DC], [len(table.values) - 3, 0], [0, 0])
    with pytest.raises(native.NativeError, match="waveform parameter slots out of range"):
        native.tran_kernel_source(table, sweep, waves=bad)
    with pytest.raises(KeyError):
        packing.make_sweep(packing.pack_circuit(ck), 2, {"v1.pulse.v2": [1.0, 2.0]})


def test_warp_program_sizes_cfg4():
    table = packing.pack_circuit(parsing.parse_netlist(workloads.rc_mesh(16)))
    st = native.warp_program_stats(table)
    assert st["nvar"] == 257 and st["max_rows_per_step"] == 17
    assert st["pool_slots"] * 16 < 12 * 1024 and st["thread_tier_slots"] > 4000
    assert st["global_slots"] > st["backsub_entries"]      # U entries + reciprocals + right-hand sides


def test_series_ld_is_a_multiple_of_32_points():
    lib = native.load_library()
    for p in (1, 31, 32, 33, 1000001, 8000001):
        ld = lib.spicey_series_ld(p)
        assert ld >= p and ld % 32 == 0 and ld - p < 32


def test_sparse_ac_kernel_source_with_global_column_compiles_without_spills(tmp_path):
    """A 150-node ladder hands 301 values from the elimination to the back-substitution: 75 in shared memory, 40 in
    registers, the 186 longest-lived in the kernel's [slot][thread] column of global memory — stored where they are
    produced, loaded a few rows ahead of their use, and next to no local-memory spills (the two opaque copies of the column's
    base keep the compiler from carrying 186 addresses across the phases)."""
    import re
    import shutil
    import subprocess
    from spicey_b200 import native, packing, parsing, workloads
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    table = packing.pack_circuit(parsing.parse_netlist(workloads.rc_ladder(150)))
    src, st = native.sparse_kernel_source(table, 1000.0, 96, 2, 75, True, 4, reg_values=40)
    assert st["saved_values"] == 301 and st["smem_slots"] == 75 and st["gmem_slots"] == 186, st
    assert src.count("    GST(gwp + ") == 186 and src.count(", gwq + ") == 186 and "double2* work;" in src
    assert "createpolicy.fractional.L2::evict_last" in src and src.count("    RST(") == 151 + 299   # 151 unknowns + 299 element currents stream past the column
    # cfg 2 fits without the column and its source does not mention it
    src2, st2 = native.sparse_kernel_source(packing.pack_circuit(parsing.parse_netlist(workloads.rc_ladder())), 1000.0, 96, 2, 75, True, 4,
                                            reg_values=64)
    assert st2["gmem_slots"] == 0 and "gwp" not in src2
    cu = tmp_path / "k.cu"
    cu.write_text(src)
    res = subprocess.run([nvcc, "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xptxas", "-v",
                          "-o", str(tmp_path / "k.cubin"), str(cu)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    # (a handful of spilled values at most; before the opaque bases the same kernel spilled 1.7 KB per thread)
    assert max(int(v) for v in re.findall(r"(\d+) bytes spill stores", res.stderr)) <= 128, res.stderr
