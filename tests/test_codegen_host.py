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
