"""Host side of the dense register-tile tier (tier 9): the thread-grid choice and the kernel text the library hands to
NVRTC, compiled here with nvcc for sm_100a (no GPU needed) within the register budget the launch bound implies."""
import os
import re
import shutil
import subprocess

import pytest

from spicey_b200 import native


def test_tile_shape_choice():
    src, sh = native.tile_kernel_source(65, 127, 1)
    # cfg 2's size: 13 x 11 threads (6 warps, two thread columns each), 5 x 6 tiles = 65 x 66 exactly, two systems per SM
    assert (sh["tr"], sh["tc"], sh["mr"], sh["mc"], sh["warps"], sh["ctas_per_sm"]) == (13, 11, 5, 6, 6, 2)
    assert sh["tr"] * sh["mr"] >= 65 and sh["tc"] * sh["mc"] >= 66
    assert "#define TL_N 65\n#define TL_TR 13\n#define TL_TC 11\n#define TL_WARPS 6\n#define TL_MINB 2\n#define TL_IELEM 1\n" in src
    assert "spicey_tile_jit" in src and "__reduce_max_sync" in src
    for n in (1, 2, 3, 8, 16, 31, 32, 33, 48, 64, 80, 96):
        got = native.tile_kernel_source(n, 4 * n, 2)
        assert got is not None, n
        s = got[1]
        assert s["tr"] * s["mr"] >= n and s["tc"] * s["mc"] >= n + 1 and s["mr"] * s["mc"] <= 36
        assert s["tr"] * (32 // s["tr"]) <= 32 and s["warps"] == -(-s["tc"] // (32 // s["tr"]))
        assert s["smem_bytes"] * s["ctas_per_sm"] <= 227 * 1024
        if n <= 32:
            assert s["warps"] <= 2, (n, s)   # small systems: one or two warps per system, many systems per SM
    assert native.tile_kernel_source(257, 700, 1) is None   # cfg 4 does not fit a register file: other tiers
    forced = native.tile_kernel_source(32, 100, 2, tr=7, tc=7)[1]
    assert (forced["tr"], forced["tc"], forced["mr"], forced["mc"], forced["warps"]) == (7, 7, 5, 5, 2)


@pytest.mark.parametrize("nvar,tr,tc,const_tables,rc_only", [(65, 0, 0, True, True), (65, 0, 0, True, False), (65, 0, 0, False, False),
                                                             (16, 0, 0, True, True), (32, 7, 7, False, False)])
def test_tile_kernel_compiles_within_its_register_budget(tmp_path, nvar, tr, tc, const_tables, rc_only):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    src, sh = native.tile_kernel_source(nvar, 200, 2, tr=tr, tc=tc, const_tables=const_tables, rc_only=rc_only)
    assert "#define TL_CONST %d\n#define TL_RC %d\n" % (const_tables, rc_only) in src
    cu = tmp_path / "tile.cu"
    cu.write_text(src)
    res = subprocess.run([nvcc, "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xptxas", "-v",
                          "-o", str(tmp_path / "tile.cubin"), str(cu)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    m = re.search(r"Used (\d+) registers", res.stderr)
    assert m, res.stderr
    spills = [int(v) for v in re.findall(r"(\d+) bytes spill stores", res.stderr)]
    # (the warp that runs the look-ahead pivot search holds the search's temporaries on top of the tile: a few spilled
    #  values at Nvar = 65, where two systems per SM leave 168 registers per thread)
    assert max(spills) <= 512, res.stderr
    threads = sh["warps"] * 32
    assert int(m.group(1)) * threads * sh["ctas_per_sm"] <= 65536, (m.group(1), sh)


@pytest.mark.parametrize("nvar,rc_only,const_tables", [(32, True, True), (17, False, True), (3, True, True), (32, False, False), (9, False, False)])
def test_warp_lu_kernel_compiles_within_its_register_budget(tmp_path, nvar, rc_only, const_tables):
    """The one-warp-per-system dense LU (Nvar <= 32): a lane's row in registers at the occupancy the host asks for."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    assert native.warp_lu_kernel_source(33) is None
    src, sh = native.warp_lu_kernel_source(nvar, rc_only=rc_only, const_tables=const_tables)
    assert "#define WL_CONST %d\n" % const_tables in src
    assert "#define WL_N %d\n" % nvar in src and "spicey_warp_lu_jit" in src and "__reduce_max_sync" in src
    cu = tmp_path / "wlu.cu"
    cu.write_text(src)
    res = subprocess.run([nvcc, "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xptxas", "-v",
                          "-o", str(tmp_path / "wlu.cubin"), str(cu)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    m = re.search(r"Used (\d+) registers", res.stderr)
    spills = [int(v) for v in re.findall(r"(\d+) bytes spill stores", res.stderr)]
    assert m and max(spills) <= 128, res.stderr
    assert int(m.group(1)) * sh["warps"] * 32 * sh["ctas_per_sm"] <= 65536, (m.group(1), sh)
