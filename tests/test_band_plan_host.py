"""CPU check of the banded + bordered AC tier's host plan (spicey_b200/csrc/band_plan.h: ordering, pilot, per-step
stamp deliveries, tie-rule flags) together with the schedule the kernel executes (band_kernel.cuh), emulated lane by
lane in tests/cpp/band_plan_check.cpp and compared with dense partial-pivoting elimination in the reference's
unknown order (solveComplex.ts:4-73), both judged against the same elimination in extended precision."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "band_plan_check")


@pytest.fixture(scope="module")
def checker():
    src = EXE + ".cpp"
    hdr = os.path.join(ROOT, "spicey_b200", "csrc", "band_plan.h")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EXE, src])
    return EXE


# (kind, seed, size, nV, L, RPL): kind 0 mesh in netlist numbering, 1 random banded RC, 2 ladder, 3 random banded RLC
CASES = [
    (0, 1, 16, 1, 0, 0),     # cfg 4: renumbered to half-bandwidth 16, 8 lanes x 2 rows
    (0, 1, 16, 1, 16, 1),    # the same with 16 lanes x 1 row and 32 x 1
    (0, 1, 16, 1, 32, 1),
    (0, 2, 8, 1, 0, 0), (0, 3, 5, 1, 0, 0), (0, 4, 3, 1, 0, 0),
    (2, 1, 64, 1, 0, 0), (2, 1, 400, 1, 0, 0), (2, 1, 3, 1, 0, 0),
    (1, 1, 100, 1, 0, 0), (1, 2, 100, 1, 0, 0), (1, 5, 100, 1, 0, 0), (1, 7, 100, 1, 0, 0),
    (1, 1, 300, 2, 0, 0), (1, 2, 300, 2, 0, 0), (1, 3, 200, 3, 0, 0), (1, 4, 150, 4, 0, 0), (1, 6, 37, 2, 0, 0),
    (3, 1, 100, 1, 0, 0), (3, 2, 120, 2, 0, 0),
]


@pytest.mark.parametrize("kind,seed,size,nv,L,RPL", CASES)
def test_band_schedule_matches_dense_pivoted_elimination(checker, kind, seed, size, nv, L, RPL):
    args = [checker, str(kind), str(seed), str(size), str(nv)] + ([str(L), str(RPL)] if L else [])
    out = subprocess.run(args, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


@pytest.mark.parametrize("lanes,rpl,nb,warps,minb,rc_only,umode", [(8, 2, 1, 4, 2, True, 2), (8, 2, 1, 4, 2, True, 0), (16, 1, 2, 4, 2, False, 2),
                                                                    (4, 1, 1, 4, 2, True, 2), (2, 1, 1, 4, 2, True, 0)])
def test_band_kernel_compiles_for_both_store_modes(tmp_path, lanes, rpl, nb, warps, minb, rc_only, umode):
    """The kernel text the library hands to NVRTC for a band shape, compiled here with nvcc for sm_100a: within the register
    file at the launch shape, and with BAND_UMODE 2 the pivot rows leave through the TMA unit (UTMASTG, one per unrolled
    step) instead of per-lane global stores."""
    import re
    import shutil
    from spicey_b200 import native
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(nvcc) or not os.path.exists(cuobjdump):
        pytest.skip("nvcc / cuobjdump not found")
    src = native.band_kernel_source(lanes, rpl, nb, 0, True, warps, minb, rc_only=rc_only, umode=umode)
    assert "#define BAND_L %d\n#define BAND_RPL %d\n#define BAND_NB %d\n" % (lanes, rpl, nb) in src and "#define BAND_UMODE %d\n" % umode in src
    cu = tmp_path / "band.cu"
    cu.write_text(src)
    res = subprocess.run([nvcc, "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xptxas", "-v",
                          "-o", str(tmp_path / "band.cubin"), str(cu)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    regs = int(re.search(r"Used (\d+) registers", res.stderr).group(1))
    assert regs * warps * 32 * minb <= 65536, (regs, warps, minb)
    assert max(int(v) for v in re.findall(r"(\d+) bytes spill stores", res.stderr)) <= 64, res.stderr
    sass = subprocess.run([cuobjdump, "-sass", str(tmp_path / "band.cubin")], capture_output=True, text=True).stdout
    n_tma = len(re.findall(r"UTMASTG", sass))
    W = lanes * rpl
    assert n_tma == (W if umode == 2 else 0), (n_tma, W)
    if umode == 2:
        assert "cp.async.bulk.tensor.3d.global.shared::cta" in src and "FENCE.VIEW.ASYNC" in sass
