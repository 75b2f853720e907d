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
