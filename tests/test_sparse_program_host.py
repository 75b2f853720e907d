"""CPU check of the sparse-program builder (spicey_b200/csrc/sparse_program.h): random sparse complex systems,
program built from the pilot, executed by a reference interpreter with the device's semantics
(tests/cpp/sparse_program_check.cpp) and compared with dense partial-pivoting elimination."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "sparse_program_check")


@pytest.fixture(scope="module")
def checker():
    src = EXE + ".cpp"
    hdr = os.path.join(ROOT, "spicey_b200", "csrc", "sparse_program.h")
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", EXE, src])
    return EXE


CASES = [(s, n, d, k) for s, (n, d, k) in enumerate([
    (1, 50, 0), (2, 100, 0), (3, 60, 2), (5, 80, 0), (12, 30, 0), (12, 30, 3), (33, 12, 0), (40, 10, 5), (65, 5, 4),
    (65, 5, 0), (100, 8, 3), (30, 100, 0), (64, 3, 2), (129, 2, 0)], start=1)]


@pytest.mark.parametrize("seed,n,density,distinct", CASES)
def test_program_matches_dense_pivoted_elimination(checker, seed, n, density, distinct):
    out = subprocess.run([checker, str(seed), str(n), str(density), str(distinct)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
