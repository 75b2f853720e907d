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


def _rc_tree(n, ppd=600):
    """A random sparse RC network: a resistor tree grown from the source's node, a capacitor at every node, n / 4 chords.
    Far nodes are attenuated by many orders of magnitude at the top of the sweep."""
    import numpy as np
    rng = np.random.default_rng(n)
    lines = ["* random sparse RC network", "v1 n1 0 dc 1 ac 1"]
    for i in range(2, n + 1):
        lines.append("r%d n%d n%d %g" % (i, i, rng.integers(1, i), rng.uniform(100, 1e4)))
    for i in range(1, n + 1):
        lines.append("c%d n%d 0 %g" % (i, i, rng.uniform(1e-9, 1e-7)))
    for k in range(n // 4):
        a, b = rng.choice(np.arange(1, n + 1), 2, replace=False)
        lines.append("r%d n%d n%d %g" % (1000 + k, a, b, rng.uniform(100, 1e4)))
    lines += [".ac dec %d 1 100k" % ppd, ".end"]
    return "\n".join(lines) + "\n"


def test_renumbered_plans_are_checked_against_the_netlist_order():
    """A renumbered band plan eliminates in another order than the reference.  Both orders are backward stable, but the
    parity bar is per entry: the plan is used by default only when the reference's algorithm, run on the host in both
    orders at the pilot and at both ends of the sweep, agrees per entry to 5e-10 (spicey_native.cu: kBandOrderTol).
    cfg 4's mesh agrees to ~1e-11; random RC trees with chords, whose far nodes are attenuated by 1e-12 and more at
    100 kHz, differ by 3e-9 (60 nodes) and 3e-6 (100 nodes) there and are left to the tiers that keep the netlist order
    (measured on the GPU before this check existed: 1e-8 / 3e-7 per entry against the oracle through the banded tier,
    1e-13 relative to the largest unknown)."""
    from spicey_b200 import native, packing, parsing, workloads
    mesh = packing.pack_circuit(parsing.parse_netlist(workloads.rc_mesh(16)))
    assert native.band_plan_stats(mesh, 300.0)["renumbered"] == 1
    for f in (1.0, 316.0, 1e5):
        d = native.band_order_deviation(mesh, 300.0, f)
        assert 0.0 <= d < 1e-10, (f, d)
    ladder = packing.pack_circuit(parsing.parse_netlist(workloads.rc_ladder(64)))
    assert native.band_order_deviation(ladder, 300.0, 1e5) == 0.0   # the netlist order is kept: nothing to compare
    for n, lo in ((60, 1e-9), (100, 1e-7)):
        tree = packing.pack_circuit(parsing.parse_netlist(_rc_tree(n)))
        assert native.band_plan_stats(tree, 300.0)["renumbered"] == 1
        assert native.band_order_deviation(tree, 300.0, 316.0) < 5e-10      # harmless in the middle of the sweep,
        assert native.band_order_deviation(tree, 300.0, 1e5) > lo           # not at its top: declined
