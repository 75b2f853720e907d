"""Synthetic netlists and sweep values of BASELINE.json's five configs.

Exact recipes from SURVEY.md §8(d): every netlist is text the reference's own
grammar parses (a `*` comment line first, so no title line is mistaken for an
element: parseNetlist.ts:158-161), values avoid the Complex.div guard (H5),
per-instance tolerances come from splitmix64 with seed 0x5EED5EED.
"""
from __future__ import annotations

import numpy as np

SEED = 0x5EED5EED
_M64 = (1 << 64) - 1

README_RC = """
Demo of a simple AC circuit

v1 1 0 dc 0 ac 1
r1 1 2 30
c1 2 0 100u
.ac dec 100 1 100

.end
"""


def rc_ladder(n_nodes: int = 64, ppd: int = 200000, f1: str = "1", f2: str = "100k") -> str:
    """cfg 2: v1 n1 0 ac 1; r{k} n{k} n{k+1} 1k; c{k} n{k+1} 0 1n, k=1..n-1."""
    lines = ["* %d-node RC ladder AC sweep" % n_nodes, "v1 n1 0 ac 1"]
    for k in range(1, n_nodes):
        lines.append("r%d n%d n%d 1k" % (k, k, k + 1))
    for k in range(1, n_nodes):
        lines.append("c%d n%d 0 1n" % (k, k + 1))
    lines.append(".ac dec %d %s %s" % (ppd, f1, f2))
    lines.append(".end")
    return "\n".join(lines) + "\n"


def rc_mesh(side: int = 16, ppd: int = 1600000, f1: str = "1", f2: str = "100k") -> str:
    """cfg 4: side x side grid, R 1k between 4-neighbours, C 1n to ground except n0_0."""
    lines = ["* %dx%d RC mesh AC sweep" % (side, side), "v1 n0_0 0 ac 1"]
    k = 0
    for r in range(side):
        for c in range(side):
            if c + 1 < side:
                k += 1
                lines.append("r%d n%d_%d n%d_%d 1k" % (k, r, c, r, c + 1))
            if r + 1 < side:
                k += 1
                lines.append("r%d n%d_%d n%d_%d 1k" % (k, r, c, r + 1, c))
    k = 0
    for r in range(side):
        for c in range(side):
            if r == 0 and c == 0:
                continue
            k += 1
            lines.append("c%d n%d_%d 0 1n" % (k, r, c))
    lines.append(".ac dec %d %s %s" % (ppd, f1, f2))
    lines.append(".end")
    return "\n".join(lines) + "\n"


def rc_dense(n_nodes: int = 64, ppd: int = 200000, f1: str = "10k", f2: str = "1g") -> str:
    """A dense MNA system of cfg 2's size (Nvar = n_nodes + 1): a resistor between EVERY pair of nodes (values spread
    over a decade so that no two matrix entries coincide), a capacitor to ground at every node but the driven one.
    The nodal matrix has no structural zero: the workload of the dense pivoting LU (lib/math/solveComplex.ts:15-53
    without its `|f| < EPS` shortcut ever applying).  The sweep (five decades, 1,000,001 points like cfg 2) covers the
    range where the capacitors load the network (w C ~ G around 1 MHz): below it every node sits at the source voltage
    and the resistor currents are differences of nearly equal numbers."""
    lines = ["* %d-node complete RC graph AC sweep" % n_nodes, "v1 n1 0 ac 1"]
    k = 0
    for i in range(1, n_nodes + 1):
        for j in range(i + 1, n_nodes + 1):
            k += 1
            lines.append("r%d n%d n%d %d" % (k, i, j, 1000 + 37 * ((i * 7 + j * 13) % 251)))
    for i in range(2, n_nodes + 1):
        lines.append("c%d n%d 0 %dp" % (i, i, 500 + 17 * i))
    lines.append(".ac dec %d %s %s" % (ppd, f1, f2))
    lines.append(".end")
    return "\n".join(lines) + "\n"


RLC_TANK = """* RLC tank Monte-Carlo
V1 1 0 PULSE(0 1 0 1n 1n 1 2)
R1 1 2 50
L1 2 0 1m
C1 2 0 1u
.tran 1u 1m
.end
"""

RECTIFIER = """* half-wave rectifier sweep
V1 in 0 PULSE(-5 5 0 0.5m 0.5m 0 1m)
D1 in out DMOD
R1 out 0 1k
C1 out 0 1u
.model DMOD D(Is=1e-14 N=1)
.tran 1u 3m
.end
"""


class SplitMix64:
    def __init__(self, seed: int = SEED):
        self.s = seed & _M64

    def next_u64(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform_pm1(self) -> float:
        return (self.next_u64() >> 11) * (2.0 / (1 << 53)) - 1.0


def splitmix_uniform_pm1(n_inst: int, n_draws: int, seed: int = SEED) -> np.ndarray:
    """[n_inst][n_draws] of U(-1,1): instance i draws in element order (vectorised splitmix64)."""
    idx = np.arange(1, n_inst * n_draws + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0
    return u.reshape(n_inst, n_draws)


def rlc_tank_overrides(n_inst: int = 65536):
    """cfg 3: +-5 % uniform tolerances on R1, L1, C1 (nominal*(1+0.05u))."""
    u = splitmix_uniform_pm1(n_inst, 3)
    return {
        "R1": 50.0 * (1 + 0.05 * u[:, 0]),
        "L1": (1 * 1e-3) * (1 + 0.05 * u[:, 1]),
        "C1": (1 * 1e-6) * (1 + 0.05 * u[:, 2]),
    }


def rectifier_overrides(n_inst: int = 100000):
    """cfg 5: R1 in logspace(100,100k), C1 +-5 %, Is in logspace(1e-15,1e-12)."""
    u = splitmix_uniform_pm1(n_inst, 1)
    frac = np.arange(n_inst, dtype=np.float64) / max(1, n_inst - 1)
    return {
        "R1": 100.0 * np.power(10.0, 3.0 * frac),
        "C1": (1 * 1e-6) * (1 + 0.05 * u[:, 0]),
        "D1.is": 1e-15 * np.power(10.0, 3.0 * frac),
    }
