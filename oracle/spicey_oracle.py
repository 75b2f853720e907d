"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Pure-Python, line-by-line CPU restatement of the reference's hot path, used
only as the parity checker (tests/, __graft_entry__.smoke(), bench.py's
cpu_baseline / --impl reference legs).  Nothing under spicey_b200/ imports it.

Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
every known-answer vector the reference's own tests hold for the path
(tests/golden/, extracted by tests/golden/make_golden.py from
tests/basics/basics01.test.ts:19-220 and the five SVG snapshots under
tests/transient/__snapshots__/).

Every function cites the reference file:line it follows.  Python floats are
IEEE-754 binary64 and no operation here is fused, so the arithmetic sequence is
the reference's; the only engine-dependent pieces are libm's
exp/hypot/sin/cos/atan2 (may differ from the JS engine's by <= 1 ulp).
"""
from __future__ import annotations

import math
from typing import Dict, List

EPS = 1e-15  # lib/constants/EPS.ts:1
VT_300K = 0.02585  # lib/constants/physics.ts:1


class SingularMatrixError(ArithmeticError):
    pass


class ComplexDivideError(ArithmeticError):
    pass


class Complex:
    """lib/math/Complex.ts:3-62 (immutable; textbook unscaled div)."""

    __slots__ = ("re", "im")

    def __init__(self, re=0.0, im=0.0):
        self.re = re
        self.im = im

    @staticmethod
    def from_polar(mag, deg=0.0):  # Complex.ts:16-19
        ph = (deg * math.pi) / 180
        return Complex(mag * math.cos(ph), mag * math.sin(ph))

    def add(self, b):  # :25-27
        return Complex(self.re + b.re, self.im + b.im)

    def sub(self, b):  # :29-31
        return Complex(self.re - b.re, self.im - b.im)

    def mul(self, b):  # :33-38
        return Complex(self.re * b.re - self.im * b.im, self.re * b.im + self.im * b.re)

    def div(self, b):  # :40-47
        d = b.re * b.re + b.im * b.im
        if d < EPS:
            raise ComplexDivideError("Complex divide by ~0")
        return Complex((self.re * b.re + self.im * b.im) / d, (self.im * b.re - self.re * b.im) / d)

    def abs(self):  # :55-57
        return math.hypot(self.re, self.im)

    def phaseDeg(self):  # :59-61
        return (math.atan2(self.im, self.re) * 180) / math.pi

    def __complex__(self):
        return complex(self.re, self.im)

    def __repr__(self):
        return "Complex(%r, %r)" % (self.re, self.im)


def solve_complex(A: List[List[Complex]], b: List[Complex]) -> List[Complex]:
    """lib/math/solveComplex.ts:4-73."""
    n = len(A)
    for i in range(n):  # :6-13 augment
        A[i] = [Complex(z.re, z.im) for z in A[i]] + [Complex(b[i].re, b[i].im)]
    for k in range(n):  # :15
        imax = k
        vmax = A[k][k].abs()
        for i in range(k + 1, n):  # :20-28 first max wins (strict >)
            v = A[i][k].abs()
            if v > vmax:
                vmax = v
                imax = i
        if vmax < EPS:  # :29
            raise SingularMatrixError("Singular matrix (complex)")
        if imax != k:  # :30-34
            A[k], A[imax] = A[imax], A[k]
        prow = A[k]
        pivot = prow[k]
        for i in range(k + 1, n):  # :40-53
            row = A[i]
            f = row[k].div(pivot)
            if f.abs() < EPS:  # :46
                continue
            for j in range(k, n + 1):
                row[j] = row[j].sub(f.mul(prow[j]))
    x: List[Complex] = [None] * n  # type: ignore
    for i in range(n - 1, -1, -1):  # :56-71
        row = A[i]
        s = row[n]
        for j in range(i + 1, n):
            s = s.sub(row[j].mul(x[j]))
        x[i] = s.div(row[i])
    return x


def solve_real(A: List[List[float]], b: List[float]) -> List[float]:
    """lib/math/solveReal.ts:3-73."""
    n = len(A)
    for i in range(n):
        A[i] = list(A[i]) + [b[i]]
    for k in range(n):
        imax = k
        vmax = abs(A[k][k])
        for i in range(k + 1, n):
            v = abs(A[i][k])
            if v > vmax:
                vmax = v
                imax = i
        if vmax < EPS:  # :28
            raise SingularMatrixError("Singular matrix (real)")
        if imax != k:
            A[k], A[imax] = A[imax], A[k]
        prow = A[k]
        pivot = prow[k]
        for i in range(k + 1, n):
            row = A[i]
            f = row[k] / pivot
            if abs(f) < EPS:
                continue
            for j in range(k, n + 1):
                row[j] = row[j] - f * prow[j]
    x = [0.0] * n
    for i in range(n - 1, -1, -1):
        row = A[i]
        s = row[n]
        for j in range(i + 1, n):
            s -= row[j] * x[j]
        x[i] = s / row[i]
    return x


# --- stamps (lib/stamping/*.ts) -------------------------------------------------

def _mi(node_id):  # NodeIndex.ts:28-31
    return -1 if node_id == 0 else node_id - 1


def stamp_admittance_complex(A, n1, n2, Y):  # stampAdmittanceComplex.ts:4-30
    i1, i2 = _mi(n1), _mi(n2)
    if i1 >= 0:
        A[i1][i1] = A[i1][i1].add(Y)
    if i2 >= 0:
        A[i2][i2] = A[i2][i2].add(Y)
    if i1 >= 0 and i2 >= 0:
        A[i1][i2] = A[i1][i2].sub(Y)
        A[i2][i1] = A[i2][i1].sub(Y)


def stamp_voltage_source_complex(A, b, vs, V):  # stampVoltageSourceComplex.ts:5-35
    i1, i2, j = _mi(vs.n1), _mi(vs.n2), vs.index
    one = Complex(1, 0)
    if i1 >= 0:
        A[i1][j] = A[i1][j].add(one)
    if i2 >= 0:
        A[i2][j] = A[i2][j].sub(one)
    if i1 >= 0:
        A[j][i1] = A[j][i1].add(one)
    if i2 >= 0:
        A[j][i2] = A[j][i2].sub(one)
    b[j] = b[j].add(V)


def stamp_admittance_real(A, n1, n2, Y):  # stampAdmittanceReal.ts:3-29
    i1, i2 = _mi(n1), _mi(n2)
    if i1 >= 0:
        A[i1][i1] = A[i1][i1] + Y
    if i2 >= 0:
        A[i2][i2] = A[i2][i2] + Y
    if i1 >= 0 and i2 >= 0:
        A[i1][i2] = A[i1][i2] - Y
        A[i2][i1] = A[i2][i1] - Y


def stamp_current_real(b, n_plus, n_minus, current):  # stampCurrentReal.ts:3-14
    ip, im = _mi(n_plus), _mi(n_minus)
    if ip >= 0:
        b[ip] = b[ip] - current
    if im >= 0:
        b[im] = b[im] + current


def stamp_current_complex(b, n_plus, n_minus, current):  # stampCurrentComplex.ts:4-15
    ip, im = _mi(n_plus), _mi(n_minus)
    if ip >= 0:
        b[ip] = b[ip].sub(current)
    if im >= 0:
        b[im] = b[im].add(current)


def stamp_voltage_source_real(A, b, vs, V):  # stampVoltageSourceReal.ts:4-32
    i1, i2, j = _mi(vs.n1), _mi(vs.n2), vs.index
    if i1 >= 0:
        A[i1][j] = A[i1][j] + 1
    if i2 >= 0:
        A[i2][j] = A[i2][j] - 1
    if i1 >= 0:
        A[j][i1] = A[j][i1] + 1
    if i2 >= 0:
        A[j][i2] = A[j][i2] - 1
    b[j] = b[j] + V


# --- AC (lib/analysis/simulateAC.ts) ---------------------------------------------

def _inductor_admittance(f, L):  # simulateAC.ts:47-51 / :111-115
    denom = Complex(0, 2 * math.pi * f * L)
    return Complex(0, 0) if denom.abs() < EPS else Complex(1, 0).div(denom)


def build_linear_system_for_ac(ckt, f, nvar):  # simulateAC.ts:24-60
    A = [[Complex(0, 0) for _ in range(nvar)] for _ in range(nvar)]
    b = [Complex(0, 0) for _ in range(nvar)]
    two_pi = 2 * math.pi
    for r in ckt.R:
        if r.R <= 0:
            raise ValueError("R %s must be > 0" % r.name)
        stamp_admittance_complex(A, r.n1, r.n2, Complex(1 / r.R, 0))
    for c in ckt.C:
        stamp_admittance_complex(A, c.n1, c.n2, Complex(0, two_pi * f * c.C))
    for l in ckt.L:
        stamp_admittance_complex(A, l.n1, l.n2, _inductor_admittance(f, l.L))
    for vs in ckt.V:
        stamp_voltage_source_complex(A, b, vs, Complex.from_polar(vs.acMag or 0, vs.acPhaseDeg or 0))
    for cs in getattr(ckt, "I", []):  # extension: the reference ships stampCurrentComplex.ts:4-15 but parses no I line
        stamp_current_complex(b, cs.n1, cs.n2, Complex.from_polar(cs.acMag or 0, cs.acPhaseDeg or 0))
    return A, b


def simulate_ac(ckt, freqs=None):
    """simulateAC.ts:62-130.  `freqs` overrides the directive's list (used to
    evaluate a subsample of a large sweep with the identical per-point code)."""
    if ckt.analyses.ac is None and freqs is None:
        return None
    from spicey_b200.parsing import build_frequency_array  # host-side list builder (simulateAC.ts:9-22)

    nvar = (ckt.nodes.count() - 1) + len(ckt.V)
    if freqs is None:
        a = ckt.analyses.ac
        freqs = build_frequency_array(a.mode, a.N, a.f1, a.f2)
    node_voltages: Dict[str, list] = {}
    for nid, name in enumerate(ckt.nodes.rev):
        if nid != 0:
            node_voltages[name] = []
    element_currents: Dict[str, list] = {}
    two_pi = 2 * math.pi
    zero = Complex(0, 0)

    def volt(x, n):
        return zero if n == 0 else x[n - 1]

    for f in freqs:  # :80
        A, b = build_linear_system_for_ac(ckt, f, nvar)
        x = solve_complex(A, b)
        for nid in range(1, ckt.nodes.count()):
            node_voltages[ckt.nodes.rev[nid]].append(x[nid - 1])
        for r in ckt.R:  # :94-100
            element_currents.setdefault(r.name, []).append(
                Complex(1 / r.R, 0).mul(volt(x, r.n1).sub(volt(x, r.n2))))
        for c in ckt.C:  # :101-107
            element_currents.setdefault(c.name, []).append(
                Complex(0, two_pi * f * c.C).mul(volt(x, c.n1).sub(volt(x, c.n2))))
        for l in ckt.L:  # :108-118
            element_currents.setdefault(l.name, []).append(
                _inductor_admittance(f, l.L).mul(volt(x, l.n1).sub(volt(x, l.n2))))
        for vs in ckt.V:  # :119-122
            element_currents.setdefault(vs.name, []).append(x[vs.index])
    return {"freqs": list(freqs), "nodeVoltages": node_voltages, "elementCurrents": element_currents}


# --- TRAN (lib/analysis/simulateTRAN.ts) -------------------------------------------

def compute_effective_time_step(dt_requested, tstop):  # simulateTRAN.ts:14-19
    dt_eff = dt_requested if dt_requested > EPS else max(tstop / 1000, EPS)
    steps = max(1, math.ceil(tstop / max(dt_eff, EPS)))
    dt = tstop / steps if steps > 0 else tstop
    return dt, steps


def stamp_all_elements_at_time(A, b, ckt, t, dt, x, it):  # simulateTRAN.ts:25-102
    for r in ckt.R:
        stamp_admittance_real(A, r.n1, r.n2, 1 / r.R)
    for c in ckt.C:
        Gc = c.C / max(dt, EPS)
        stamp_admittance_real(A, c.n1, c.n2, Gc)
        stamp_current_real(b, c.n1, c.n2, -Gc * c.vPrev)
    for l in ckt.L:
        Gl = max(dt, EPS) / l.L
        stamp_admittance_real(A, l.n1, l.n2, Gl)
        stamp_current_real(b, l.n1, l.n2, l.iPrev)
    for sw in ckt.S:
        if sw.model is None:
            continue
        Rv = sw.model.Ron if sw.isOn else sw.model.Roff
        stamp_admittance_real(A, sw.n1, sw.n2, 1 / max(abs(Rv), EPS))
    for vs in ckt.V:
        Vt = vs.waveform(t) if vs.waveform else (vs.dc or 0)
        stamp_voltage_source_real(A, b, vs, Vt)
    for d in ckt.D:
        if d.model is None:
            continue
        vp = 0 if d.nPlus == 0 else x[d.nPlus - 1]
        vn = 0 if d.nMinus == 0 else x[d.nMinus - 1]
        vd = d.vdPrev if it == 0 else vp - vn  # :85 (hazard H3)
        vth = d.model.N * VT_300K
        vlim = vd
        if vd > 0.8:
            vlim = 0.8
        if vd < -1.0:
            vlim = -1.0
        e = math.exp(vlim / vth)
        idd = d.model.Is * (e - 1)
        gd = max((d.model.Is / vth) * e, 1e-12)
        ieq = idd - gd * vlim
        stamp_admittance_real(A, d.nPlus, d.nMinus, gd)
        stamp_current_real(b, d.nPlus, d.nMinus, ieq)
    for cs in getattr(ckt, "I", []):  # extension: constant current, stampCurrentReal.ts:3-14
        stamp_current_real(b, cs.n1, cs.n2, cs.dc or 0)


def update_switch_states_from_solution(ckt, x):  # simulateTRAN.ts:108-128
    switched = False
    for sw in ckt.S:
        if sw.model is None:
            continue
        vp = 0 if sw.ncPos == 0 else x[sw.ncPos - 1]
        vn = 0 if sw.ncNeg == 0 else x[sw.ncNeg - 1]
        vctrl = vp - vn
        nxt = sw.isOn
        if sw.isOn:
            if vctrl < sw.model.Voff:
                nxt = False
        elif vctrl > sw.model.Von:
            nxt = True
        if nxt != sw.isOn:
            sw.isOn = nxt
            switched = True
    return switched


def simulate_tran(ckt):
    """simulateTRAN.ts:130-252 (mutates ckt state exactly as the reference)."""
    if ckt.analyses.tran is None:
        return None
    dt, steps = compute_effective_time_step(ckt.analyses.tran.dt, ckt.analyses.tran.tstop)
    nvar = (ckt.nodes.count() - 1) + len(ckt.V)
    times = []
    node_voltages: Dict[str, list] = {}
    for nid, name in enumerate(ckt.nodes.rev):
        if nid != 0:
            node_voltages[name] = []
    element_currents: Dict[str, list] = {}

    def volt(x, n):
        return 0 if n == 0 else x[n - 1]

    for step in range(steps + 1):  # :146-147  t = step*dt, never accumulated
        t = step * dt
        times.append(t)
        x = [0.0] * nvar  # :149 (hazard H4)
        for it in range(20):  # :151
            A = [[0.0] * nvar for _ in range(nvar)]
            b = [0.0] * nvar
            stamp_all_elements_at_time(A, b, ckt, t, dt, x, it)
            x = solve_real(A, b)
            if not update_switch_states_from_solution(ckt, x):
                break
        for nid in range(1, ckt.nodes.count()):
            node_voltages[ckt.nodes.rev[nid]].append(x[nid - 1])
        for r in ckt.R:  # :173-178
            element_currents.setdefault(r.name, []).append((volt(x, r.n1) - volt(x, r.n2)) / r.R)
        for c in ckt.C:  # :179-184 (old vPrev: H6)
            element_currents.setdefault(c.name, []).append(
                (c.C * (volt(x, c.n1) - volt(x, c.n2) - c.vPrev)) / max(dt, EPS))
        for l in ckt.L:  # :185-191
            Gl = max(dt, EPS) / l.L
            element_currents.setdefault(l.name, []).append(Gl * (volt(x, l.n1) - volt(x, l.n2)) + l.iPrev)
        for vs in ckt.V:  # :192-195
            element_currents.setdefault(vs.name, []).append(x[vs.index])
        for sw in ckt.S:  # :196-205
            if sw.model is None:
                continue
            Rv = sw.model.Ron if sw.isOn else sw.model.Roff
            element_currents.setdefault(sw.name, []).append(
                (volt(x, sw.n1) - volt(x, sw.n2)) / max(abs(Rv), EPS))
        for d in ckt.D:  # :208-219 (unclamped vd: H6)
            if d.model is None:
                continue
            vd = volt(x, d.nPlus) - volt(x, d.nMinus)
            try:
                e = math.exp(vd / (d.model.N * VT_300K))
            except OverflowError:
                e = math.inf
            element_currents.setdefault(d.name, []).append(d.model.Is * (e - 1))
        for cs in getattr(ckt, "I", []):  # extension: the source's own value
            element_currents.setdefault(cs.name, []).append(cs.dc or 0)
        for c in ckt.C:  # :221-225
            c.vPrev = volt(x, c.n1) - volt(x, c.n2)
        for l in ckt.L:  # :226-231
            Gl = max(dt, EPS) / l.L
            l.iPrev = Gl * (volt(x, l.n1) - volt(x, l.n2)) + l.iPrev
        for d in ckt.D:  # :233-237
            d.vdPrev = volt(x, d.nPlus) - volt(x, d.nMinus)

    if len(ckt.probes.tran) > 0:  # :240-249
        upper = [p.upper() for p in ckt.probes.tran]
        node_voltages = {k: v for k, v in node_voltages.items() if k.upper() in upper}
    return {"times": times, "nodeVoltages": node_voltages, "elementCurrents": element_currents}


def simulate(text):  # lib/analysis/simulate.ts:5-10
    from spicey_b200.parsing import parse_netlist

    ckt = parse_netlist(text)
    return {"circuit": ckt, "ac": simulate_ac(ckt), "tran": simulate_tran(ckt)}
