"""Host-side netlist front end: text -> ParsedCircuit.

This is the Python mirror of the reference's L1 layer, which the
north star keeps on the host.  It exists so that inputs to the CUDA path are
built with exactly the reference's arithmetic (hazard H1: ``3m`` is
``3 * 1e-3``, never the literal ``0.003``) and so that the parity tests read
like the reference's own tests.  It is not on the data-parallel hot path.

Reference behaviour mirrored (read-only, never copied):
  lib/parsing/parseNetlist.ts:109-481      tokeniser, directives, elements
  lib/parsing/parseNumberWithUnits.ts:1-30 unit suffix arithmetic
  lib/parsing/parsePulseArgs.ts:4-25       PULSE(...)
  lib/parsing/parsePwlArgs.ts:3-18         PWL(...)
  lib/parsing/pulseValue.ts:4-22           PULSE value at t
  lib/parsing/pwlValue.ts:3-16             PWL value at t
  lib/parsing/NodeIndex.ts:1-32            case-folding node index
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional

EPS = 1e-15  # lib/constants/EPS.ts:1
VT_300K = 0.02585  # lib/constants/physics.ts:1

_FLOAT_PREFIX = re.compile(r"^\s*[+-]?(?:Infinity|\d+\.?\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?)")
_INT_PREFIX = re.compile(r"^\s*[+-]?\d+")
_PLAIN_NUMBER = re.compile(r"^[+-]?\d*\.?\d+(?:[eE][+-]?\d+)?$")
_NUMBER_SUFFIX = re.compile(r"^([+-]?\d*\.?\d+(?:[eE][+-]?\d+)?)([a-zA-Z]+)$")
_UNIT_TAIL = re.compile(r"(ohm|v|a|s|h|f)$")
_UNIT_MUL = {
    "t": 1e12,
    "g": 1e9,
    "meg": 1e6,
    "k": 1e3,
    "m": 1e-3,
    "u": 1e-6,
    "n": 1e-9,
    "p": 1e-12,
    "f": 1e-15,
}


def js_parse_float(s: str) -> float:
    """ECMAScript ``parseFloat``: longest numeric prefix, else NaN."""
    m = _FLOAT_PREFIX.match(s)
    if not m:
        return math.nan
    txt = m.group(0).strip()
    if txt.endswith("Infinity"):
        return -math.inf if txt.startswith("-") else math.inf
    return float(txt)


def js_parse_int(s: str) -> float:
    """ECMAScript ``parseInt(s, 10)``; NaN when no digits lead the string."""
    m = _INT_PREFIX.match(s)
    if not m:
        return math.nan
    return int(m.group(0))


def parse_number_with_units(raw) -> float:
    """SPICE number with unit suffix -> double (parseNumberWithUnits.ts:1-30).

    The product ``parseFloat(num) * multiplier`` is performed in binary64 just
    as the reference does, so ``20u`` is 1.9999999999999998e-05.
    """
    if raw is None:
        return math.nan
    s = str(raw).strip()
    if s == "":
        return math.nan
    if _PLAIN_NUMBER.match(s):
        return js_parse_float(s)
    m = _NUMBER_SUFFIX.match(s)
    if not m:
        return js_parse_float(s)
    val = js_parse_float(m.group(1))
    suf = m.group(2).lower()
    suf = _UNIT_TAIL.sub("", suf)
    if suf == "meg":
        return val * _UNIT_MUL["meg"]
    if len(suf) == 1 and suf in _UNIT_MUL:
        return val * _UNIT_MUL[suf]
    return val


@dataclass
class PulseSpec:  # lib/types/simulation.ts:1-10
    v1: float
    v2: float
    td: float
    tr: float
    tf: float
    ton: float
    period: float
    ncycles: float


def _split_args(token: str, keyword: str) -> List[str]:
    clean = re.sub(r"^%s\s*\(" % keyword, "(", token.strip(), flags=re.I)
    inside = re.sub(r"\)$", "", re.sub(r"^\(", "", clean)).strip()
    return [p for p in re.split(r"[\s,]+", inside) if len(p)]


def parse_pulse_args(token: str) -> PulseSpec:
    parts = _split_args(token, "pulse")
    if len(parts) < 7:
        raise ValueError("PULSE(...) requires 7 or 8 args")
    vals = [parse_number_with_units(p) for p in parts]
    if any(math.isnan(v) for v in vals):
        raise ValueError("Invalid PULSE() numeric value")
    return PulseSpec(
        v1=vals[0], v2=vals[1], td=vals[2], tr=vals[3], tf=vals[4], ton=vals[5],
        period=vals[6], ncycles=vals[7] if len(parts) > 7 else math.inf,
    )


def parse_pwl_args(token: str):
    parts = _split_args(token, "pwl")
    if len(parts) == 0 or len(parts) % 2 != 0:
        raise ValueError("PWL(...) requires an even number of time/value pairs")
    pairs = []
    for i in range(0, len(parts), 2):
        t = parse_number_with_units(parts[i])
        v = parse_number_with_units(parts[i + 1])
        if math.isnan(t) or math.isnan(v):
            raise ValueError("Invalid PWL() numeric value")
        pairs.append((t, v))
    return pairs


def pulse_value(p: PulseSpec, t: float) -> float:
    """PULSE source value at time t (pulseValue.ts:4-22)."""
    if t < p.td:
        return p.v1
    tt = t - p.td
    cycles_done = math.floor(tt / p.period) if p.period != 0 else (math.inf if tt > 0 else math.nan)
    if cycles_done >= p.ncycles:
        return p.v1
    tc = tt - cycles_done * p.period
    if tc < p.tr:
        a = tc / max(p.tr, EPS)
        return p.v1 + (p.v2 - p.v1) * a
    if tc < p.tr + p.ton:
        return p.v2
    if tc < p.tr + p.ton + p.tf:
        a = (tc - (p.tr + p.ton)) / max(p.tf, EPS)
        return p.v2 + (p.v1 - p.v2) * a
    return p.v1


def pwl_value(pairs, t: float) -> float:
    """PWL source value at time t (pwlValue.ts:3-16)."""
    if len(pairs) == 0:
        return 0.0
    if t <= pairs[0][0]:
        return pairs[0][1]
    for i in range(1, len(pairs)):
        pt, pv = pairs[i - 1]
        ct, cv = pairs[i]
        if t <= ct:
            dt = max(ct - pt, EPS)
            a = (t - pt) / dt
            return pv + (cv - pv) * a
    return pairs[-1][1]


class NodeIndex:
    """Case-insensitive node name -> id, first spelling canonical (NodeIndex.ts)."""

    def __init__(self):
        self._map: Dict[str, int] = {"0": 0}
        self.rev: List[str] = ["0"]

    def get_or_create(self, name: str) -> int:
        orig = str(name)
        key = orig.upper()
        if key in self._map:
            return self._map[key]
        idx = len(self.rev)
        self._map[key] = idx
        self.rev.append(orig)
        return idx

    def get(self, name: str) -> Optional[int]:
        return self._map.get(str(name).upper())

    def count(self) -> int:
        return len(self.rev)

    @staticmethod
    def matrix_index_of_node(node_id: int) -> int:
        return -1 if node_id == 0 else node_id - 1


@dataclass
class Resistor:
    name: str
    n1: int
    n2: int
    R: float


@dataclass
class Capacitor:
    name: str
    n1: int
    n2: int
    C: float
    vPrev: float = 0.0


@dataclass
class Inductor:
    name: str
    n1: int
    n2: int
    L: float
    iPrev: float = 0.0


@dataclass
class VoltageSource:
    name: str
    n1: int
    n2: int
    dc: float = 0.0
    acMag: float = 0.0
    acPhaseDeg: float = 0.0
    waveform: Optional[Callable[[float], float]] = None
    index: int = -1
    # Kept beside the closure: the reference hides these in
    # the closure; device-side evaluation for source sweeps needs them.
    pulse: Optional[PulseSpec] = None
    pwl: Optional[list] = None


@dataclass
class CurrentSource:
    """Independent current source, n1 (n+) -> n2 (n-) through the source.  NOT part of the reference's ParsedCircuit:
    its parser skips `I` lines (parseNetlist.ts:444-446) although lib/stamping/stampCurrent{Real,Complex}.ts exist.
    parse_netlist(text, current_sources=True) is the opt-in extension that fills ckt.I (north star: "R, L, C, V and I")."""
    name: str
    n1: int
    n2: int
    dc: float = 0.0
    acMag: float = 0.0
    acPhaseDeg: float = 0.0


@dataclass
class VSwitchModel:
    name: str
    Ron: float = 1.0
    Roff: float = 1e12
    Von: float = 0.0
    Voff: float = 0.0


@dataclass
class DiodeModel:
    name: str
    Is: float = 1e-14
    N: float = 1.0


@dataclass
class Switch:
    name: str
    n1: int
    n2: int
    ncPos: int
    ncNeg: int
    modelName: str
    model: Optional[VSwitchModel] = None
    isOn: bool = False


@dataclass
class Diode:
    name: str
    nPlus: int
    nMinus: int
    modelName: str
    model: Optional[DiodeModel] = None
    vdPrev: float = 0.0


@dataclass
class ACAnalysis:
    mode: str
    N: float
    f1: float
    f2: float


@dataclass
class TranAnalysis:
    dt: float
    tstop: float


@dataclass
class Analyses:
    ac: Optional[ACAnalysis] = None
    tran: Optional[TranAnalysis] = None


@dataclass
class Probes:
    tran: List[str] = field(default_factory=list)


@dataclass
class Models:
    vswitch: Dict[str, VSwitchModel] = field(default_factory=dict)
    diode: Dict[str, DiodeModel] = field(default_factory=dict)


@dataclass
class ParsedCircuit:  # parseNetlist.ts:83-103
    nodes: NodeIndex = field(default_factory=NodeIndex)
    R: List[Resistor] = field(default_factory=list)
    C: List[Capacitor] = field(default_factory=list)
    L: List[Inductor] = field(default_factory=list)
    V: List[VoltageSource] = field(default_factory=list)
    S: List[Switch] = field(default_factory=list)
    D: List[Diode] = field(default_factory=list)
    I: List["CurrentSource"] = field(default_factory=list)   # extension, empty unless parse_netlist(current_sources=True)
    analyses: Analyses = field(default_factory=Analyses)
    probes: Probes = field(default_factory=Probes)
    skipped: List[str] = field(default_factory=list)
    models: Models = field(default_factory=Models)


_TOKEN_RE = re.compile(r'"[^"]*"|\w+\s*\([^)]*\)|\([^()]*\)|\S+', re.ASCII)
_ELEMENT_FIRST = re.compile(r"^[rclvgsmiqd]\w*$", re.I | re.ASCII)
_PROBE_RE = re.compile(r"^v\(([^)]+)\)$", re.I)


def smart_tokens(line: str) -> List[str]:
    return _TOKEN_RE.findall(line)


def _require(tokens: List[str], index: int, context: str) -> str:
    if index >= len(tokens):
        raise ValueError(context)
    return tokens[index]


def _parse_model_params(params: str) -> List[tuple]:
    out = []
    if len(params) > 0:
        for assignment in [a for a in re.split(r"[\s,]+", params) if a]:
            bits = assignment.split("=")
            if len(bits) < 2 or not bits[0]:
                continue
            value = parse_number_with_units(bits[1])
            if math.isnan(value):
                continue
            out.append((bits[0].lower(), value))
    return out


def parse_netlist(text: str, current_sources: bool = False) -> ParsedCircuit:
    """Netlist text -> ParsedCircuit with the reference's grammar and quirks.
    current_sources=True: `I<name> n+ n- [dc] [DC v] [AC mag [phase]]` lines fill ckt.I instead of ckt.skipped (the
    reference skips them, parseNetlist.ts:444-446; same value grammar as a V line without waveforms)."""
    ckt = ParsedCircuit()
    seen_title = False

    for raw in re.split(r"\r?\n", text):
        line = raw.strip()
        if not line:
            continue
        if line.startswith("*"):
            continue
        if re.match(r"^\s*\.end\b", line, re.I):
            break
        line = re.sub(r"//.*$", "", line)
        line = re.sub(r";.*$", "", line)

        tokens = smart_tokens(line)
        if len(tokens) == 0:
            continue
        first = tokens[0]
        if len(first) == 0:
            continue

        if not seen_title and not _ELEMENT_FIRST.match(first) and not first.startswith("."):
            seen_title = True
            continue

        if first.startswith("."):
            d = first.lower()
            if d == ".ac":
                mode = _require(tokens, 1, ".ac missing mode").lower()
                if mode not in ("dec", "lin"):
                    raise ValueError(".ac supports 'dec' or 'lin'")
                N = js_parse_int(_require(tokens, 2, ".ac missing point count"))
                f1 = parse_number_with_units(_require(tokens, 3, ".ac missing start frequency"))
                f2 = parse_number_with_units(_require(tokens, 4, ".ac missing stop frequency"))
                ckt.analyses.ac = ACAnalysis(mode, N, f1, f2)
            elif d == ".tran":
                dt = parse_number_with_units(_require(tokens, 1, ".tran missing timestep"))
                tstop = parse_number_with_units(_require(tokens, 2, ".tran missing stop time"))
                ckt.analyses.tran = TranAnalysis(dt, tstop)
            elif d == ".print":
                kind = _require(tokens, 1, ".print missing analysis type").lower()
                if kind == "tran":
                    for token in tokens[2:]:
                        m = _PROBE_RE.match(token)
                        if m and m.group(1):
                            name = m.group(1)
                            if not any(p.upper() == name.upper() for p in ckt.probes.tran):
                                ckt.probes.tran.append(name)
                else:
                    ckt.skipped.append(line)
            elif d == ".model":
                name_token = _require(tokens, 1, ".model missing name")
                type_token = _require(tokens, 2, ".model missing type")
                mtype = type_token
                params = ""
                if "(" in mtype:
                    idx = mtype.index("(")
                    params = mtype[idx + 1:]
                    mtype = mtype[:idx]
                if not params:
                    rest = " ".join(tokens[3:])
                    params = re.sub(r"\)$", "", re.sub(r"^\(", "", rest))
                else:
                    rest = re.sub(r"\)$", "", " ".join(tokens[3:]))
                    params = ("%s %s" % (params, rest)).strip()
                params = re.sub(r"\)$", "", re.sub(r"^\(", "", params)).strip()
                tl = mtype.lower()
                if tl in ("vswitch", "sw"):
                    model = VSwitchModel(name=name_token)
                    vt = None
                    vh = None
                    for key, value in _parse_model_params(params):
                        if key == "ron":
                            model.Ron = value
                        elif key == "roff":
                            model.Roff = value
                        elif key == "von":
                            model.Von = value
                        elif key == "voff":
                            model.Voff = value
                        elif key == "vt":
                            vt = value
                        elif key == "vh":
                            vh = value
                    if vt is not None:
                        Vh = vh if vh is not None else 0
                        model.Von = vt + Vh / 2
                        model.Voff = vt - Vh / 2
                    ckt.models.vswitch[name_token.lower()] = model
                elif tl == "d":
                    dm = DiodeModel(name=name_token)
                    for key, value in _parse_model_params(params):
                        if key == "is":
                            dm.Is = value
                        elif key == "n":
                            dm.N = value
                    ckt.models.diode[name_token.lower()] = dm
                else:
                    ckt.skipped.append(line)
            else:
                ckt.skipped.append(line)
            continue

        tc = first[0].lower()
        name = first
        try:
            if tc == "r":
                n1 = ckt.nodes.get_or_create(_require(tokens, 1, "Resistor missing node"))
                n2 = ckt.nodes.get_or_create(_require(tokens, 2, "Resistor missing node"))
                val = parse_number_with_units(_require(tokens, 3, "Resistor missing value"))
                ckt.R.append(Resistor(name, n1, n2, val))
            elif tc == "c":
                n1 = ckt.nodes.get_or_create(_require(tokens, 1, "Capacitor missing node"))
                n2 = ckt.nodes.get_or_create(_require(tokens, 2, "Capacitor missing node"))
                val = parse_number_with_units(_require(tokens, 3, "Capacitor missing value"))
                ckt.C.append(Capacitor(name, n1, n2, val, 0.0))
            elif tc == "l":
                n1 = ckt.nodes.get_or_create(_require(tokens, 1, "Inductor missing node"))
                n2 = ckt.nodes.get_or_create(_require(tokens, 2, "Inductor missing node"))
                val = parse_number_with_units(_require(tokens, 3, "Inductor missing value"))
                ckt.L.append(Inductor(name, n1, n2, val, 0.0))
            elif tc == "v":
                n1 = ckt.nodes.get_or_create(_require(tokens, 1, "Voltage source missing node"))
                n2 = ckt.nodes.get_or_create(_require(tokens, 2, "Voltage source missing node"))
                vs = VoltageSource(name, n1, n2)
                i = 3
                if i < len(tokens) and not re.match(r"^[a-zA-Z]", tokens[i]):
                    vs.dc = parse_number_with_units(tokens[i])
                    i += 1
                while i < len(tokens):
                    key = tokens[i].lower()
                    if key == "dc":
                        vs.dc = parse_number_with_units(_require(tokens, i + 1, "DC value missing"))
                        i += 2
                    elif key == "ac":
                        vs.acMag = parse_number_with_units(_require(tokens, i + 1, "AC magnitude missing"))
                        phase = tokens[i + 2] if i + 2 < len(tokens) else None
                        if phase is not None and re.match(r"^[+-]?\d", phase):
                            vs.acPhaseDeg = parse_number_with_units(phase)
                            i += 3
                        else:
                            i += 2
                    elif key.startswith("pulse"):
                        arg = key if "(" in key else _require(tokens, i + 1, "PULSE() missing arguments")
                        if not arg or not re.search(r"\(.*\)", arg):
                            raise ValueError("Malformed PULSE() specification")
                        p = parse_pulse_args(arg)
                        vs.pulse, vs.pwl = p, None
                        vs.waveform = (lambda t, _p=p: pulse_value(_p, t))
                        i += 1 if "(" in key else 2
                    elif key.startswith("pwl"):
                        arg = key if "(" in key else _require(tokens, i + 1, "PWL() missing arguments")
                        if not arg or not re.search(r"\(.*\)", arg):
                            raise ValueError("Malformed PWL() specification")
                        pairs = parse_pwl_args(arg)
                        vs.pwl, vs.pulse = pairs, None
                        vs.waveform = (lambda t, _q=pairs: pwl_value(_q, t))
                        i += 1 if "(" in key else 2
                    else:
                        i += 1
                ckt.V.append(vs)
            elif tc == "i" and current_sources:
                n1 = ckt.nodes.get_or_create(_require(tokens, 1, "Current source missing node"))
                n2 = ckt.nodes.get_or_create(_require(tokens, 2, "Current source missing node"))
                cs = CurrentSource(name, n1, n2)
                i = 3
                if i < len(tokens) and not re.match(r"^[a-zA-Z]", tokens[i]):
                    cs.dc = parse_number_with_units(tokens[i])
                    i += 1
                while i < len(tokens):
                    key = tokens[i].lower()
                    if key == "dc":
                        cs.dc = parse_number_with_units(_require(tokens, i + 1, "DC value missing"))
                        i += 2
                    elif key == "ac":
                        cs.acMag = parse_number_with_units(_require(tokens, i + 1, "AC magnitude missing"))
                        phase = tokens[i + 2] if i + 2 < len(tokens) else None
                        if phase is not None and re.match(r"^[+-]?\d", phase):
                            cs.acPhaseDeg = parse_number_with_units(phase)
                            i += 3
                        else:
                            i += 2
                    else:
                        i += 1
                ckt.I.append(cs)
            elif tc == "s":
                n1 = ckt.nodes.get_or_create(_require(tokens, 1, "Switch missing node"))
                n2 = ckt.nodes.get_or_create(_require(tokens, 2, "Switch missing node"))
                cp = ckt.nodes.get_or_create(_require(tokens, 3, "Switch missing control node"))
                cn = ckt.nodes.get_or_create(_require(tokens, 4, "Switch missing control node"))
                mname = _require(tokens, 5, "Switch missing model")
                ckt.S.append(Switch(name, n1, n2, cp, cn, mname.lower()))
            elif tc == "d":
                if len(tokens) == 4:
                    np_ = ckt.nodes.get_or_create(_require(tokens, 1, "Diode missing node"))
                    nm = ckt.nodes.get_or_create(_require(tokens, 2, "Diode missing node"))
                    mname = _require(tokens, 3, "Diode missing model")
                    ckt.D.append(Diode(name, np_, nm, mname.lower()))
                else:
                    ckt.skipped.append(line)
            else:
                ckt.skipped.append(line)
        except ValueError as err:
            raise ValueError('Parse error on line: "%s"\n%s' % (line, err)) from None

    n_nodes = ckt.nodes.count() - 1
    for i, vs in enumerate(ckt.V):
        vs.index = n_nodes + i
    for sw in ckt.S:
        model = ckt.models.vswitch.get(sw.modelName)
        if model is None:
            raise ValueError("Unknown .model %s referenced by switch %s" % (sw.modelName, sw.name))
        sw.model = model
        sw.isOn = False
    for d in ckt.D:
        model = ckt.models.diode.get(d.modelName)
        if model is None:
            raise ValueError("Unknown .model %s referenced by diode %s" % (d.modelName, d.name))
        d.model = model
    return ckt


def logspace(f1: float, f2: float, points_per_decade: float) -> List[float]:
    """`.ac dec` frequency list (lib/utils/logspace.ts:3-15)."""
    if f1 <= 0 or f2 <= 0:
        raise ValueError(".ac frequencies must be > 0")
    if f2 < f1:
        f1, f2 = f2, f1
    decades = math.log10(f2 / f1)
    n = max(1, math.ceil(decades * points_per_decade))
    arr = [f1 * math.pow(10, i / points_per_decade) for i in range(n + 1)]
    if arr[-1] < f2 * (1 - EPS):
        arr.append(f2)
    return arr


def build_frequency_array(mode: str, N: float, f1: float, f2: float) -> List[float]:
    """simulateAC.ts:9-22."""
    if mode == "dec":
        return logspace(f1, f2, N)
    npts = max(2, N)
    step = (f2 - f1) / (npts - 1)
    return [f1 + i * step for i in range(npts)]


def compute_effective_time_step(dt_requested: float, tstop: float):
    """simulateTRAN.ts:14-19 — hazard H2: ulp-sensitive, host only."""
    dt_eff = dt_requested if dt_requested > EPS else max(tstop / 1000, EPS)
    steps = max(1, math.ceil(tstop / max(dt_eff, EPS)))
    dt = tstop / steps if steps > 0 else tstop
    return dt, steps
