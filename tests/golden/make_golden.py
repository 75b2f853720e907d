#!/usr/bin/env python3
"""Extract the reference's own known-answer vectors into small JSON fixtures.

Run in the build container, where /root/reference exists (it does not exist on
the GPU box, so tests only ever read the JSON written here):

    python tests/golden/make_golden.py

Sources (data only — expected outputs and the input netlists of the tests):
  tests/basics/basics01.test.ts:4-13,19-220      README RC low-pass, 201 rows
  tests/transient/*.test.ts                      input netlists
  tests/transient/__snapshots__/*.snap.svg       plotted spicey series

SVG decoding: plot rectangle x=100..1152, y=64..520; the axis
labels at y=520 / y=64 give (vmin, vmax), those at x=100 / x=1152 give the
time range in ms; each <path class="simulation-line"> is one series in legend
order, "M x y L x y ...".  Only the spicey series are kept (names without
"(ngspice)").
"""
import json
import os
import re
import sys

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def first_template_literal(path):
    src = open(path, encoding="utf-8").read()
    m = re.search(r"const\s+\w+\s*=\s*`(.*?)`", src, re.S)
    return m.group(1)


def basics01():
    path = os.path.join(REF, "tests/basics/basics01.test.ts")
    src = open(path, encoding="utf-8").read()
    netlist = first_template_literal(path)
    snap = re.search(r"toMatchInlineSnapshot\(`\s*\"(.*?)\"\s*`\)", src, re.S).group(1)
    lines = [ln.strip() for ln in snap.split("\n")]
    header, rows = lines[0], []
    for ln in lines[1:]:
        f, v1, v2 = ln.split(", ")
        m1, p1 = v1.split(",")
        m2, p2 = v2.split(",")
        rows.append([f, m1, p1, m2, p2])
    json.dump({"source": "tests/basics/basics01.test.ts:4-13,19-220", "netlist": netlist,
               "header": header, "rows": rows},
              open(os.path.join(OUT, "basics01_ac.json"), "w"), indent=0)
    print("basics01:", len(rows), "rows")


def axis_labels(svg, axis):
    out = []
    for m in re.finditer(r'<text class="axis-label axis-label-%s" x="([^"]+)" y="([^"]+)"[^>]*>([^<]*)</text>' % axis, svg):
        out.append((float(m.group(1)), float(m.group(2)), float(m.group(3))))
    return out


def svg_case(test_file, snap_file, out_name):
    netlist = first_template_literal(os.path.join(REF, "tests/transient", test_file))
    svg = open(os.path.join(REF, "tests/transient/__snapshots__", snap_file), encoding="utf-8").read()
    xl = axis_labels(svg, "x")
    yl = axis_labels(svg, "y")
    tmin = [v for (x, _, v) in xl if x == 100.0][0]
    tmax = [v for (x, _, v) in xl if x == 1152.0][0]
    vmin = [v for (_, y, v) in yl if y == 520.0][0]
    vmax = [v for (_, y, v) in yl if y == 64.0][0]
    names = re.findall(r'<text class="legend-label"[^>]*>([^<]*)</text>', svg)
    paths = re.findall(r'<path class="simulation-line" d="([^"]*)"', svg)
    assert len(names) == len(paths), (names, len(paths))
    series = {}
    for name, d in zip(names, paths):
        if "(ngspice)" in name:
            continue
        nums = re.findall(r"[ML]\s+(-?[\d.]+)\s+(-?[\d.]+)", d)
        series[name] = [[float(x), float(y)] for x, y in nums]
    json.dump({"source": "tests/transient/%s + __snapshots__/%s" % (test_file, snap_file),
               "netlist": netlist, "plot": {"x0": 100, "x1": 1152, "y0": 520, "y1": 64},
               "t_ms_range": [tmin, tmax], "v_range": [vmin, vmax], "series_px": series},
              open(os.path.join(OUT, out_name), "w"))
    print(out_name, {k: len(v) for k, v in series.items()}, (tmin, tmax), (vmin, vmax))


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    basics01()
    svg_case("transient01.test.ts", "transient01-rc-pulse-comparison.snap.svg", "transient01_rc_pulse.json")
    svg_case("two-probes.test.ts", "two-probes-two-probes-graph.snap.svg", "two_probes.json")
    svg_case("switch-vt-vh.test.ts", "switch-vt-vh-switch-vt-vh-graph.snap.svg", "switch_vt_vh.json")
    svg_case("vswitch-pwl.test.ts", "vswitch-pwl-vswitch-pwl-control.snap.svg", "vswitch_pwl.json")
    svg_case("boost-converter-probe.test.ts", "boost-converter-probe-boost-converter-probe.snap.svg",
             "boost_converter_probe.json")
    for name, f in (("diode_switch", "diode-switch.test.ts"), ("case_insensitive_nodes", "case-insensitive-nodes.test.ts")):
        json.dump({"source": "tests/transient/" + f,
                   "netlist": first_template_literal(os.path.join(REF, "tests/transient", f))},
                  open(os.path.join(OUT, name + ".json"), "w"))


if __name__ == "__main__":
    main()
