#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: stall-reason totals and the hottest SASS lines."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = rows[2:]
    tot = sum(int(r[col["# Samples"]] or 0) for r in body)
    print("kernel:", rows[0][1])
    print("total samples", tot, " sass lines", len(body))
    agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
    for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print("  %-24s %8d %5.1f%%" % (s, v, 100.0 * v / max(1, tot)))
    print("top lines:")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        n = int(r[col["# Samples"]] or 0)
        dom = max(stalls, key=lambda s: int(r[col[s]] or 0))
        print("  %5d %5.1f%% #%-5d ex=%-9s %-22s %s" % (n, 100.0 * n / max(1, tot), i, r[col["Instructions Executed"]],
                                                  dom, r[col["Source"]].strip()[:90]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
