#!/usr/bin/env python
"""Condense an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <command>`) into one row per
(kernel, grid size): launches, mean / min / max duration, summed time and share of the total.  Usage: launch_summary.py X.csv"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = []
    with open(sys.argv[1], newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        if len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(d["Metric Value"].replace(",", ""))
        unit = d.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        name = re.sub(r"\(.*", "", d["Kernel Name"])
        rows.append((name, d.get("Grid Size", ""), ns))
    agg = OrderedDict()
    for name, grid, ns in rows:
        a = agg.setdefault((name, grid), [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += ns; a[2] = min(a[2], ns); a[3] = max(a[3], ns)
    total = sum(a[1] for a in agg.values()) or 1.0
    print(f"# launch list summary: {sys.argv[1]} ({len(rows)} launches, {total / 1e6:.2f} ms of kernel time; per-launch times are cold-cache and serialised)")
    print()
    print("| kernel | grid | launches | mean us | min us | max us | sum ms | share |")
    print("|---|---|---|---|---|---|---|---|")
    for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {grid} | {a[0]} | {a[1] / a[0] / 1e3:.2f} | {a[2] / 1e3:.2f} | {a[3] / 1e3:.2f} | {a[1] / 1e6:.3f} | {a[1] / total:.3f} |")


if __name__ == "__main__":
    main()
