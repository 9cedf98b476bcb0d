#!/usr/bin/env python
"""Hot-spot view of an ncu source page (`ncu -i rep --page source --csv`): samples per 200-instruction
region and the top stalled SASS instructions.  Usage: python scripts/ncu_source_hot.py report.ncu-rep [kernel-index]"""
import collections
import csv
import io
import subprocess
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = heads[which]
    end = heads[which + 1] - 1 if which + 1 < len(heads) else len(rows)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(num(r[ix["# Samples"]]) for r in data)
    print(rows[hi - 1][:2], "instructions", len(data), "samples", tot)
    keys = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_no_inst", "stall_branch_resolving", "stall_barrier", "stall_math", "stall_selected"]
    b = collections.OrderedDict()
    for n, r in enumerate(data):
        v = b.setdefault(n // 200, [0] * (len(keys) + 2))
        v[0] += num(r[ix["# Samples"]])
        v[1] += num(r[ix["Instructions Executed"]])
        for j, k in enumerate(keys):
            v[j + 2] += num(r[ix[k]])
    print("region  samples  warp-instr ", keys)
    for k, v in b.items():
        print(f"{k * 200:6d} {v[0]:7d} {v[1]:9d}  {v[2:]}   {data[k * 200][ix['Source']][:50]}")
    for r in sorted(data, key=lambda r: -num(r[ix["# Samples"]]))[:25]:
        print(num(r[ix["# Samples"]]), {k[6:]: num(r[ix[k]]) for k in keys if num(r[ix[k]])}, r[ix["Source"]][:80])


if __name__ == "__main__":
    main()
