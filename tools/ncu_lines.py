#!/usr/bin/env python
"""Stall samples of an .ncu-rep aggregated per CUDA source line (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(io.StringIO(out)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


for k, hi in enumerate(hdr):
    end = hdr[k + 1] - 2 if k + 1 < len(hdr) else len(rows)
    head = rows[hi]
    si, ie = head.index("# Samples"), head.index("Instructions Executed")
    stalls = [i for i, h in enumerate(head) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[hi + 1:end] if len(r) > si and r[0].isdigit() and r[2] == "-"]
    tot = sum(num(r[si]) for r in body)
    if not tot:
        continue
    print("== section %d: %d samples, %d warp instructions" % (k, tot, sum(num(r[ie]) for r in body)))
    for r in sorted(body, key=lambda r: -num(r[si]))[:n]:
        top = sorted(((num(r[j]), head[j][6:]) for j in stalls), reverse=True)[:2]
        print("%5s %6d %5.1f%% inst %9s  %-28s %s" % (r[0], num(r[si]), 100.0 * num(r[si]) / tot, r[ie],
                                                      " ".join("%s=%d" % (b, a) for a, b in top if a), r[1].strip()[:100]))
