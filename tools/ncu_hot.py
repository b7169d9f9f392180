#!/usr/bin/env python
"""Top stall locations (SASS) of an .ncu-rep: python tools/ncu_hot.py file.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
head = rows[hi]
body = rows[hi + 1:]
si = head.index("# Samples")
stalls = [i for i, h in enumerate(head) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
order = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:n]
print("total samples", tot)
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[j] or 0), head[j]) for j in stalls), reverse=True)[:2]
    print("%5d %6s %5.1f%%  %-70s %s" % (i, r[si], 100.0 * int(r[si]) / max(tot, 1), r[1].strip()[:70],
                                        " ".join("%s=%d" % (b[6:], a) for a, b in top if a)))
