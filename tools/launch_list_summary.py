#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: python tools/launch_list_summary.py file.csv[.gz]"""
import collections, csv, gzip, re, sys
path = sys.argv[1]
op = gzip.open if path.endswith(".gz") else open
with op(path, "rt", newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
head = next(rd)
ki, ui, vi = head.index("Kernel Name"), head.index("Metric Unit"), head.index("Metric Value")
U = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}
agg = collections.OrderedDict()
for r in rd:
    name = r[ki]
    short = re.sub(r"^void ", "", name)
    short = re.sub(r"<.*", "", short.split("(")[0])
    ours = "ctta" in name or "tap_sum" in name or "flash_attn" in name or "resblock_pair" in name
    key = (ours, short)
    v = float(r[vi].replace(",", "")) * U.get(r[ui], 1e-6)
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
ours = [(k[1], v) for k, v in agg.items() if k[0]]
other = [(k[1], v) for k, v in agg.items() if not k[0]]
tot = sum(v[1] for _, v in ours)
print("libctta kernels over every launch of the process (%d launches, %.1f ms; other (torch fill / copy) kernels: %d launches, %.1f ms):"
      % (sum(v[0] for _, v in ours), tot, sum(v[0] for _, v in other), sum(v[1] for _, v in other)))
for n, (c, ms) in sorted(ours, key=lambda kv: -kv[1][1]):
    print("%-40s %5d launches %10.3f ms %5.1f%%" % (n[:40], c, ms, 100 * ms / tot))
print("other kernels:")
for n, (c, ms) in sorted(other, key=lambda kv: -kv[1][1])[:8]:
    print("%-60s %5d launches %10.3f ms" % (n[:60], c, ms))
