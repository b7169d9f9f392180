#!/usr/bin/env python
"""Counts the Blackwell-only SASS opcodes per object file of libctta (cuobjdump -sass on consistencytta_b200/build/*.o):
UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG / UTMASTG / UTMAREDG (TMA loads / stores / reduce-adds), LDTM / STTM
(tcgen05.ld / st), UTCBAR (tcgen05.commit), plus the legacy HMMA (mma.sync) for contrast.

    python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDTM", "STTM", "SYNCS", "HMMA", "MUFU.EX2",
       "FFMA2", "FADD2", "FMUL2", "FMNMX3"]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "consistencytta_b200", "build", "*.o")))
    if not objs:
        sys.exit("no objects under consistencytta_b200/build: run python -m consistencytta_b200.build --force")
    print("# cuobjdump -sass opcode counts per object (sm_100a); columns: " + " ".join(OPS))
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        kernels = collections.OrderedDict()
        cur = None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                kernels[cur] = collections.Counter()
                continue
            if cur is None:
                continue
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                kernels[cur]["_total"] += 1
                for k in OPS:
                    if op == k or op.startswith(k + "."):
                        kernels[cur][k] += 1
        tot = collections.Counter()
        for c in kernels.values():
            tot.update(c)
        print("\n%s: %d kernels, %d SASS instructions" % (os.path.basename(o), len(kernels), tot["_total"]))
        print("  TOTAL  " + "  ".join("%s=%d" % (k, tot[k]) for k in OPS if tot[k]))
        for name, c in kernels.items():
            hot = "  ".join("%s=%d" % (k, c[k]) for k in OPS if c[k])
            short = subprocess.run(["c++filt", "-p", name], capture_output=True, text=True).stdout.strip() or name
            print("  %-90s %6d  %s" % (short[:90], c["_total"], hot))


if __name__ == "__main__":
    main()
