#!/usr/bin/env python
"""DRAM traffic of the dominant kernel over ONE step, from an ncu metrics pass of the eager step:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_tc_kernel --csv \
        --log-file gpurun_out/gemm_dram.csv python tools/profile_layers.py --batch 64
    python tools/gemm_traffic.py gpurun_out/gemm_dram.csv 64 > profiles/gemm_dram_traffic.json

tools/profile_layers.py runs two eager steps; the second (steady state: weights packed, buffers allocated) is used.
bench.py reads the JSON for `roofline.traffic` (average DRAM bytes per launch of the kernel)."""
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    path, batch = sys.argv[1], int(sys.argv[2])
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    head = next(rd)
    ki, mi, ui, vi, ii = (head.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = {}
    for r in rd:
        if "gemm_tc_kernel" not in r[ki]:
            continue
        v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        per.setdefault(int(r[ii]), {})[r[mi]] = v
    ids = sorted(per)
    n = len(ids) // 2            # two steps were profiled
    last = ids[len(ids) - n:]
    rd_b = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in last)
    wr_b = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in last)
    print(json.dumps({"kernel": "gemm_tc_kernel", "batch": batch, "launches_per_step": n,
                      "dram_read_bytes_per_step": rd_b, "dram_write_bytes_per_step": wr_b,
                      "dram_bytes_per_launch": (rd_b + wr_b) / max(n, 1),
                      "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum over every gemm_tc_kernel launch of one "
                                "eager step (tools/gemm_traffic.py)"}, indent=1))


if __name__ == "__main__":
    main()
