#!/usr/bin/env python
"""Phase timeline of CTA 0 of the fused ResBlock-pair kernel (a library built with -DRBP_TRACE=1, see tools/exp_pair_trace.sh):
   CTTA_LIB=consistencytta_b200/libctta_trace.so python tools/trace_pair.py --c 64 --taps 3 --dil 1 --t 81920 --batch 64"""
import argparse, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=64)
ap.add_argument("--taps", type=int, default=3)
ap.add_argument("--dil", type=int, default=1)
ap.add_argument("--t", type=int, default=81920)
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
trace = torch.zeros(48 * 20, dtype=torch.int64, device="cuda")
os.environ["CTTA_RBP_TRACE_PTR"] = hex(trace.data_ptr())
from consistencytta_b200 import ops
DT = ops.OPERAND_DTYPE
lx = torch.randn(a.batch, a.t, a.c, device="cuda").to(DT)
w1 = torch.randn(a.c, a.c, a.taps, device="cuda") / math.sqrt(a.c * a.taps)
w2 = torch.randn(a.c, a.c, a.taps, device="cuda") / math.sqrt(a.c * a.taps)
b = torch.randn(a.c, device="cuda") * 0.1
pw1, pw2 = ops.pack_conv1d(w1, b, dilation=a.dil), ops.pack_conv1d(w2, b, dilation=1)
out = torch.empty_like(lx)
for _ in range(3):
    ops.resblock_pair(lx, pw1, pw2, 0.1, out=out)
torch.cuda.synchronize()
tr = trace.cpu().view(48, 20).numpy()
names = ["tma_x", "c1_go", "c1_end", "c2_go", "c2_end", "e1_acc1", "e1_ld", "e1_hfree", "e1_hfull", "e2_top", "e2_flush",
         "e2_acc2", "e2_ld", "e2_math", "e2_end"]
t0 = tr[tr > 0].min()
print("c=%d taps=%d dil=%d  (cycles since the first stamp; tile rows)" % (a.c, a.taps, a.dil))
print("tile " + " ".join("%8s" % n for n in names))
for n in range(10, 16):
    print("%4d " % n + " ".join("%8d" % (tr[n, e] - t0) for e in range(15)))
per = (tr[40, 14] - tr[10, 14]) / 30.0
print("period per tile (tiles 10..40): %.0f cycles" % per)
for e0, e1, label in [(1, 2, "c1 issue"), (3, 4, "c2 issue"), (5, 6, "E1 tmem ld"), (6, 7, "E1 wait h_free"), (7, 15, "E1 math+st"),
                      (15, 8, "E1 fence+arrive"), (9, 17, "E2 ldg issue"), (17, 10, "E2 flush"), (10, 11, "E2 wait acc2"),
                      (11, 12, "E2 tmem ld"), (12, 16, "E2 math+st"), (16, 13, "E2 fence"), (13, 14, "E2 store issue")]:
    d = (tr[10:40, e1] - tr[10:40, e0]).mean()
    print("  %-20s %7.0f" % (label, d))
print("  E1 total (acc1 passed -> h_full) %7.0f" % (tr[10:40, 8] - tr[10:40, 5]).mean())
print("  E2 total (top -> end)            %7.0f" % (tr[10:40, 14] - tr[10:40, 9]).mean())
print("  h_full -> c2_go                  %7.0f" % (tr[10:40, 3] - tr[10:40, 8]).mean())
print("  c2_end -> e2_acc2                %7.0f" % (tr[10:40, 11] - tr[10:40, 4]).mean())
print("  c1_end -> e1_acc1                %7.0f" % (tr[10:40, 5] - tr[10:40, 2]).mean())
