#!/usr/bin/env python
"""Times one fused HiFi-GAN ResBlock pair (ctta_resblock_pair) against the two-launch ctta_gemm path it replaces.
   python tools/run_one_pair.py --c 32 --taps 11 --dil 5 --t 163872 --batch 64 [--seconds 1]"""
import argparse
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from consistencytta_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=32)
ap.add_argument("--taps", type=int, default=11)
ap.add_argument("--dil", type=int, default=1)
ap.add_argument("--t", type=int, default=163872)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--seconds", type=float, default=0.0)
a = ap.parse_args()
DT = ops.OPERAND_DTYPE
lx = torch.randn(a.batch, a.t, a.c, device="cuda").to(DT)
w1 = torch.randn(a.c, a.c, a.taps, device="cuda") / math.sqrt(a.c * a.taps)
w2 = torch.randn(a.c, a.c, a.taps, device="cuda") / math.sqrt(a.c * a.taps)
b = torch.randn(a.c, device="cuda") * 0.1
pw1, pw2 = ops.pack_conv1d(w1, b, dilation=a.dil), ops.pack_conv1d(w2, b, dilation=1)
tmp, out = torch.empty_like(lx), torch.empty_like(lx)


def fused():
    ops.resblock_pair(lx, pw1, pw2, 0.1, out=out)


def unfused():
    ops.conv1d(lx, pw1, out2=tmp, act2=ops.ACT_LRELU, act2_slope=0.1)
    ops.conv1d(tmp, pw2, residual=lx, res_neg_scale=10.0, out2=out, act2=ops.ACT_LRELU, act2_slope=0.1)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if a.seconds > 0:
        t0 = time.time()
        while time.time() - t0 < a.seconds:
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


flops = 2 * 2.0 * a.batch * a.t * a.c * a.c * a.taps
gbytes = 2 * 2.0 * a.batch * a.t * a.c / 1e9
res = {}
if ops.resblock_pair_supported(a.c, a.taps, a.dil):
    res["fused"] = timed(fused)
res["unfused"] = timed(unfused)
for k, ms in res.items():
    print("pair c=%d taps=%d dil=%d rows=%d %-8s %.3f ms  %.0f TFLOP/s  %.0f GB/s of the fused path's algorithmic bytes"
          % (a.c, a.taps, a.dil, a.batch * a.t, k, ms, flops / ms / 1e9, gbytes / ms * 1e3))
