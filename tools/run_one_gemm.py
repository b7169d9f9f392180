#!/usr/bin/env python
"""Runs a single ctta_gemm problem a few times (for ncu captures).
   python tools/run_one_gemm.py conv1d --c 64 --taps 3 --rows 81936 --batch 64 --kind c1
"""
import argparse, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from consistencytta_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["conv1d", "conv2d", "linear"])
ap.add_argument("--c", type=int, default=64)
ap.add_argument("--n", type=int, default=None)
ap.add_argument("--taps", type=int, default=3)
ap.add_argument("--dil", type=int, default=1)
ap.add_argument("--rows", type=int, default=81936)
ap.add_argument("--h", type=int, default=256)
ap.add_argument("--w", type=int, default=16)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--kind", default="c1", choices=["c1", "c2", "c2h", "f32res", "f16", "geglu"])
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--stats", action="store_true", help="also emit the fused GroupNorm moments of the result (32 groups)")
ap.add_argument("--seconds", type=float, default=0.0, help="sustained mode: repeat for this long first (power-capped clocks)")
a = ap.parse_args()
dev = "cuda"
n = a.n or a.c
DT = ops.OPERAND_DTYPE
if a.what == "conv1d":
    x = torch.randn(a.batch, a.rows, a.c, device=dev).to(DT)
    pw = ops.pack_conv1d(torch.randn(n, a.c, a.taps, device=dev) / math.sqrt(a.c * a.taps), torch.randn(n, device=dev), a.dil)
    shape = (a.batch, a.rows, n)
    fn = ops.conv1d
elif a.what == "conv2d":
    x = torch.randn(a.batch, a.h, a.w, a.c, device=dev).to(DT)
    pw = ops.pack_conv2d(torch.randn(n, a.c, 3, 3, device=dev) / math.sqrt(a.c * 9), torch.randn(n, device=dev))
    shape = (a.batch, a.h, a.w, n)
    fn = ops.conv2d
else:
    x = torch.randn(a.rows, a.c, device=dev).to(DT)
    pw = ops.pack_linear(torch.randn(n, a.c, device=dev) / math.sqrt(a.c), torch.randn(n, device=dev))
    shape = (a.rows, n)
    fn = ops.linear
kw = {}
if a.kind == "c1":
    kw = dict(out2=torch.empty(shape, device=dev, dtype=DT), act2=ops.ACT_LRELU, act2_slope=0.1)
elif a.kind == "c2":
    kw = dict(out=torch.empty(shape, device=dev), residual=torch.randn(shape, device=dev),
              out2=torch.empty(shape, device=dev, dtype=DT), act2=ops.ACT_LRELU, act2_slope=0.1)
elif a.kind == "c2h":   # HiFi-GAN c2: 16-bit LeakyReLU'ed residual -> LeakyReLU'ed 16-bit operand
    kw = dict(residual=torch.randn(shape, device=dev).to(DT), res_neg_scale=10.0,
              out2=torch.empty(shape, device=dev, dtype=DT), act2=ops.ACT_LRELU, act2_slope=0.1)
elif a.kind == "geglu":   # UNet feed-forward: interleaved (value, gate) columns -> n / 2 16-bit outputs
    kw = dict(out=torch.empty(shape[:-1] + (n // 2,), device=dev, dtype=DT), act=ops.ACT_GEGLU)
elif a.kind == "f32res":
    kw = dict(out=torch.empty(shape, device=dev), residual=torch.randn(shape, device=dev))
else:
    kw = dict(out=torch.empty(shape, device=dev, dtype=DT))
if a.stats:
    kw.update(stats=torch.empty(a.batch, 32, 2, device=dev), stats_groups=32)
    if a.what == "linear":
        kw.update(stats_rows_per_img=a.rows // a.batch)
for _ in range(a.iters):
    fn(x, pw, **kw)
torch.cuda.synchronize()
if a.seconds > 0:   # bring the chip to its sustained (power-capped) clocks with the same kernel before timing
    import time
    t0 = time.time()
    while time.time() - t0 < a.seconds:
        for _ in range(50):
            fn(x, pw, **kw)
        torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    fn(x, pw, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
rows = x.numel() // a.c
print("%s c=%d n=%d taps=%d rows=%d kind=%s%s: %.3f ms  %.1f TFLOP/s" % (a.what, a.c, n, pw.ntaps, rows, a.kind, "+stats" if a.stats else "", ms,
      2.0 * rows * n * pw.ntaps * a.c / ms / 1e9))
