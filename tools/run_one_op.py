#!/usr/bin/env python
"""Runs one bandwidth-bound / attention libctta kernel a few times and prints its CUDA-event time (for ncu captures
and roofline bookkeeping).

   python tools/run_one_op.py attention --b 64 --heads 5 --lq 4096 --lk 4096
   python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128 [--in16]
   python tools/run_one_op.py gn_stats --n 64 --h 1024 --w 64 --c 128 [--in16]
   python tools/run_one_op.py layernorm --rows 262144 --d 255
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from consistencytta_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["attention", "gn_apply", "gn_stats", "layernorm", "tap_sum", "mrf_combine"])
ap.add_argument("--b", type=int, default=64)
ap.add_argument("--heads", type=int, default=5)
ap.add_argument("--lq", type=int, default=4096)
ap.add_argument("--lk", type=int, default=4096)
ap.add_argument("--n", type=int, default=64)
ap.add_argument("--h", type=int, default=1024)
ap.add_argument("--w", type=int, default=64)
ap.add_argument("--c", type=int, default=128)
ap.add_argument("--rows", type=int, default=262144)
ap.add_argument("--d", type=int, default=255)
ap.add_argument("--in16", action="store_true")
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()
dev = "cuda"
DT = ops.OPERAND_DTYPE

if a.what == "attention":
    qkv = torch.randn(a.b, a.lq, 3, a.heads, 64, device=dev).to(DT)
    if a.lk != a.lq:
        kv = torch.randn(a.b, a.lk, 2, a.heads, 64, device=dev).to(DT)
        q, k, v = qkv[:, :, 0], kv[:, :, 0], kv[:, :, 1]
    else:
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    o = torch.empty(a.b, a.lq, a.heads, 64, device=dev, dtype=DT)
    fn = lambda: ops.attention(q, k, v, 51 ** -0.5, out=o)
    work = 4.0 * a.b * a.heads * a.lq * a.lk * 64
    unit, scale = "TFLOP/s (d padded to 64)", 1e9
elif a.what in ("gn_apply", "gn_stats"):
    x = torch.randn(a.n, a.h, a.w, a.c, device=dev)
    if a.in16:
        x = x.to(DT)
    st = ops.groupnorm_stats(x, 32)
    g, bta = torch.ones(a.c, device=dev), torch.zeros(a.c, device=dev)
    y = torch.empty(a.n, a.h, a.w, a.c, device=dev, dtype=DT)
    if a.what == "gn_apply":
        fn = lambda: ops.groupnorm_apply(x, 32, st, g, bta, eps=1e-6, act=ops.ACT_SILU, out=y)
        work = x.numel() * (x.element_size() + 2)
    else:
        fn = lambda: ops.groupnorm_stats(x, 32, stats=st)
        work = x.numel() * x.element_size()
    unit, scale = "GB/s", 1e6
elif a.what == "tap_sum":   # VAE conv_out tail: 9 taps of a [n, h, w, 16] fp32 partial-product tensor
    sc = ops.pack_single_channel_conv(torch.randn(1, a.c, 3, 3, device=dev), torch.randn(1, device=dev))
    z = torch.randn(a.n * a.h * a.w, sc.n_pad, device=dev)
    y = torch.empty(a.n * a.h * a.w, device=dev)
    y16 = torch.empty(a.n * a.h * a.w, device=dev, dtype=DT)
    fn = lambda: ops.tap_sum(z, sc, a.n, a.h, a.w, ops.ACT_NONE, y, y16)
    work = z.shape[0] * (9 * 4 + 4 + 2)     # the 9 taps that are summed + both outputs
    unit, scale = "GB/s (algorithmic)", 1e6
elif a.what == "mrf_combine":   # HiFi-GAN stage sum: 3 x 16-bit in, 16-bit out
    xs = [torch.randn(a.n, a.rows, a.c, device=dev).to(DT) for _ in range(3)]
    y = torch.empty_like(xs[0])
    fn = lambda: ops.mrf_combine(xs, 0.1, 1.0 / 3, 0.1, out=y)
    work = xs[0].numel() * 8
    unit, scale = "GB/s", 1e6
else:
    ld = ops.round_up(a.d, 64)
    x = torch.randn(a.rows, ld, device=dev)
    g, bta = torch.ones(ld, device=dev), torch.zeros(ld, device=dev)
    y = torch.empty(a.rows, ld, device=dev, dtype=DT)
    fn = lambda: ops.layernorm(x, a.d, g, bta, 1e-5, out=y)
    work = x.numel() * 6
    unit, scale = "GB/s", 1e6

for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print("%s %s: %.3f ms  %.1f %s" % (a.what, vars(a), ms, work / ms / scale, unit))
