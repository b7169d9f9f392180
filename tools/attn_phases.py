#!/usr/bin/env python
"""Phase timestamps of one softmax warp of the tcgen05 attention kernel (CTA 0, key tiles 8..15).
   CTTA_ATTN_DEBUG=1 python tools/attn_phases.py"""
import ctypes as C
import os
import sys
os.environ["CTTA_ATTN_DEBUG"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from consistencytta_b200 import _lib, ops  # noqa: E402

b, h, l = 16, 5, 4096
qkv = torch.randn(b, l, 3, h, 64, device="cuda").to(ops.OPERAND_DTYPE)
o = torch.empty(b, l, h, 64, device="cuda", dtype=ops.OPERAND_DTYPE)
for _ in range(3):
    ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], 51 ** -0.5, out=o)
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
fn = _lib.lib().ctta_attention_debug
fn.argtypes = [C.POINTER(C.c_longlong)]
fn.restype = C.c_int
assert fn(buf) == 0
names = ["wait S", "ld S + s_free", "max/rescale", "exp+sum", "stagger/arrive", "wait PV / pack+STS begin", "pack+STS+arrive"]
t = [[buf[j * 8 + k] for k in range(7)] for j in range(8)]
print("per key tile (cycles): " + " | ".join(names[:4] + ["(wait pv)", "pack+STS+fence+arrive", "loop back"]))
for j in range(8):
    d = [t[j][k + 1] - t[j][k] for k in range(6)]
    nxt = (t[j + 1][0] - t[j][6]) if j < 7 else 0
    print("tile %2d: " % (j + 8) + " ".join("%6d" % x for x in d) + " %6d   total %6d" % (nxt, (t[j + 1][0] - t[j][0]) if j < 7 else 0))
