#!/usr/bin/env python
"""Repeat-run determinism + parity stress of the fused ResBlock-pair kernel: every supported (C, taps, dilation) family at
HiFi-GAN-like sizes, REPS launches each, every launch compared with the two-launch ctta_gemm path bit-for-bit against
the first launch and within 1e-3 rel-L2 against the unfused result."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from consistencytta_b200 import ops
DT = ops.OPERAND_DTYPE
reps = int(os.environ.get("REPS", "200"))
tot_bad = 0
for c, k, dil, t, bsz in [(64, 3, 1, 81936, 3), (64, 3, 5, 81936, 2), (64, 7, 3, 81936, 2), (32, 3, 1, 163872, 2), (32, 7, 5, 163872, 2),
                          (32, 11, 3, 163872, 1), (32, 11, 1, 20000, 12), (64, 7, 1, 5000, 40)]:
    if not ops.resblock_pair_supported(c, k, dil, t):
        print("unsupported", c, k, dil); continue
    torch.manual_seed(c + k + dil)
    lx = F.leaky_relu(torch.randn(bsz, t, c, device="cuda"), 0.1).to(DT)
    w1 = torch.randn(c, c, k, device="cuda") / math.sqrt(k * c)
    w2 = torch.randn(c, c, k, device="cuda") / math.sqrt(k * c)
    b1 = torch.randn(c, device="cuda") * 0.1
    pw1, pw2 = ops.pack_conv1d(w1, b1, dilation=dil), ops.pack_conv1d(w2, b1, dilation=1)
    tmp, two = torch.empty_like(lx), torch.empty_like(lx)
    ops.conv1d(lx, pw1, out2=tmp, act2=ops.ACT_LRELU, act2_slope=0.1)
    ops.conv1d(tmp, pw2, residual=lx, res_neg_scale=10.0, out2=two, act2=ops.ACT_LRELU, act2_slope=0.1)
    first = ops.resblock_pair(lx, pw1, pw2, 0.1).clone()
    rel = ((first.float() - two.float()).norm() / two.float().norm()).item()
    bad = 0
    for _ in range(reps):
        out = ops.resblock_pair(lx, pw1, pw2, 0.1)
        bad += int(not torch.equal(out, first))
    tot_bad += bad + (rel > 1e-3)
    print("c=%d k=%d dil=%d t=%d b=%d: rel vs unfused %.2e, %d of %d launches differ from the first" % (c, k, dil, t, bsz, rel, bad, reps))
print("TOTAL bad:", tot_bad)
sys.exit(1 if tot_bad else 0)
