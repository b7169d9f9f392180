import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from consistencytta_b200 import ops
DT = ops.OPERAND_DTYPE
torch.manual_seed(0)
def r16(x): return x.to(DT).float()
def run(c, k, dil, t, bsz, mode):
    slope = 0.1
    x = torch.randn(bsz, t, c, device="cuda")
    lx = F.leaky_relu(x, slope).to(DT)
    w1 = r16(torch.randn(c, c, k, device="cuda") / math.sqrt(k * c))
    w2 = r16(torch.randn(c, c, k, device="cuda") / math.sqrt(k * c))
    b1, b2 = torch.randn(c, device="cuda") * 0.1, torch.randn(c, device="cuda") * 0.1
    if mode == "w2zero": w2.zero_()
    if mode == "w1zero": w1.zero_()
    if mode == "center":   # only the centre tap of both convs
        m = torch.zeros(k, device="cuda"); m[k // 2] = 1; w1 *= m; w2 *= m
    pw1, pw2 = ops.pack_conv1d(w1, b1, dilation=dil), ops.pack_conv1d(w2, b2, dilation=1)
    out = ops.resblock_pair(lx, pw1, pw2, slope).float()
    lxf = lx.float()
    x_rec = torch.where(lxf < 0, lxf / slope, lxf)
    h = F.conv1d(lxf.permute(0, 2, 1), w1, b1, dilation=dil, padding=(k * dil - dil) // 2)
    h16 = r16(F.leaky_relu(h, slope))
    y = x_rec.permute(0, 2, 1) + F.conv1d(h16, w2, b2, padding=(k - 1) // 2)
    ref = F.leaky_relu(y, slope).permute(0, 2, 1)
    d = (out - ref)
    e = (d.norm() / ref.norm()).item()
    # per-row error profile
    row_err = d.pow(2).sum(-1).sqrt()[0]
    bad = (row_err > 1e-2 * ref.pow(2).sum(-1).sqrt()[0].clamp_min(1e-3)).nonzero().flatten()
    ch_err = d.pow(2).sum((0, 1)).sqrt()
    print("c=%d k=%d dil=%d t=%d mode=%-7s rel=%.3e bad rows %d/%d first %s  ch_err[:8]=%s" % (
        c, k, dil, t, mode, e, bad.numel(), t, bad[:12].tolist(), [round(v, 3) for v in ch_err[:8].tolist()]))
    if e > 1e-2 and mode == "w2zero":
        print(" out[0,0,:8]", out[0, 0, :8].tolist()); print(" ref[0,0,:8]", ref[0, 0, :8].tolist())
        print(" out[0,1,:8]", out[0, 1, :8].tolist()); print(" ref[0,1,:8]", ref[0, 1, :8].tolist())
for mode in ("w2zero", "w1zero", "center", "full"):
    run(64, 3, 1, 600, 1, mode)
for mode in ("w2zero", "w1zero", "center", "full"):
    run(32, 3, 1, 600, 1, mode)
