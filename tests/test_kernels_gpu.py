"""Per-kernel parity: every libctta kernel (called through the C-ABI via ctypes) against the same op in plain
PyTorch fp32 on identical inputs.  Operands are pre-rounded to fp16 so the comparison isolates the kernel."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from consistencytta_b200 import _lib, ops  # noqa: E402

DEV = "cuda"
DT = ops.OPERAND_DTYPE


def rel(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def r16(x):
    return x.to(DT).float()


@pytest.mark.parametrize("m,k,n", [(300, 256, 320), (128, 64, 16), (1000, 1024, 8), (40000, 256, 256), (77, 320, 1280)])
def test_linear(m, k, n):
    torch.manual_seed(0)
    a = r16(torch.randn(m, k, device=DEV))
    w = r16(torch.randn(n, k, device=DEV) / math.sqrt(k))
    b = torch.randn(n, device=DEV)
    res = torch.randn(m, n, device=DEV)
    pw = ops.pack_linear(w, b)
    out = torch.empty(m, n, device=DEV)
    ops.linear(a.to(DT), pw, out=out, residual=res)
    ref = a @ w.t() + b + res
    assert rel(out, ref) < 2e-5
    # 16-bit output, no residual, padded K (k_pad_to) and SiLU second output
    out16 = torch.empty(m, n, device=DEV, dtype=DT)
    ops.linear(a.to(DT), pw, out=out16)
    assert rel(out16, a @ w.t() + b) < 1e-3


@pytest.mark.parametrize("m,k,d", [(500, 256, 1024), (300, 128, 48), (1000, 320, 640)])
def test_linear_geglu(m, k, d):
    """GEGLU epilogue: 64-output (128-byte) store rows when the N tile is a multiple of 128, 16-output rows otherwise."""
    torch.manual_seed(1)
    a = r16(torch.randn(m, k, device=DEV))
    w = r16(torch.randn(2 * d, k, device=DEV) / math.sqrt(k))
    b = torch.randn(2 * d, device=DEV)
    # interleave (value, gate)
    wi = torch.stack([w[:d], w[d:]], dim=1).reshape(2 * d, k)
    bi = torch.stack([b[:d], b[d:]], dim=1).reshape(2 * d)
    pw = ops.pack_linear(wi, bi)
    out = torch.empty(m, d, device=DEV, dtype=DT)
    ops.linear(a.to(DT), pw, out=out, act=ops.ACT_GEGLU)
    h = a @ w.t() + b
    ref = h[:, :d] * F.gelu(h[:, d:])
    assert rel(out, ref) < 1e-3


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 16, 64, 128), (3, 256, 16, 8, 256), (1, 64, 4, 192, 8),
                                            (5, 32, 2, 128, 64), (1, 128, 64, 128, 1), (2, 16, 32, 320, 512),
                                            # large enough for M super-tiles in stream mode (>= 2 x 148 tiles of 256 pixels)
                                            (24, 256, 16, 64, 96), (2, 1024, 64, 128, 128), (40, 128, 8, 72, 256),
                                            (70, 64, 4, 128, 64), (3, 512, 32, 256, 512)])
def test_conv2d_3x3(n, h, w, cin, cout):
    torch.manual_seed(2)
    x = r16(torch.randn(n, cin, h, w, device=DEV))
    wt = r16(torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(9 * cin))
    b = torch.randn(cout, device=DEV)
    temb = torch.randn(n, cout, device=DEV)
    ref = F.conv2d(x, wt, b, padding=1) + temb[:, :, None, None]
    pw = ops.pack_conv2d(wt, b)
    cp = ops.round_up(cin, 8)
    a = torch.zeros(n, h, w, cp, device=DEV, dtype=DT)
    a[..., :cin] = x.permute(0, 2, 3, 1)
    ld = ops.round_up(cout, 8) if cout >= 8 else cout
    out = torch.zeros(n, h, w, ld, device=DEV)
    ops.conv2d(a, pw, out=out, rowadd=temb, rowadd_rows=h * w)
    assert rel(out[..., :cout].permute(0, 3, 1, 2), ref) < 2e-5


def test_conv2d_1x1_residual_out2():
    torch.manual_seed(3)
    n, h, w, cin, cout = 2, 64, 4, 256, 128
    x = r16(torch.randn(n, cin, h, w, device=DEV))
    wt = r16(torch.randn(cout, cin, 1, 1, device=DEV) / math.sqrt(cin))
    b = torch.randn(cout, device=DEV)
    res = torch.randn(n, h, w, cout, device=DEV)
    ref = (F.conv2d(x, wt, b).permute(0, 2, 3, 1) + res) * 0.5
    pw = ops.pack_conv2d(wt, b)
    a = x.permute(0, 2, 3, 1).contiguous().to(DT)
    out = torch.empty(n, h, w, cout, device=DEV)
    out2 = torch.empty(n, h, w, cout, device=DEV, dtype=DT)
    ops.conv2d(a, pw, out=out, residual=res, out_scale=0.5, out2=out2, act2=ops.ACT_SILU)
    assert rel(out, ref) < 2e-5
    assert rel(out2, F.silu(ref)) < 1e-3


@pytest.mark.parametrize("k,dil,c,t,bsz", [(3, 1, 64, 300, 2), (7, 3, 128, 5121, 2), (11, 5, 32, 1000, 2), (7, 1, 64, 129, 2),
                                           # stream mode with M super-tiles (weights streamed, halo'd activation boxes)
                                           (11, 5, 128, 40968, 3), (7, 3, 256, 20484, 4), (3, 1, 512, 5121, 16),
                                           (11, 1, 256, 700, 2)])
def test_conv1d(k, dil, c, t, bsz):
    torch.manual_seed(4)
    x = r16(torch.randn(bsz, c, t, device=DEV))
    wt = r16(torch.randn(c, c, k, device=DEV) / math.sqrt(k * c))
    b = torch.randn(c, device=DEV)
    ref = F.conv1d(F.leaky_relu(x, 0.1), wt, b, dilation=dil, padding=(k * dil - dil) // 2)
    pw = ops.pack_conv1d(wt, b, dilation=dil)
    a = F.leaky_relu(x, 0.1).permute(0, 2, 1).contiguous().to(DT)
    xs = torch.randn(bsz, t, c, device=DEV)
    out = xs.clone()
    out2 = torch.empty(bsz, t, c, device=DEV, dtype=DT)
    res = torch.randn(bsz, t, c, device=DEV)
    # a is rounded to fp16 after lrelu; recompute the reference from it
    ref = F.conv1d(a.float().permute(0, 2, 1), wt, b, dilation=dil, padding=(k * dil - dil) // 2).permute(0, 2, 1)
    # residual + second (activated 16-bit) output
    o1 = torch.empty(bsz, t, c, device=DEV)
    ops.conv1d(a, pw, out=o1, residual=res, out2=out2, act2=ops.ACT_LRELU, act2_slope=0.1)
    assert rel(o1, ref + res) < 2e-5
    assert rel(out2, F.leaky_relu(ref + res, 0.1)) < 1e-3
    # accumulate: out += (conv + residual) * scale
    ops.conv1d(a, pw, out=out, residual=res, accumulate=True, out_scale=1 / 3)
    assert rel(out, xs + (ref + res) / 3) < 2e-5
    # operand-only output
    ops.conv1d(a, pw, out2=out2, act2=ops.ACT_LRELU, act2_slope=0.1)
    assert rel(out2, F.leaky_relu(ref, 0.1)) < 1e-3
    y16 = ops.lrelu_cast(out, 0.01)
    assert rel(y16, F.leaky_relu(out, 0.01)) < 1e-3


@pytest.mark.parametrize("k,s,cin,cout,t,bsz", [(16, 5, 128, 64, 50, 2), (16, 4, 64, 32, 131, 2), (8, 2, 64, 64, 200, 2),
                                                (4, 2, 64, 32, 257, 2), (16, 5, 1024, 512, 1024, 8),
                                                (16, 4, 512, 256, 5121, 8), (8, 2, 256, 128, 20484, 6)])
def test_conv_transpose1d(k, s, cin, cout, t, bsz):
    torch.manual_seed(5)
    p = (k - s) // 2
    x = r16(torch.randn(bsz, cin, t, device=DEV))
    wt = r16(torch.randn(cin, cout, k, device=DEV) / math.sqrt(k * cin))
    b = torch.randn(cout, device=DEV)
    ref = F.conv_transpose1d(x, wt, b, stride=s, padding=p)
    t_out = ref.shape[-1]
    assert t_out == (t - 1) * s - 2 * p + k
    phases = ops.pack_conv_transpose1d(wt, b, s, p)
    a = x.permute(0, 2, 1).contiguous().to(DT)
    out = torch.full((bsz, t_out, cout), float("nan"), device=DEV)
    ops.conv_transpose1d(a, phases, t_out, out=out)
    assert not torch.isnan(out).any()
    assert rel(out.permute(0, 2, 1), ref) < 2e-5


@pytest.mark.parametrize("k,dil,c,t,bsz", [(3, 1, 32, 1000, 2), (7, 3, 64, 2500, 3), (11, 5, 128, 40968, 3),
                                           (11, 1, 256, 20484, 4), (7, 1, 512, 5121, 8), (3, 5, 32, 163872, 4),
                                           # odd row count: N = 32 cannot be viewed as [rows / 2, 64] (64-byte-row fallback)
                                           (3, 1, 32, 1001, 2), (7, 1, 96, 777, 2)])
def test_conv1d_lrelu_residual_stream(k, dil, c, t, bsz):
    """HiFi-GAN ResBlock pair with the activation held only as lx = lrelu(x): c2's epilogue adds the residual recovered
    from lx (negative values * 1/slope) and emits lrelu(x') again (16-bit residual ring + inverse LeakyReLU)."""
    torch.manual_seed(31)
    slope = 0.1
    x = torch.randn(bsz, t, c, device=DEV)
    lx = F.leaky_relu(x, slope).to(DT)                       # what lives in HBM
    a = r16(torch.randn(bsz, t, c, device=DEV))              # lrelu(c1(...)) operand
    wt = r16(torch.randn(c, c, k, device=DEV) / math.sqrt(k * c))
    b = torch.randn(c, device=DEV)
    pw = ops.pack_conv1d(wt, b, dilation=dil)
    out2 = torch.empty(bsz, t, c, device=DEV, dtype=DT)
    ops.conv1d(a.to(DT), pw, residual=lx, res_neg_scale=1.0 / slope, out2=out2, act2=ops.ACT_LRELU, act2_slope=slope)
    lxf = lx.float()
    x_rec = torch.where(lxf < 0, lxf / slope, lxf)
    ref = F.conv1d(a.permute(0, 2, 1), wt, b, dilation=dil, padding=(k * dil - dil) // 2).permute(0, 2, 1) + x_rec
    assert rel(out2, F.leaky_relu(ref, slope)) < 1e-3


@pytest.mark.parametrize("c,k,dil,t,bsz", [(32, 3, 1, 1000, 2), (32, 11, 5, 5000, 2), (64, 7, 3, 2500, 3), (64, 3, 5, 300, 2),
                                           (32, 7, 1, 40968, 2), (32, 11, 1, 246, 1), (64, 7, 5, 251, 2), (32, 3, 3, 18, 3),
                                           (64, 3, 1, 81936, 3), (32, 11, 3, 163872, 2)])
def test_resblock_pair_fused(c, k, dil, t, bsz):
    """One (c1, c2) pair of a HiFi-GAN ResBlock in ONE kernel (hidden tile in shared memory) against F.conv1d in fp32
    (hifigan/models.py:56-63), on the LeakyReLU'ed 16-bit streams; also against the two-launch ctta_gemm path it replaces.
    Sequence lengths around the 256 - (k - 1) row tile exercise the zero padding of BOTH convs at the sequence ends."""
    torch.manual_seed(41)
    slope = 0.1
    assert ops.resblock_pair_supported(c, k, dil, t)
    x = torch.randn(bsz, t, c, device=DEV)
    lx = F.leaky_relu(x, slope).to(DT)
    w1 = r16(torch.randn(c, c, k, device=DEV) / math.sqrt(k * c))
    w2 = r16(torch.randn(c, c, k, device=DEV) / math.sqrt(k * c))
    b1, b2 = torch.randn(c, device=DEV) * 0.1, torch.randn(c, device=DEV) * 0.1
    pw1, pw2 = ops.pack_conv1d(w1, b1, dilation=dil), ops.pack_conv1d(w2, b2, dilation=1)
    out = ops.resblock_pair(lx, pw1, pw2, slope)
    assert out.shape == lx.shape and out.dtype == DT and torch.isfinite(out.float()).all()
    lxf = lx.float()
    x_rec = torch.where(lxf < 0, lxf / slope, lxf)
    h = F.conv1d(lxf.permute(0, 2, 1), w1, b1, dilation=dil, padding=(k * dil - dil) // 2)
    h16 = r16(F.leaky_relu(h, slope))                       # the hidden tensor is held in 16 bits (as in the unfused path)
    y = x_rec.permute(0, 2, 1) + F.conv1d(h16, w2, b2, padding=(k - 1) // 2)
    ref = F.leaky_relu(y, slope).permute(0, 2, 1)
    e = rel(out, ref)
    assert e < 1e-3, e
    # the unfused path: two implicit-GEMM launches with the hidden tensor in HBM
    tmp = torch.empty_like(lx)
    two = torch.empty_like(lx)
    ops.conv1d(lx, pw1, out2=tmp, act2=ops.ACT_LRELU, act2_slope=slope)
    ops.conv1d(tmp, pw2, residual=lx, res_neg_scale=1.0 / slope, out2=two, act2=ops.ACT_LRELU, act2_slope=slope)
    assert rel(out, two) < 1e-3


@pytest.mark.parametrize("c,k,fold,t,bsz", [(32, 3, 4, 1000, 2), (32, 7, 4, 4096, 3), (32, 11, 4, 163872, 2), (64, 3, 2, 81936, 2),
                                            (64, 7, 2, 2502, 3), (32, 11, 2, 64, 1), (64, 11, 2, 300, 2)])
def test_conv1d_time_folded(c, k, fold, t, bsz):
    """Dilation-1 HiFi-GAN convs in the time-folded form (ops.pack_conv1d_folded): `fold` time steps per GEMM row and a
    block-Toeplitz weight — same result as F.conv1d with 'same' zero padding (hifigan/models.py:16-17,59,61), on the
    LeakyReLU'ed 16-bit stream with the residual epilogue of the ResBlock."""
    torch.manual_seed(17)
    slope = 0.1
    x = torch.randn(bsz, t, c, device=DEV)
    lx = F.leaky_relu(x, slope).to(DT)
    a = r16(torch.randn(bsz, t, c, device=DEV))
    wt = r16(torch.randn(c, c, k, device=DEV) / math.sqrt(k * c))
    b = torch.randn(c, device=DEV)
    pw = ops.pack_conv1d_folded(wt, b, fold)
    assert pw.n == fold * c and pw.c == fold * c and pw.ntaps <= (k + 2 * fold - 2) // fold + 1
    out2 = torch.empty(bsz, t, c, device=DEV, dtype=DT)

    def fv(z):
        return z.view(bsz, t // fold, fold * c)
    ops.conv1d(fv(a.to(DT)), pw, residual=fv(lx), res_neg_scale=1.0 / slope, out2=fv(out2), act2=ops.ACT_LRELU,
               act2_slope=slope)
    lxf = lx.float()
    x_rec = torch.where(lxf < 0, lxf / slope, lxf)
    ref = F.conv1d(a.permute(0, 2, 1), wt, b, padding=(k - 1) // 2).permute(0, 2, 1) + x_rec
    assert rel(out2, F.leaky_relu(ref, slope)) < 1e-3
    # and against the unfolded kernel path
    direct = torch.empty_like(out2)
    ops.conv1d(a.to(DT), ops.pack_conv1d(wt, b), residual=lx, res_neg_scale=1.0 / slope, out2=direct, act2=ops.ACT_LRELU,
               act2_slope=slope)
    assert rel(out2, direct) < 1e-3


@pytest.mark.parametrize("c,k,dil,t,bsz", [(64, 3, 1, 40000, 3), (64, 7, 3, 20000, 3), (32, 3, 5, 60000, 3), (32, 11, 5, 50000, 2)])
def test_resblock_pair_bit_identical_to_two_launches_and_reproducible(c, k, dil, t, bsz):
    """Many tiles per CTA (the two-slot pipeline in steady state): every launch gives the same bits, and those bits are the
    two-launch ctta_gemm path's (same products, same fp32 operation order in the epilogue)."""
    torch.manual_seed(c + k + dil)
    lx = F.leaky_relu(torch.randn(bsz, t, c, device=DEV), 0.1).to(DT)
    w1 = torch.randn(c, c, k, device=DEV) / math.sqrt(k * c)
    w2 = torch.randn(c, c, k, device=DEV) / math.sqrt(k * c)
    b1, b2 = torch.randn(c, device=DEV) * 0.1, torch.randn(c, device=DEV) * 0.1
    pw1, pw2 = ops.pack_conv1d(w1, b1, dilation=dil), ops.pack_conv1d(w2, b2, dilation=1)
    tmp, two = torch.empty_like(lx), torch.empty_like(lx)
    ops.conv1d(lx, pw1, out2=tmp, act2=ops.ACT_LRELU, act2_slope=0.1)
    ops.conv1d(tmp, pw2, residual=lx, res_neg_scale=10.0, out2=two, act2=ops.ACT_LRELU, act2_slope=0.1)
    first = ops.resblock_pair(lx, pw1, pw2, 0.1).clone()
    assert torch.equal(first, two)
    for _ in range(25):
        assert torch.equal(ops.resblock_pair(lx, pw1, pw2, 0.1), first)


def test_resblock_pair_input_aligned_to_16_bytes_only():
    """The residual rows are read with 256-bit loads when the input is 32-byte aligned, else with 128-bit loads."""
    torch.manual_seed(3)
    c, k, t, bsz = 32, 7, 3000, 2
    buf = torch.empty(bsz * t * c + 8, device=DEV, dtype=DT)
    lx = buf[8:].view(bsz, t, c)
    assert lx.data_ptr() % 32 == 16
    lx.copy_(F.leaky_relu(torch.randn(bsz, t, c, device=DEV), 0.1))
    w1 = torch.randn(c, c, k, device=DEV) / math.sqrt(k * c)
    w2 = torch.randn(c, c, k, device=DEV) / math.sqrt(k * c)
    b = torch.randn(c, device=DEV) * 0.1
    pw1, pw2 = ops.pack_conv1d(w1, b, dilation=3), ops.pack_conv1d(w2, b, dilation=1)
    aligned = lx.clone()
    assert aligned.data_ptr() % 32 == 0
    assert torch.equal(ops.resblock_pair(lx, pw1, pw2, 0.1), ops.resblock_pair(aligned, pw1, pw2, 0.1))


def test_resblock_pair_unsupported_shapes_are_refused():
    assert not ops.resblock_pair_supported(64, 11, 1)        # resident weights of both convs exceed shared memory
    assert not ops.resblock_pair_supported(128, 3, 1) and not ops.resblock_pair_supported(32, 4, 1) and not ops.resblock_pair_supported(32, 3, 1, 1001)
    lx = torch.zeros(1, 64, 128, device=DEV, dtype=DT)
    pw = ops.pack_conv1d(torch.zeros(128, 128, 3, device=DEV), torch.zeros(128, device=DEV))
    with pytest.raises(_lib.CttaError):
        ops.resblock_pair(lx, pw, pw, 0.1)


@pytest.mark.parametrize("kind", ["conv1d_stream", "linear", "conv2d_stream"])
def test_multicast_cta_pairs_bit_identical(kind, monkeypatch):
    """The CTA-pair (cluster of 2, TMA-multicast weight chunks) launch computes exactly what independent CTAs compute:
    same K order, same epilogue.  Shapes with an ODD number of M tiles exercise the padding tile of the last pair
    (loads zero-filled, stores clipped), with fused GroupNorm moments and a residual in flight."""
    torch.manual_seed(41)
    if kind == "conv1d_stream":
        bsz, t, c = 3, 20484, 256                                  # 161 M tiles per sample x 3 = 483 (odd)
        x = torch.randn(bsz, t, c, device=DEV).to(DT)
        res = torch.randn(bsz, t, c, device=DEV).to(DT)
        pw = ops.pack_conv1d(r16(torch.randn(c, c, 7, device=DEV) / math.sqrt(7 * c)), torch.randn(c, device=DEV), dilation=3)

        def run():
            out = torch.empty(bsz, t, c, device=DEV, dtype=DT)
            ops.conv1d(x, pw, residual=res, res_neg_scale=10.0, out2=out, act2=ops.ACT_LRELU, act2_slope=0.1)
            return out, None
    elif kind == "linear":
        m, k, n = 301 * 128 - 40, 320, 512                          # 301 M tiles (odd), ragged last tile, 2 N tiles
        a = torch.randn(m, k, device=DEV).to(DT)
        res = torch.randn(m, n, device=DEV)
        pw = ops.pack_linear(r16(torch.randn(n, k, device=DEV) / math.sqrt(k)), torch.randn(n, device=DEV))

        def run():
            out = torch.empty(m, n, device=DEV)
            ops.linear(a, pw, out=out, residual=res)
            return out, None
    else:
        n, h, w, c = 37, 264, 16, 128                              # 33 tiles of 128 pixels per image x 37 = 1221 (odd)
        x = torch.randn(n, h, w, c, device=DEV).to(DT)
        pw = ops.pack_conv2d(r16(torch.randn(256, c, 3, 3, device=DEV) / math.sqrt(9 * c)), torch.randn(256, device=DEV))

        def run():
            out = torch.empty(n, h, w, 256, device=DEV)
            st = torch.empty(n, 32, 2, device=DEV)
            ops.conv2d(x, pw, out=out, stats=st, stats_groups=32)
            return out, st
    monkeypatch.setenv("CTTA_NO_MCAST", "1")
    ref, ref_st = run()
    monkeypatch.delenv("CTTA_NO_MCAST")
    got, got_st = run()
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    if ref_st is not None:      # moments are accumulated with fp32 atomics: equal up to summation order
        assert rel(got_st, ref_st) < 1e-5


def test_bf16_operands_generic_epilogue():
    """The C-ABI also takes bf16 operands (generic epilogue instance): fp32-output linear, the HiFi-GAN residual stream
    with 128-byte epilogue rows (C = 64), the paired-row N = 32 layout, and a GEGLU linear with bf16 output."""
    torch.manual_seed(43)
    bf = torch.bfloat16
    rb = lambda t: t.to(bf).float()
    # linear, fp32 out + residual
    m, k, n = 1000, 256, 320
    a, w, b = rb(torch.randn(m, k, device=DEV)), rb(torch.randn(n, k, device=DEV) / math.sqrt(k)), torch.randn(n, device=DEV)
    res = torch.randn(m, n, device=DEV)
    out = torch.empty(m, n, device=DEV)
    ops.linear(a.to(bf), ops.pack_linear(w, b, dtype=bf), out=out, residual=res)
    assert rel(out, a @ w.t() + b + res) < 2e-5
    # GEGLU, bf16 out
    d = 256
    w2, b2 = rb(torch.randn(2 * d, k, device=DEV) / math.sqrt(k)), torch.randn(2 * d, device=DEV)
    wi = torch.stack([w2[:d], w2[d:]], dim=1).reshape(2 * d, k)
    bi = torch.stack([b2[:d], b2[d:]], dim=1).reshape(2 * d)
    o16 = torch.empty(m, d, device=DEV, dtype=bf)
    ops.linear(a.to(bf), ops.pack_linear(wi, bi, dtype=bf), out=o16, act=ops.ACT_GEGLU)
    h = a @ w2.t() + b2
    assert rel(o16, h[:, :d] * F.gelu(h[:, d:])) < 6e-3
    # conv1d residual stream (lrelu'ed 16-bit residual -> lrelu'ed 16-bit operand): C = 64 (wide rows), C = 32 (paired rows)
    for c, t in ((64, 3000), (32, 3000)):
        slope = 0.1
        lx = F.leaky_relu(torch.randn(2, t, c, device=DEV), slope).to(bf)
        x = rb(torch.randn(2, t, c, device=DEV))
        wt, bc = rb(torch.randn(c, c, 7, device=DEV) / math.sqrt(7 * c)), torch.randn(c, device=DEV)
        out2 = torch.empty(2, t, c, device=DEV, dtype=bf)
        ops.conv1d(x.to(bf), ops.pack_conv1d(wt, bc, dilation=1, dtype=bf), residual=lx, res_neg_scale=1.0 / slope, out2=out2,
                   act2=ops.ACT_LRELU, act2_slope=slope)
        lxf = lx.float()
        ref = F.conv1d(x.permute(0, 2, 1), wt, bc, padding=3).permute(0, 2, 1) + torch.where(lxf < 0, lxf / slope, lxf)
        assert rel(out2, F.leaky_relu(ref, slope)) < 6e-3


def test_bmm_nt_batched_weights():
    """Per-image weight matrices (VAE AttnBlock: q.k^T, P.V, W_v.a^T for every sample in one launch)."""
    torch.manual_seed(33)
    bsz, m, n, k = 5, 300, 640, 192
    a = r16(torch.randn(bsz, m, k, device=DEV))
    b = r16(torch.randn(bsz, n, k, device=DEV) / math.sqrt(k))
    bias = torch.randn(n, device=DEV)
    out = torch.empty(bsz, m, n, device=DEV)
    ops.bmm_nt(a.to(DT), b.to(DT), bias=bias, out=out)
    assert rel(out, torch.einsum("bmk,bnk->bmn", a, b) + bias) < 2e-5
    out16 = torch.empty(bsz, m, n, device=DEV, dtype=DT)
    ops.bmm_nt(a.to(DT), b.to(DT), out=out16)
    assert rel(out16, torch.einsum("bmk,bnk->bmn", a, b)) < 1e-3


@pytest.mark.parametrize("mode,c,k", [("conv2d", 128, 3), ("conv1d", 32, 7), ("conv1d", 64, 3)])
def test_single_output_channel_conv(mode, c, k):
    """n == 1 multi-tap convolutions (VAE conv_out, HiFi-GAN conv_post + tanh) through the generic implicit-GEMM path (direct-epilogue tcgen05; the models use the pointwise-GEMM + tap-sum
    restatement tested below)."""
    torch.manual_seed(34)
    if mode == "conv2d":
        n, h, w = 3, 96, 64
        x = r16(torch.randn(n, c, h, w, device=DEV))
        wt = r16(torch.randn(1, c, k, k, device=DEV) / math.sqrt(k * k * c))
        b = torch.randn(1, device=DEV)
        ref = F.conv2d(x, wt, b, padding=k // 2).permute(0, 2, 3, 1)
        out = torch.full((n, h, w, 1), float("nan"), device=DEV)
        out16 = torch.empty(n, h, w, 1, device=DEV, dtype=DT)
        ops.conv2d(x.permute(0, 2, 3, 1).contiguous().to(DT), ops.pack_conv2d(wt, b), out=out, out2=out16)
        assert rel(out, ref) < 2e-5 and rel(out16, ref) < 1e-3
    else:
        bsz, t = 3, 5000
        x = r16(torch.randn(bsz, c, t, device=DEV))
        wt = r16(torch.randn(1, c, k, device=DEV) / math.sqrt(k * c))
        b = torch.randn(1, device=DEV)
        ref = torch.tanh(F.conv1d(x, wt, b, padding=k // 2)).permute(0, 2, 1)
        out = torch.full((bsz, t, 1), float("nan"), device=DEV)
        ops.conv1d(x.permute(0, 2, 1).contiguous().to(DT), ops.pack_conv1d(wt, b), out=out, act=ops.ACT_TANH)
        assert rel(out, ref) < 2e-5


@pytest.mark.parametrize("mode,c,k,dil", [("conv2d", 128, 3, 1), ("conv1d", 32, 7, 1), ("conv1d", 64, 3, 1),
                                          ("conv1d", 32, 11, 3), ("conv2d", 48, 3, 1)])
def test_single_channel_conv_as_pointwise_gemm_plus_tap_sum(mode, c, k, dil):
    """The product path for n == 1 convolutions: z = x . w_taps (pointwise tcgen05 GEMM, N = taps) then
    y[p] = act(bias + sum_j z[p + shift_j, j]) (ctta_tap_sum), against F.conv1d / F.conv2d with zero padding."""
    torch.manual_seed(35)
    if mode == "conv2d":
        n, h, w = 3, 96, 64
        x = r16(torch.randn(n, c, h, w, device=DEV))
        wt = r16(torch.randn(1, c, k, k, device=DEV) / math.sqrt(k * k * c))
        b = torch.randn(1, device=DEV)
        ref = F.conv2d(x, wt, b, padding=k // 2).permute(0, 2, 3, 1)
        out = torch.full((n, h, w, 1), float("nan"), device=DEV)
        out16 = torch.empty(n, h, w, 1, device=DEV, dtype=DT)
        sc = ops.pack_single_channel_conv(wt, b)
        ops.single_channel_conv(x.permute(0, 2, 3, 1).contiguous().to(DT), sc, n, h, w, out=out, out16=out16)
        assert rel(out, ref) < 2e-5 and rel(out16, ref) < 1e-3
    else:
        bsz, t = 3, 5004
        x = r16(torch.randn(bsz, c, t, device=DEV))
        wt = r16(torch.randn(1, c, k, device=DEV) / math.sqrt(k * c))
        b = torch.randn(1, device=DEV)
        ref = torch.tanh(F.conv1d(x, wt, b, padding=dil * (k // 2), dilation=dil)).permute(0, 2, 1)
        out = torch.full((bsz, t), float("nan"), device=DEV)
        sc = ops.pack_single_channel_conv(wt, b, dilation=dil)
        ops.single_channel_conv(x.permute(0, 2, 1).contiguous().to(DT), sc, bsz, 1, t, act=ops.ACT_TANH, out=out)
        assert rel(out.view(bsz, t, 1), ref) < 2e-5


@pytest.mark.parametrize("n,h,w,c,cout", [(2, 256, 16, 64, 128), (2, 64, 4, 96, 256), (3, 128, 8, 64, 128),
                                          (2, 512, 32, 64, 256), (1, 128, 64, 32, 128)])
def test_upsample2x_conv_phases(n, h, w, c, cout):
    """nearest-2x upsample + conv3x3 as four 2x2 phase convs on the low-resolution input, with fused GroupNorm moments."""
    torch.manual_seed(35)
    x = r16(torch.randn(n, c, h, w, device=DEV))
    wt = r16(torch.randn(cout, c, 3, 3, device=DEV) / math.sqrt(9 * c))
    b = torch.randn(cout, device=DEV)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), wt, b, padding=1).permute(0, 2, 3, 1)
    phases = ops.pack_upsample2x_conv2d(wt, b)
    out = torch.full((n, 2 * h, 2 * w, cout), float("nan"), device=DEV)
    stats = torch.full((n, 32, 2), float("nan"), device=DEV)
    assert ops.upsample2x_conv_supported(h, w)
    ops.conv2d_upsample2x(x.permute(0, 2, 3, 1).contiguous().to(DT), phases, out, stats=stats, stats_groups=32)
    assert not torch.isnan(out).any()
    assert rel(out, ref) < 1e-3      # summed taps are rounded to fp16 once (reference sums fp32 products)
    assert rel(stats, _moments(ref, 32).float()) < 2e-3


def test_mrf_combine():
    torch.manual_seed(32)
    slope = 0.1
    xs = [torch.randn(3, 1000, 64, device=DEV) for _ in range(3)]
    lxs = [F.leaky_relu(x, slope).to(DT) for x in xs]
    rec = [torch.where(l.float() < 0, l.float() / slope, l.float()) for l in lxs]
    for out_slope in (0.1, 0.01):
        y = ops.mrf_combine(lxs, slope, 1.0 / 3, out_slope)
        assert rel(y, F.leaky_relu(sum(rec) / 3, out_slope)) < 1e-3


def _moments(y_nhwc, groups):
    """y [N, ..., C] fp32 -> [N, groups, 2] (sum, sum of squares) in float64."""
    n, c = y_nhwc.shape[0], y_nhwc.shape[-1]
    yg = y_nhwc.double().reshape(n, -1, groups, c // groups)
    return torch.stack([yg.sum(dim=(1, 3)), (yg * yg).sum(dim=(1, 3))], dim=-1)


@pytest.mark.parametrize("n,h,w,cin,cout,out16", [(3, 64, 16, 64, 128, False), (2, 128, 8, 32, 256, True),
                                                  (5, 32, 2, 128, 512, False), (2, 256, 16, 64, 1024, True),
                                                  (24, 256, 16, 64, 128, True)])
def test_conv2d_fused_groupnorm_moments(n, h, w, cin, cout, out16):
    """The conv epilogue accumulates the next GroupNorm's (sum, sum^2) per (image, group): channels per group 4..32,
    fp32 and 16-bit outputs, with the time-embedding row add and a residual in the value being measured."""
    torch.manual_seed(21)
    g = 32
    x = r16(torch.randn(n, h, w, cin, device=DEV))
    wt = r16(torch.randn(cout, cin, 3, 3, device=DEV) / math.sqrt(9 * cin))
    b = torch.randn(cout, device=DEV)
    temb = torch.randn(n, cout, device=DEV)
    ref = F.conv2d(x.permute(0, 3, 1, 2), wt, b, padding=1).permute(0, 2, 3, 1) + temb[:, None, None, :]
    pw = ops.pack_conv2d(wt, b)
    stats = torch.full((n, g, 2), float("nan"), device=DEV)
    if out16:
        out = torch.empty(n, h, w, cout, device=DEV, dtype=DT)
        ops.conv2d(x.to(DT), pw, out=out, rowadd=temb, rowadd_rows=h * w, stats=stats, stats_groups=g)
    else:
        res = torch.randn(n, h, w, cout, device=DEV)
        ref = ref + res
        out = torch.empty(n, h, w, cout, device=DEV)
        ops.conv2d(x.to(DT), pw, out=out, rowadd=temb, rowadd_rows=h * w, residual=res, stats=stats, stats_groups=g)
    assert rel(out, ref) < (1e-3 if out16 else 2e-5)
    want = _moments(ref, g)
    assert rel(stats, want.float()) < 1e-4
    # and they normalise like F.group_norm
    gam, bet = torch.randn(cout, device=DEV), torch.randn(cout, device=DEV)
    y = ops.groupnorm_apply(out, g, stats, gam, bet, eps=1e-5, act=ops.ACT_SILU)
    yref = F.silu(F.group_norm(ref.permute(0, 3, 1, 2), g, gam, bet, 1e-5)).permute(0, 2, 3, 1)
    assert rel(y, yref) < (3e-3 if out16 else 1.5e-3)


def test_linear_fused_groupnorm_moments():
    """ROWS mode: image = row / rows_per_image (transformer proj_out, stride-2 downsampler GEMM); ragged last tile."""
    torch.manual_seed(22)
    n_img, hw, k, c, g = 5, 96, 320, 256, 32
    a = r16(torch.randn(n_img * hw, k, device=DEV))
    wt = r16(torch.randn(c, k, device=DEV) / math.sqrt(k))
    b = torch.randn(c, device=DEV)
    res = torch.randn(n_img * hw, c, device=DEV)
    ref = a @ wt.t() + b + res
    stats = torch.empty(n_img, g, 2, device=DEV)
    out = torch.empty(n_img * hw, c, device=DEV)
    ops.linear(a.to(DT), ops.pack_linear(wt, b), out=out, residual=res, stats=stats, stats_groups=g, stats_rows_per_img=hw)
    assert rel(out, ref) < 2e-5
    assert rel(stats, _moments(ref.view(n_img, hw, c), g).float()) < 1e-4


def test_im2col_s2_conv():
    torch.manual_seed(6)
    n, h, w, c, cout = 2, 32, 8, 64, 96
    x = r16(torch.randn(n, c, h, w, device=DEV))
    wt = r16(torch.randn(cout, c, 3, 3, device=DEV) / math.sqrt(9 * c))
    b = torch.randn(cout, device=DEV)
    ref = F.conv2d(x, wt, b, stride=2, padding=1)
    a = ops.im2col_s2(x.permute(0, 2, 3, 1).contiguous().to(DT))
    pw = ops.pack_conv2d_im2col(wt, b)
    out = torch.empty(a.shape[0], cout, device=DEV)
    ops.linear(a, pw, out=out)
    assert rel(out.reshape(n, h // 2, w // 2, cout).permute(0, 3, 1, 2), ref) < 2e-5


@pytest.mark.parametrize("n,h,w,c,c2,g", [(2, 64, 4, 256, 0, 32), (3, 32, 16, 512, 256, 32), (1, 128, 64, 128, 0, 32)])
def test_groupnorm(n, h, w, c, c2, g):
    torch.manual_seed(7)
    x = torch.randn(n, h, w, c, device=DEV) * 2 + 0.5
    x2 = torch.randn(n, h, w, c2, device=DEV) if c2 else None
    gamma = torch.randn(c + c2, device=DEV)
    beta = torch.randn(c + c2, device=DEV)
    xc = torch.cat([x, x2], -1) if c2 else x
    ref = F.silu(F.group_norm(xc.permute(0, 3, 1, 2), g, gamma, beta, eps=1e-6)).permute(0, 2, 3, 1)
    st = ops.groupnorm_stats(x, g, x2=x2)
    raw = torch.empty(n, h, w, c + c2, device=DEV, dtype=DT)
    y = ops.groupnorm_apply(x, g, st, gamma, beta, eps=1e-6, x2=x2, raw_out=raw)
    assert rel(y, ref) < 1e-3
    assert rel(raw, xc) < 1e-3
    yu = ops.groupnorm_apply(x, g, st, gamma, beta, eps=1e-6, x2=x2, upsample=True)
    refu = ref.repeat_interleave(2, 1).repeat_interleave(2, 2)
    assert rel(yu, refu) < 1e-3
    # plain cast path
    yc = ops.groupnorm_apply(x, g, None, None, None, act=ops.ACT_NONE, x2=x2)
    assert rel(yc, xc) < 1e-3


@pytest.mark.parametrize("m,d", [(1000, 255), (77, 510), (300, 1020)])
def test_layernorm(m, d):
    torch.manual_seed(8)
    ld = ops.round_up(d, 64)
    x = torch.zeros(m, ld, device=DEV)
    x[:, :d] = torch.randn(m, d, device=DEV) * 3 + 1
    gamma = torch.randn(d, device=DEV)
    beta = torch.randn(d, device=DEV)
    gp = torch.zeros(ld, device=DEV)
    bp = torch.zeros(ld, device=DEV)
    gp[:d] = gamma
    bp[:d] = beta
    y = ops.layernorm(x, d, gp, bp, 1e-5)
    ref = F.layer_norm(x[:, :d], (d,), gamma, beta, 1e-5)
    assert rel(y[:, :d], ref) < 1e-3
    assert (y[:, d:] == 0).all()


@pytest.mark.parametrize("b,h,lq,lk,masked", [(2, 5, 256, 256, False), (1, 10, 1024, 1024, False),
                                              (3, 20, 64, 32, True), (2, 5, 4096, 77, True), (2, 4, 100, 200, False)])
def test_attention(b, h, lq, lk, masked):
    torch.manual_seed(9)
    d = 64
    q = r16(torch.randn(b, lq, h, d, device=DEV))
    k = r16(torch.randn(b, lk, h, d, device=DEV))
    v = r16(torch.randn(b, lk, h, d, device=DEV))
    q[..., 51:] = 0
    k[..., 51:] = 0
    v[..., 51:] = 0
    scale = 51 ** -0.5
    kv_len = None
    bias = None
    if masked:
        kv_len = torch.randint(1, lk + 1, (b,), device=DEV, dtype=torch.int32)
        mask = torch.arange(lk, device=DEV)[None, :] < kv_len[:, None]
        bias = ((1 - mask.float()) * -10000.0)[:, None, None, :]
    ref = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3),
                                         attn_mask=bias, scale=scale).permute(0, 2, 1, 3)
    out = ops.attention(q.to(DT), k.to(DT), v.to(DT), scale, kv_len=kv_len)
    assert rel(out, ref) < 2e-3


@pytest.mark.parametrize("b,h,lq,lk,masked,amp", [(2, 5, 4096, 4096, False, 1.0), (2, 3, 300, 500, True, 1.0),
                                                  (1, 2, 128, 128, False, 1.0), (2, 2, 640, 1152, True, 8.0),
                                                  (1, 1, 256, 2048, False, 20.0), (1, 2, 256, 1024, True, 80.0)])
def test_attention_tcgen05_strided(b, h, lq, lk, masked, amp):
    """Self-attention sized problems run on the tcgen05 kernel: q/k/v are strided views of one fused buffer (the
    UNet's qkv GEMM output); `amp` scales the scores so that later key tiles beat the running max: by more than 2^15
    (amp 8 / 20: the optimistic tile evaluation fails its row-sum check and is redone against the true max) and by
    more than 2^128 (amp 80: the optimistic exponentials overflow to inf); ragged lq / lk exercise TMA zero fill and the
    key-length mask."""
    torch.manual_seed(12)
    d = 64
    lm = max(lq, lk)
    buf = torch.randn(b, lm, 3, h, d, device=DEV)
    buf[:, :, 0] *= amp
    # make the max grow along the key axis so the rescale path is taken repeatedly
    buf[:, :, 1] *= torch.linspace(0.2, 1.0, lm, device=DEV)[None, :, None, None]
    buf[..., 51:] = 0
    buf16 = buf.to(DT)
    q, k, v = buf16[:, :lq, 0], buf16[:, :lk, 1], buf16[:, :lk, 2]
    scale = 51 ** -0.5
    kv_len = None
    bias = None
    if masked:
        kv_len = torch.randint(1, lk + 1, (b,), device=DEV, dtype=torch.int32)
        kv_len[0] = lk
        mask = torch.arange(lk, device=DEV)[None, :] < kv_len[:, None]
        bias = torch.zeros(b, 1, 1, lk, device=DEV).masked_fill(~mask[:, None, None, :], float("-inf"))
    ref = F.scaled_dot_product_attention(q.float().permute(0, 2, 1, 3), k.float().permute(0, 2, 1, 3),
                                         v.float().permute(0, 2, 1, 3), attn_mask=bias, scale=scale).permute(0, 2, 1, 3)
    out = ops.attention(q, k, v, scale, kv_len=kv_len)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref) < 3e-3


def test_softmax_rows():
    torch.manual_seed(10)
    x = torch.randn(300, 4096, device=DEV) * 20
    y = ops.softmax_rows(x, 512 ** -0.5)
    assert rel(y, torch.softmax(x * 512 ** -0.5, -1)) < 1e-3


def test_small_ops():
    torch.manual_seed(11)
    b = 5
    x = torch.randn(b, 8, 256, 16, device=DEV)
    y = ops.nchw_to_nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(y), x)
    t = torch.full((b,), 999.0, device=DEV)
    w = torch.rand(b, device=DEV) * 5
    gw = torch.randn(512, device=DEV)
    tf, gf = ops.time_features(t, w, gw)
    f = torch.exp(-math.log(10000) * torch.arange(128, device=DEV, dtype=torch.float32) / 128)
    arg = t[:, None] * f[None]
    assert torch.allclose(tf, torch.cat([arg.cos(), arg.sin()], 1), atol=2e-4)
    garg = (w.double()[:, None] * gw.double()[None]) * 2 * math.pi
    assert torch.allclose(gf, torch.cat([garg.cos(), garg.sin()], 1).float(), atol=1e-5)
    xw = torch.randn(b, 1024, device=DEV)
    wt = torch.randn(300, 1024, device=DEV) / 32
    bs = torch.randn(300, device=DEV)
    out = ops.small_linear(xw, wt, bs, act_in=ops.ACT_SILU, act_out=ops.ACT_SILU)
    assert rel(out, F.silu(F.silu(xw) @ wt.t() + bs)) < 1e-5
    wav = torch.randn(3, 163872, device=DEV) * 0.3 + 0.05
    i16, mm = ops.wave_to_int16(wav)
    c = (wav.max() + wav.min()) / 2
    ref = ((wav - c).cpu().numpy() * 32768).astype("int16")
    assert (i16.cpu().numpy() == ref).all()
    z = torch.randn(4, 8, 256, 16, device=DEV)
    assert torch.allclose(ops.cfg_mix(z, 3.0), (1 - 3.0) * z[:2] + 3.0 * z[2:], atol=1e-6)
