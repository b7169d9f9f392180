"""Python wrappers over the C-ABI kernels of libctta.so.

torch is used for device memory and streams only; every arithmetic op on the hot path is a libctta kernel.
All activations are channels-last: images [N, H, W, C], sequences [B, T, C], token matrices [M, C].
"""
import ctypes as C
import math

import torch

from . import _lib
from ._lib import (A_CONV1D, A_CONV2D, A_ROWS, ACT_GEGLU, ACT_LRELU, ACT_NONE, ACT_SILU, ACT_TANH, BF16, F16, F32,
                   GemmDesc, check, lib)

# 16-bit operand type of the tensor-core path.  fp16 (10-bit mantissa) is what keeps the UNet / VAE within the
# 1e-2 rel-L2 tolerance against the fp32 reference (the reference's own bf16 autocast does not).
OPERAND_DTYPE = torch.float16


class operand_dtype:
    """Context manager: run a stage with another 16-bit operand type (torch.bfloat16 when fp16's 65504 range overflowed —
    the reason the reference config carries `upcast_attention`: SD-2.x activations do overflow fp16).  bf16 has fp32's
    range and 3 fewer mantissa bits; packed operands are kept per dtype (PackedModule.packed)."""

    def __init__(self, dtype):
        assert dtype in (torch.float16, torch.bfloat16)
        self.dtype = dtype

    def __enter__(self):
        global OPERAND_DTYPE
        self.prev = OPERAND_DTYPE
        OPERAND_DTYPE = self.dtype
        return self

    def __exit__(self, *exc):
        global OPERAND_DTYPE
        OPERAND_DTYPE = self.prev
        return False

_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


def _require_cuda(t):
    if not t.is_cuda:
        raise _lib.CttaError("consistencytta_b200 ops need CUDA tensors (no CPU fallback)")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the current device's stream; a tensor of another GPU would be dereferenced on the wrong one
        raise _lib.CttaError("tensor lives on %s but the current CUDA device is %d: wrap the call in "
                             "torch.cuda.device(tensor.device)" % (t.device, torch.cuda.current_device()))


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def round_up(x, m):
    return (x + m - 1) // m * m


# ------------------------------------------------------------------------------------------------ weight packing
class PackedWeight:
    """K-major 16-bit GEMM operand W[n, ntaps * c_pad] (+ fp32 bias) with the tap shifts of the implicit GEMM."""

    def __init__(self, w16, bias, ntaps, c, n, d0, d1, out_stride=1, out_off=0):
        self.w = w16
        self.bias = bias
        self.ntaps = ntaps
        self.c = c
        self.n = n
        self.d0 = list(d0)
        self.d1 = list(d1)
        self.out_stride = out_stride
        self.out_off = out_off


def _pack_taps(w_ntc, dtype):
    """w_ntc: [n, ntaps, c] fp32 -> [n, ntaps * c_pad] 16-bit, zero padded per tap to a multiple of 64."""
    n, ntaps, c = w_ntc.shape
    c_pad = round_up(c, 64)
    out = torch.zeros(n, ntaps, c_pad, dtype=dtype, device=w_ntc.device)
    out[:, :, :c] = w_ntc.to(dtype)
    return out.reshape(n, ntaps * c_pad).contiguous()


def _bias(b, n, device):
    if b is None:
        return None
    return b.detach().to(device=device, dtype=torch.float32).contiguous()


def pack_linear(weight, bias=None, dtype=None, k_pad_to=None, n_pad_to=None):
    """nn.Linear weight [n, k] (optionally zero padded to [n_pad_to, k_pad_to])."""
    dtype = dtype or OPERAND_DTYPE
    w = weight.detach().float()
    n, k = w.shape
    kp = k_pad_to or k
    np_ = n_pad_to or n
    wp = torch.zeros(np_, kp, device=w.device)
    wp[:n, :k] = w
    b = None
    if bias is not None:
        b = torch.zeros(np_, device=w.device)
        b[:n] = bias.detach().float()
    return PackedWeight(_pack_taps(wp[:, None, :], dtype), b, 1, kp, np_, [0], [0])


def pack_conv2d(weight, bias=None, dtype=None):
    """nn.Conv2d weight [cout, cin, kh, kw], stride 1, 'same' padding -> taps ordered (kh, kw)."""
    dtype = dtype or OPERAND_DTYPE
    w = weight.detach().float()
    cout, cin, kh, kw = w.shape
    w_ntc = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    d0, d1 = [], []
    for i in range(kh):
        for j in range(kw):
            d1.append(i - (kh - 1) // 2)
            d0.append(j - (kw - 1) // 2)
    return PackedWeight(_pack_taps(w_ntc, dtype), _bias(bias, cout, w.device), kh * kw, cin, cout, d0, d1)


def pack_conv2d_im2col(weight, bias=None, dtype=None):
    """3x3 conv weight for a pre-gathered A [M, 9 * cin] (stride-2 downsamplers): one tap, K = 9 * cin."""
    dtype = dtype or OPERAND_DTYPE
    w = weight.detach().float()
    cout, cin, kh, kw = w.shape
    w_k = w.permute(0, 2, 3, 1).reshape(cout, 1, kh * kw * cin)
    return PackedWeight(_pack_taps(w_k, dtype), _bias(bias, cout, w.device), 1, kh * kw * cin, cout, [0], [0])


def pack_upsample2x_conv2d(weight, bias=None, dtype=None):
    """nearest-2x upsample followed by a 3x3 'same' conv == four 2x2 convs on the low-resolution input, one per output
    phase (ph, pw): output row 2h + ph reads low-res rows h + floor((ph + kh - 1) / 2), i.e. {h-1, h} for ph = 0 (kernel
    rows {0}, {1, 2}) and {h, h+1} for ph = 1 (kernel rows {0, 1}, {2}); the same along the width.  Kernel taps that land
    on the same low-res pixel are summed.  Returns [(PackedWeight, phase_code)] with phase_code = 1 + 2 ph + pw."""
    dtype = dtype or OPERAND_DTYPE
    w = weight.detach().float()
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    rows = {0: [(-1, [0]), (0, [1, 2])], 1: [(0, [0, 1]), (1, [2])]}
    out = []
    for ph in (0, 1):
        for pw_ in (0, 1):
            taps, d0, d1 = [], [], []
            for dh, khs in rows[ph]:
                for dw, kws in rows[pw_]:
                    taps.append(w[:, :, khs][:, :, :, kws].sum(dim=(2, 3)))   # [cout, cin]
                    d1.append(dh)
                    d0.append(dw)
            w_ntc = torch.stack(taps, dim=1)                                   # [cout, 4, cin]
            out.append((PackedWeight(_pack_taps(w_ntc, dtype), _bias(bias, cout, w.device), 4, cin, cout, d0, d1),
                        1 + 2 * ph + pw_))
    return out


def pack_conv1d(weight, bias=None, dilation=1, dtype=None):
    """nn.Conv1d weight [cout, cin, k], stride 1, padding (k*d - d)/2 (hifigan/models.py:16-17)."""
    dtype = dtype or OPERAND_DTYPE
    w = weight.detach().float()
    cout, cin, k = w.shape
    w_ntc = w.permute(0, 2, 1).contiguous()
    d0 = [(j - (k - 1) // 2) * dilation for j in range(k)]
    return PackedWeight(_pack_taps(w_ntc, dtype), _bias(bias, cout, w.device), k, cin, cout, d0, [0] * k)


def pack_conv1d_folded(weight, bias, fold, dtype=None):
    """nn.Conv1d weight [cout, cin, k] (stride 1, dilation 1, 'same' padding) for the FOLDED view of the sequence:
    `fold` consecutive time steps form one GEMM row, x[B, T, cin] read as [B, T / fold, fold * cin] (the same memory) and
    y[B, T, cout] written as [B, T / fold, fold * cout].  Output sub-position g of folded row m needs input time
    fold*m + g + s (s = j - (k-1)/2) = fold*(m + r) + f, so kernel tap j = fold*r + f - g + (k-1)/2 lands in folded tap r,
    input block f, output block g of a block-Toeplitz weight [fold*cout, taps', fold*cin] (zero where j falls outside
    [0, k)).  Exact (zeros are added), and the conv's zero padding in time is the folded conv's zero padding in rows.
    Why: the narrow HiFi-GAN stages (C = 32 / 64) are bound by the tcgen05.mma issue floor (~60 cycles per instruction at
    N <= 64, whatever N and K are); folding to N = K-chunk = 128 does 2-4x the useful work per instruction with 3-5 folded
    taps instead of 3-11.  T must be a multiple of `fold`."""
    dtype = dtype or OPERAND_DTYPE
    w = weight.detach().float()
    cout, cin, k = w.shape
    p = (k - 1) // 2
    r_lo = (0 - p - (fold - 1)) // fold          # floor((g + s - f) / fold) over s in [-p, p], g, f in [0, fold)
    r_hi = (fold - 1 + p) // fold
    taps = []
    for r in range(r_lo, r_hi + 1):
        blk = torch.zeros(fold, cout, fold, cin, device=w.device)
        for g in range(fold):
            for f in range(fold):
                j = fold * r + f - g + p
                if 0 <= j < k:
                    blk[g, :, f, :] = w[:, :, j]
        taps.append(blk.reshape(fold * cout, fold * cin))
    # drop all-zero outer taps (possible when fold does not divide the reach evenly)
    while len(taps) > 1 and not taps[0].any():
        taps.pop(0)
        r_lo += 1
    while len(taps) > 1 and not taps[-1].any():
        taps.pop()
        r_hi -= 1
    w_ntc = torch.stack(taps, dim=1).contiguous()                      # [fold*cout, taps', fold*cin]
    b = None if bias is None else bias.detach().float().repeat(fold)
    ntaps = len(taps)
    return PackedWeight(_pack_taps(w_ntc, dtype), _bias(b, fold * cout, w.device), ntaps, fold * cin, fold * cout,
                        list(range(r_lo, r_hi + 1)), [0] * ntaps)


def pack_conv_transpose1d(weight, bias, stride, padding, dtype=None):
    """nn.ConvTranspose1d weight [cin, cout, k] -> one PackedWeight per output phase r in [0, stride).

    y[t] = sum_{i*s + j - p = t} x[i] w[:, :, j].  With t + p = s*q + r:  y[s*q + r - p] = sum_m x[q - m] w[r + s*m],
    i.e. phase r is a stride-1 conv over q with taps shifted by -m, written to output rows q*s + (r - p).
    """
    dtype = dtype or OPERAND_DTYPE
    w = weight.detach().float()
    cin, cout, k = w.shape
    phases = []
    for r in range(stride):
        js = list(range(r, k, stride))
        w_ntc = torch.stack([w[:, :, j].t() for j in js], dim=1).contiguous()  # [cout, ntaps, cin]
        d0 = [-m for m in range(len(js))]
        phases.append(PackedWeight(_pack_taps(w_ntc, dtype), _bias(bias, cout, w.device), len(js), cin, cout, d0,
                                   [0] * len(js), out_stride=stride, out_off=r - padding))
    return phases


class SingleChannelConv:
    """A conv with ONE output channel, restated as a pointwise GEMM over the taps + a shifted sum (see ctta_tap_sum)."""

    def __init__(self, pw, d0, d1, bias, fold, n_pad):
        self.pw = pw            # PackedWeight of the pointwise GEMM (block diagonal over `fold` consecutive pixels), no bias
        self.fold = fold        # pixels per GEMM row: [rows, C] is read as [rows / fold, fold * C]
        self.n_pad = n_pad      # taps zero padded to 8 / 16 = row pitch of z
        self.d0 = (C.c_int16 * len(d0))(*d0)
        self.d1 = (C.c_int16 * len(d1))(*d1)
        self.ntaps = len(d0)
        self.bias = bias        # fp32 [1] or None


def pack_single_channel_conv(weight, bias=None, dilation=1, dtype=None):
    """nn.Conv1d [1, C, k] / nn.Conv2d [1, C, kh, kw] weight ('same' padding, stride 1) -> SingleChannelConv."""
    w = weight.detach().float()
    assert w.shape[0] == 1
    if w.dim() == 3:
        k = w.shape[2]
        taps = w[0].t().contiguous()                                        # [k, C]
        d0 = [(j - (k - 1) // 2) * dilation for j in range(k)]
        d1 = [0] * k
    else:
        kh, kw = w.shape[2], w.shape[3]
        taps = w[0].permute(1, 2, 0).reshape(kh * kw, w.shape[1]).contiguous()  # [(kh, kw), C]
        d1 = [i - (kh - 1) // 2 for i in range(kh) for _ in range(kw)]
        d0 = [j - (kw - 1) // 2 for _ in range(kh) for j in range(kw)]
    n_pad = 8 if len(d0) <= 8 else 16
    assert len(d0) <= 16
    # `fold` consecutive pixels form one GEMM row (the same memory, read as [rows / fold, fold * C]) against a block
    # diagonal weight, so that N = fold * n_pad = 32 reaches the TMA-staged epilogue with M super-tiles: z[p, j] lands
    # at the same address, and a 128-row tile carries 4x / 2x the pixels for the same number of tensor-core instructions
    fold = 32 // n_pad
    c = taps.shape[1]
    tp = torch.zeros(n_pad, c, device=w.device)
    tp[:taps.shape[0]] = taps
    wf = torch.zeros(fold * n_pad, fold * c, device=w.device)
    for i in range(fold):
        wf[i * n_pad:(i + 1) * n_pad, i * c:(i + 1) * c] = tp
    pw = pack_linear(wf, None, dtype=dtype)
    return SingleChannelConv(pw, d0, d1, _bias(bias, 1, w.device), fold, n_pad)


# ------------------------------------------------------------------------------------------------ GEMM
def gemm(a, pw, mode=A_ROWS, n_img=1, h=1, w=None, rows_per_img=None, a_ld=None, out=None, out_ld=None,
         out_rows_per_img=None, residual=None, res_ld=None, rowadd=None, rowadd_rows=1, act=ACT_NONE, act_slope=0.0,
         accumulate=False, out_scale=1.0, out2=None, out2_ld=None, act2=ACT_NONE, act2_slope=0.0, use_bias=True,
         stats=None, stats_groups=32, stats_rows_per_img=0, res_neg_scale=1.0, wgt_img_stride=0, out_up_phase=0,
         stats_keep=False):
    """Launches ctta_gemm.  `a` is a 16-bit channels-last tensor; shapes are given explicitly by the caller.
    `stats` (fp32 [n_img, stats_groups, 2]) receives the GroupNorm moments of the result (fused statistics pass).
    A 16-bit `residual` may be the LeakyReLU'ed copy of the true residual: negative values are multiplied by
    `res_neg_scale` (= 1 / slope) before the add."""
    _require_cuda(a)
    d = GemmDesc()
    d.a = a.data_ptr()
    d.a_mode = mode
    d.ab_dtype = _DT[a.dtype]
    assert pw.w.dtype == a.dtype, "operand dtype mismatch"
    d.c = pw.c
    d.a_ld = a_ld if a_ld is not None else a.shape[-1]
    d.n_img = n_img
    d.h = h
    d.w = w if w is not None else (a.shape[0] if mode == A_ROWS else 0)
    if rows_per_img is None:
        rows_per_img = d.h * d.w
    d.rows_per_img = rows_per_img
    d.ntaps = pw.ntaps
    for j in range(pw.ntaps):
        d.tap_d0[j] = pw.d0[j]
        d.tap_d1[j] = pw.d1[j]
    d.wgt = pw.w.data_ptr()
    d.n = pw.n
    d.bias = pw.bias.data_ptr() if (use_bias and pw.bias is not None) else None
    if rowadd is not None:
        d.rowadd = rowadd.data_ptr()
        d.rowadd_ld = rowadd.stride(0)
        d.rowadd_rows = rowadd_rows
    d.act = act
    d.act_slope = act_slope
    if residual is not None:
        d.residual = residual.data_ptr()
        d.res_dtype = _DT[residual.dtype]
        d.res_ld = res_ld if res_ld is not None else residual.shape[-1]
        d.res_neg_scale = res_neg_scale
    d.accumulate = 1 if accumulate else 0
    d.out_scale = out_scale
    if out is not None:
        d.out = out.data_ptr()
        d.out_dtype = _DT[out.dtype]
        d.out_ld = out_ld if out_ld is not None else out.shape[-1]
    if out2 is not None:
        d.out2 = out2.data_ptr()
        d.out2_ld = out2_ld if out2_ld is not None else out2.shape[-1]
        d.act2 = act2
        d.act2_slope = act2_slope
    d.out_rows_per_img = out_rows_per_img if out_rows_per_img is not None else rows_per_img
    d.out_stride = pw.out_stride
    d.out_off = pw.out_off
    d.wgt_img_stride = wgt_img_stride
    d.out_up_phase = out_up_phase
    d.stats_keep = 1 if stats_keep else 0
    if stats is not None:
        d.stats = stats.data_ptr()
        d.stats_groups = stats_groups
        d.stats_rows_per_img = stats_rows_per_img
    check(lib().ctta_gemm(C.byref(d), _stream()))
    return out if out is not None else out2


def linear(a, pw, **kw):
    """a: [M, K] 16-bit."""
    return gemm(a, pw, mode=A_ROWS, n_img=1, h=1, w=a.shape[0], rows_per_img=a.shape[0], a_ld=a.stride(0), **kw)


def conv2d(a, pw, **kw):
    """a: [N, H, W, C] 16-bit channels-last; 'same' conv given by pw's taps."""
    n, h, w, _ = a.shape
    return gemm(a, pw, mode=A_CONV2D, n_img=n, h=h, w=w, rows_per_img=h * w, a_ld=a.stride(2), **kw)


def conv2d_upsample2x(a, phases, out, stats=None, stats_groups=32):
    """a: LOW-resolution 16-bit [N, H, W, C]; out: fp32 [N, 2H, 2W, Cout] = conv3x3(nearest_upsample_2x(a))."""
    n, h, w, _ = a.shape
    assert out.shape[1] == 2 * h and out.shape[2] == 2 * w and out.dtype == torch.float32 and out.is_contiguous()
    for i, (pw, code) in enumerate(phases):
        gemm(a, pw, mode=A_CONV2D, n_img=n, h=h, w=w, rows_per_img=h * w, a_ld=a.stride(2), out=out,
             out_rows_per_img=h * w, out_up_phase=code, stats=stats, stats_groups=stats_groups, stats_keep=i > 0)
    return out


def upsample2x_conv_supported(h, w):
    """The fused path needs the stream mainloop (tiles of full-width rows inside one image) and boxable warps."""
    return h * w >= 128 and w <= 128 and 128 % w == 0 and (w % 32 == 0 if w >= 32 else 32 % w == 0)


def conv1d(a, pw, rows_per_img=None, out_rows_per_img=None, **kw):
    """a: [B, T, C] 16-bit channels-last."""
    b, t, _ = a.shape
    rp = rows_per_img if rows_per_img is not None else t
    return gemm(a, pw, mode=A_CONV1D, n_img=b, h=1, w=t, rows_per_img=rp,
                out_rows_per_img=out_rows_per_img if out_rows_per_img is not None else rp, a_ld=a.stride(1), **kw)


def single_channel_conv(a, sc, n_img, h, w, act=ACT_NONE, out=None, out16=None):
    """a: 16-bit channels-last [n_img, h, w, C] (h = 1 for sequences).  Returns fp32 `out` [n_img * h * w] (and / or
    its 16-bit copy `out16`)."""
    rows = n_img * h * w
    assert a.is_contiguous() and rows % sc.fold == 0, "single_channel_conv: pixel count must be a multiple of the fold"
    a2 = a.reshape(rows // sc.fold, sc.fold * a.shape[-1])
    z = torch.empty(rows, sc.n_pad, device=a.device, dtype=torch.float32)
    linear(a2, sc.pw, out=z.view(rows // sc.fold, sc.fold * sc.n_pad))
    if out is None and out16 is None:
        out = torch.empty(rows, device=a.device, dtype=torch.float32)
    tap_sum(z, sc, n_img, h, w, act, out, out16)
    return out if out is not None else out16


def tap_sum(z, sc, n_img, h, w, act=ACT_NONE, out=None, out16=None):
    """y[p] = act(bias + sum_j z[p + shift_j, j]) over the taps of `sc` (zero outside each h x w image)."""
    check(lib().ctta_tap_sum(_ptr(z), sc.n_pad, n_img, h, w, sc.ntaps, sc.d0, sc.d1,
                             _ptr(sc.bias) if sc.bias is not None else None, act, _ptr(out) if out is not None else None,
                             _ptr(out16) if out16 is not None else None, _DT[out16.dtype] if out16 is not None else 0,
                             _stream()))


def bmm_nt(a, b, bias=None, out=None):
    """Batched out[i] = a[i] @ b[i]^T (+ bias): a [B, M, K], b [B, N, K] 16-bit (K % 64 == 0), out [B, M, N]."""
    bsz, m, k = a.shape
    n = b.shape[1]
    assert b.shape[0] == bsz and b.shape[2] == k and k % 64 == 0 and a.is_contiguous() and b.is_contiguous()
    pw = PackedWeight(b, bias, 1, k, n, [0], [0])
    return gemm(a, pw, mode=A_CONV1D, n_img=bsz, h=1, w=m, rows_per_img=m, out_rows_per_img=m, a_ld=k, out=out,
                wgt_img_stride=n * k)


def conv_transpose1d(a, phases, t_out, **kw):
    """a: [B, T, C]; phases from pack_conv_transpose1d; writes every output row exactly once."""
    b, t, _ = a.shape
    res = None
    for pw in phases:
        # q ranges over [0, ceil((t_out - out_off) / stride)) — rows mapping outside [0, t_out) are dropped
        q = (t_out - pw.out_off + pw.out_stride - 1) // pw.out_stride
        res = conv1d(a, pw, rows_per_img=max(q, 1), out_rows_per_img=t_out, **kw)
    return res


# ------------------------------------------------------------------------------------------------ norms
def groupnorm_stats(x, groups, x2=None, stats=None):
    """x: [N, H, W, C] (fp32 or 16-bit), optional x2 [N, H, W, C2] concatenated on channels -> stats [N, G, 2]."""
    _require_cuda(x)
    n, h, w, c = x.shape
    c2 = x2.shape[-1] if x2 is not None else 0
    if stats is None:
        stats = torch.empty(n, groups, 2, device=x.device, dtype=torch.float32)
    check(lib().ctta_groupnorm_stats(_ptr(x), _DT[x.dtype], c, x.stride(2), _ptr(x2), c2,
                                     x2.stride(2) if x2 is not None else 0, n, h * w, groups, _ptr(stats), _stream()))
    return stats


def groupnorm_apply(x, groups, stats, gamma, beta, eps=1e-5, act=ACT_SILU, x2=None, upsample=False, out=None, raw_out=None,
                    out_dtype=None):
    """Normalise (+SiLU) -> 16-bit operand [N, H', W', C+C2]; stats=None: plain cast/concat/upsample."""
    _require_cuda(x)
    n, h, w, c = x.shape
    c2 = x2.shape[-1] if x2 is not None else 0
    s = 2 if upsample else 1
    if out is None:
        out = torch.empty(n, h * s, w * s, c + c2, device=x.device, dtype=out_dtype or OPERAND_DTYPE)
    check(lib().ctta_groupnorm_apply(_ptr(x), _DT[x.dtype], c, x.stride(2), _ptr(x2), c2,
                                     x2.stride(2) if x2 is not None else 0, n, h, w, groups, _ptr(stats), _ptr(gamma),
                                     _ptr(beta), eps, act, 1 if upsample else 0, _ptr(out), _DT[out.dtype], out.stride(2),
                                     _ptr(raw_out), raw_out.stride(2) if raw_out is not None else 0, _stream()))
    return out


def layernorm(x, d, gamma, beta, eps, out=None):
    """x: fp32 [M, ld] (first d columns valid) -> 16-bit [M, ld] with pad columns zero."""
    _require_cuda(x)
    m, ld = x.shape
    if out is None:
        out = torch.empty(m, ld, device=x.device, dtype=OPERAND_DTYPE)
    check(lib().ctta_layernorm(_ptr(x), m, d, x.stride(0), _ptr(gamma), _ptr(beta), eps, _ptr(out), _DT[out.dtype],
                               out.stride(0), _stream()))
    return out


# ------------------------------------------------------------------------------------------------ attention
def attention(q, k, v, scale, kv_len=None, out=None):
    """q: [B, Lq, H, D], k/v: [B, Lk, H, D] 16-bit views (last dim contiguous, head stride = D)."""
    _require_cuda(q)
    b, lq, h, dh = q.shape
    lk = k.shape[1]
    if out is None:
        out = torch.empty(b, lq, h, dh, device=q.device, dtype=q.dtype)
    for t in (q, k, v, out):
        assert t.stride(3) == 1 and t.stride(2) == dh
    check(lib().ctta_attention(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _DT[q.dtype], b, h, lq, lk, dh,
                               q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1),
                               out.stride(0), out.stride(1), _ptr(kv_len), scale, _stream()))
    return out


def softmax_rows(x, scale, out=None):
    """x: fp32 [rows, cols] -> 16-bit softmax(scale * x) rows."""
    rows, cols = x.shape
    if out is None:
        out = torch.empty(rows, cols, device=x.device, dtype=OPERAND_DTYPE)
    check(lib().ctta_softmax_rows(_ptr(x), rows, cols, x.stride(0), scale, _ptr(out), _DT[out.dtype], out.stride(0),
                                  _stream()))
    return out


# ------------------------------------------------------------------------------------------------ misc
def im2col_s2(x):
    n, h, w, c = x.shape
    assert x.is_contiguous()
    out = torch.empty(n * (h // 2) * (w // 2), 9 * c, device=x.device, dtype=x.dtype)
    check(lib().ctta_im2col_s2(_ptr(x), n, h, w, c, _ptr(out), _stream()))
    return out


def nchw_to_nhwc(x, dtype=torch.float32, scale=1.0, c_pad=None, out=None, scale_dev=None):
    """fp32 NCHW -> channels-last [N, H, W, c_pad] (zero padded channels), multiplied by `scale` (and by the device
    scalar `scale_dev`, fp32 [1], when given)."""
    _require_cuda(x)
    x = x.contiguous().float()
    n, c, h, w = x.shape
    cp = c_pad or c
    if out is None:
        out = (torch.zeros if cp != c else torch.empty)(n, h, w, cp, device=x.device, dtype=dtype)
    dtype = out.dtype
    check(lib().ctta_nchw_to_nhwc(_ptr(x), n, c, h * w, _ptr(out), _DT[dtype], cp, scale, _ptr(scale_dev), _stream()))
    return out


def nhwc_to_nchw(x, c=None, out=None):
    """fp32 channels-last [N, H, W, ld] -> fp32 NCHW [N, c, H, W]."""
    n, h, w, ld = x.shape
    c = c or ld
    if out is None:
        out = torch.empty(n, c, h, w, device=x.device, dtype=torch.float32)
    check(lib().ctta_nhwc_to_nchw(_ptr(x), n, c, h * w, x.stride(2), _ptr(out), _stream()))
    return out


def time_features(t, w, gw):
    b = t.shape[0]
    tf = torch.empty(b, 256, device=t.device, dtype=torch.float32)
    gf = torch.empty(b, gw.shape[0] * 2, device=t.device, dtype=torch.float32)
    check(lib().ctta_time_features(_ptr(t), _ptr(w), _ptr(gw), b, _ptr(tf), _ptr(gf), _stream()))
    return tf, gf


def small_linear(x, weight, bias, act_in=ACT_NONE, act_out=ACT_NONE, accumulate=False, out=None):
    m, k = x.shape
    n = weight.shape[0]
    if out is None:
        out = torch.empty(m, n, device=x.device, dtype=torch.float32)
    check(lib().ctta_small_linear(_ptr(x), m, k, _ptr(weight), _ptr(bias), n, act_in, act_out, 1 if accumulate else 0,
                                  _ptr(out), _stream()))
    return out


def wave_to_int16(wav, minmax=None, out=None):
    """Batch-global centring + int16 truncation (hifigan/utilities.py:84-86). Returns (int16 tensor, minmax)."""
    numel = wav.numel()
    if minmax is None:
        minmax = torch.empty(2, device=wav.device, dtype=torch.float32)
        check(lib().ctta_wave_minmax(_ptr(wav), numel, _ptr(minmax), _stream()))
    if out is None:
        out = torch.empty(wav.shape, device=wav.device, dtype=torch.int16)
    check(lib().ctta_wave_to_int16(_ptr(wav), numel, _ptr(minmax), _ptr(out), _stream()))
    return out, minmax


def lrelu_cast(x, slope, out=None):
    """fp32 tensor -> 16-bit leaky_relu(x, slope)."""
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=OPERAND_DTYPE)
    check(lib().ctta_lrelu_cast(_ptr(x), x.numel(), slope, _ptr(out), _DT[out.dtype], _stream()))
    return out


def resblock_pair_supported(c, taps, dilation, t=2):
    return bool(lib().ctta_resblock_pair_supported(c, taps, dilation, t))


def resblock_pair(lx, pw1, pw2, slope, out=None):
    """One fused (c1, c2) pair of a HiFi-GAN ResBlock (hifigan/models.py:56-63) on the LeakyReLU'ed 16-bit stream:
    lx = lrelu(x) [B, T, C] -> lrelu(x + c2(lrelu(c1(lrelu(x))))) [B, T, C].  pw1 / pw2 from pack_conv1d (pw1 dilated)."""
    _require_cuda(lx)
    b, t, c = lx.shape
    assert lx.is_contiguous() and pw1.ntaps == pw2.ntaps and pw1.n == c and pw2.n == c and pw1.c == c and pw2.c == c
    assert pw1.w.dtype == lx.dtype and pw2.w.dtype == lx.dtype and pw1.bias is not None and pw2.bias is not None
    taps = pw1.ntaps
    dil = (pw1.d0[1] - pw1.d0[0]) if taps > 1 else 1
    assert (pw2.d0[1] - pw2.d0[0] if taps > 1 else 1) == 1, "c2 is the undilated conv"
    if out is None:
        out = torch.empty_like(lx)
    assert out.is_contiguous() and out.shape == lx.shape and out.dtype == lx.dtype and out.data_ptr() != lx.data_ptr()
    check(lib().ctta_resblock_pair(_ptr(lx), _ptr(out), _DT[lx.dtype], b, t, c, _ptr(pw1.w), _ptr(pw1.bias), _ptr(pw2.w),
                                   _ptr(pw2.bias), taps, dil, slope, _stream()))
    return out


def mrf_combine(xs, in_slope, out_scale, out_slope, out=None):
    """HiFi-GAN MRF sum over the LeakyReLU'ed 16-bit ResBlock outputs `xs` -> LeakyReLU'ed 16-bit operand of the next stage."""
    x0 = xs[0]
    _require_cuda(x0)
    if out is None:
        out = torch.empty_like(x0)
    ptrs = (C.c_void_p * len(xs))(*[x.data_ptr() for x in xs])
    for x in xs:
        assert x.shape == x0.shape and x.dtype == x0.dtype and x.is_contiguous()
    check(lib().ctta_mrf_combine(ptrs, len(xs), x0.numel(), _DT[x0.dtype], in_slope, out_scale, out_slope, _ptr(out),
                                 _stream()))
    return out


def cfg_mix(x, s, out=None):
    half = x.shape[0] // 2
    if out is None:
        out = torch.empty((half,) + tuple(x.shape[1:]), device=x.device, dtype=torch.float32)
    check(lib().ctta_cfg_mix(_ptr(x), out.numel(), s, _ptr(out), _stream()))
    return out


def launch_count():
    return int(lib().ctta_launch_count())
