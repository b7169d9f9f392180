"""Drop-in AutoencoderKL (decode side) + HiFi-GAN Generator running on libctta kernels.

Mirrors audioldm/variational_autoencoder/autoencoder.py:9-111 (AutoencoderKL: decode_first_stage,
decode_to_waveform, .decoder/.post_quant_conv/.vocoder/.scale_factor), modules.py:546-683 (Decoder) and
audioldm/hifigan/models.py:66-117 (Generator) with the reference's state_dict keys.  The encode side
(encoder.*, quant_conv.*) is training/eval only and out of scope: its keys are accepted and ignored.
"""
import os

import torch
from torch import nn

from . import ops, weights
from .module_base import PackedModule, _Node, next_pack_version, register_tree
from .ops import ACT_LRELU, ACT_NONE, ACT_SILU, ACT_TANH

GROUPS = 32
EPS = 1e-6  # Normalize(), modules.py:38-41


class Generator(PackedModule):
    """HiFi-GAN V1 generator, HIFIGAN_16K_64 (hifigan/utilities.py:9-39). forward(mel [B, 64, T]) -> [B, 1, T*160+32]."""

    def __init__(self, h=None):
        super().__init__()
        self.h = dict(weights.HIFIGAN_CONFIG)
        if h is not None:
            self.h.update({k: h[k] for k in weights.HIFIGAN_CONFIG if k in h})
        self.num_kernels = len(self.h["resblock_kernel_sizes"])
        self.num_upsamples = len(self.h["upsample_rates"])
        # Fused ResBlock-pair kernel on the C <= 64 stages (ctta_resblock_pair) instead of two ctta_gemm launches per pair:
        # bit-identical output, 9-40 % less time per pair (profiles/r2_resblock_pair.txt).  CTTA_FUSE_PAIRS=0 switches back
        # to the two-launch path (A/B runs).
        self.fuse_pairs = os.environ.get("CTTA_FUSE_PAIRS", "1") not in ("", "0")
        register_tree(self, weights.vocoder_schema(self.h, prefix=""))

    def remove_weight_norm(self):
        """Weights are stored folded (hifigan/utilities.py:71); nothing to do."""

    def _pack(self, sd, dev):
        pk = {}
        sd = {k: v.to(dev) for k, v in sd.items()}
        pk["conv_pre"] = ops.pack_conv1d(sd["conv_pre.weight"], sd["conv_pre.bias"])
        for i, (u, k) in enumerate(zip(self.h["upsample_rates"], self.h["upsample_kernel_sizes"])):
            pk["ups.%d" % i] = ops.pack_conv_transpose1d(sd["ups.%d.weight" % i], sd["ups.%d.bias" % i], u, (k - u) // 2)
            for j, (ks, dils) in enumerate(zip(self.h["resblock_kernel_sizes"], self.h["resblock_dilation_sizes"])):
                n = i * self.num_kernels + j
                for m, d in enumerate(dils):
                    pk["rb.%d.c1.%d" % (n, m)] = ops.pack_conv1d(sd["resblocks.%d.convs1.%d.weight" % (n, m)],
                                                                 sd["resblocks.%d.convs1.%d.bias" % (n, m)], dilation=d)
                    pk["rb.%d.c2.%d" % (n, m)] = ops.pack_conv1d(sd["resblocks.%d.convs2.%d.weight" % (n, m)],
                                                                 sd["resblocks.%d.convs2.%d.bias" % (n, m)], dilation=1)
                    # narrow stages: dilation-1 convs also in the time-folded form (ops.pack_conv1d_folded)
                    ch = sd["resblocks.%d.convs2.%d.weight" % (n, m)].shape[0]
                    fold = self.fold_factor(ch, ks)
                    if fold > 1:
                        pk["rb.%d.c2.%d.fold" % (n, m)] = ops.pack_conv1d_folded(
                            sd["resblocks.%d.convs2.%d.weight" % (n, m)], sd["resblocks.%d.convs2.%d.bias" % (n, m)], fold)
                        if d == 1:
                            pk["rb.%d.c1.%d.fold" % (n, m)] = ops.pack_conv1d_folded(
                                sd["resblocks.%d.convs1.%d.weight" % (n, m)], sd["resblocks.%d.convs1.%d.bias" % (n, m)], fold)
        # one output channel: pointwise GEMM over the 7 taps + shifted sum (ops.single_channel_conv)
        pk["conv_post"] = ops.pack_single_channel_conv(sd["conv_post.weight"], sd["conv_post.bias"])
        return pk

    @staticmethod
    def fold_factor(ch, ks):
        """Time steps folded into one GEMM row for a dilation-1 conv of `ch` channels and `ks` taps (1 = not folded):
        C = 32 -> 4 (N = K-chunk = 128, 3 / 3 / 5 folded taps for k = 3 / 7 / 11), C = 64 -> 2 for k <= 7 (3 / 5 folded
        taps; k = 11 would need 7 and measures slower).  OPT-IN (CTTA_FOLD=1): exact and 15-25 % faster per launch in the
        eager profile, but the captured pipeline measures the same clips/s (369.4 vs 369.1, profiles/r2_fold.txt) — the
        folded problems run like the C = 128 stage, which is bound by its 16-bit epilogue / TMA rows, not by the MMA floor."""
        if os.environ.get("CTTA_FOLD") is None:
            return 1
        if ch == 32:
            return 4
        if ch == 64 and ks <= 7:
            return 2
        return 1

    def out_length(self, t):
        for u, k in zip(self.h["upsample_rates"], self.h["upsample_kernel_sizes"]):
            t = (t - 1) * u - 2 * ((k - u) // 2) + k
        return t

    def forward_btc(self, mel16):
        """mel16: 16-bit channels-last [B, T, 64] -> fp32 waveform [B, T_out].

        Every activation of models.py:56-63,101-117 lives in HBM only as its LeakyReLU'ed 16-bit copy lx = lrelu(x, 0.1):
        it is the operand of the next conv AND (LeakyReLU being invertible: x = lx >= 0 ? lx : 10 lx) the residual of
        the ResBlock add, which the conv epilogue undoes on load.  So a ResBlock pair costs 2+2 (c1) + 2+2+2 (c2) bytes
        per element instead of 16 with a separate fp32 residual stream, and the MRF sum / divide-by-3 / next LeakyReLU is
        one pass over the three 16-bit branch outputs."""
        pk = self.packed()
        dev = mel16.device
        f16 = ops.OPERAND_DTYPE
        b, t, _ = mel16.shape
        c = self.h["upsample_initial_channel"]
        cur16 = torch.empty(b, t, c, device=dev, dtype=f16)
        ops.conv1d(mel16, pk["conv_pre"], out2=cur16, act2=ACT_LRELU, act2_slope=0.1)  # models.py:102,104
        nk = self.num_kernels
        slope = 0.1
        for i, (u, k) in enumerate(zip(self.h["upsample_rates"], self.h["upsample_kernel_sizes"])):
            c //= 2
            t = (t - 1) * u - 2 * ((k - u) // 2) + k
            lx = torch.empty(b, t, c, device=dev, dtype=f16)
            ops.conv_transpose1d(cur16, pk["ups.%d" % i], t, out2=lx, act2=ACT_LRELU, act2_slope=slope)
            tmp16 = torch.empty(b, t, c, device=dev, dtype=f16)
            ping = [torch.empty(b, t, c, device=dev, dtype=f16) for _ in range(2)]
            branch = [torch.empty(b, t, c, device=dev, dtype=f16) for _ in range(nk)]
            last = i == self.num_upsamples - 1
            for j in range(nk):
                n = i * nk + j
                cur = lx
                for m in range(3):
                    dst = branch[j] if m == 2 else ping[m]
                    pw1, pw2 = pk["rb.%d.c1.%d" % (n, m)], pk["rb.%d.c2.%d" % (n, m)]
                    if self.fuse_pairs and ops.resblock_pair_supported(c, pw1.ntaps, pw1.d0[1] - pw1.d0[0], t):
                        # narrow stages (C <= 64): c1 -> lrelu -> c2 -> + x in ONE kernel, hidden tile in shared memory
                        ops.resblock_pair(cur, pw1, pw2, slope, out=dst)
                        cur = dst
                        continue
                    f1, f2 = pk.get("rb.%d.c1.%d.fold" % (n, m)), pk.get("rb.%d.c2.%d.fold" % (n, m))
                    fold = (f2.n // c) if f2 is not None else 1
                    if fold > 1 and t % fold != 0:
                        f1 = f2 = None

                    def fv(x):   # [B, T, C] read as [B, T / fold, fold * C]: the same memory
                        return x.view(b, t // fold, fold * c)
                    if f1 is not None:
                        ops.conv1d(fv(cur), f1, out2=fv(tmp16), act2=ACT_LRELU, act2_slope=slope)
                    else:
                        ops.conv1d(cur, pw1, out2=tmp16, act2=ACT_LRELU, act2_slope=slope)
                    # x' = x + c2(lrelu(c1(lrelu(x)))) with x recovered from lrelu(x); emitted again as lrelu(x')
                    if f2 is not None:
                        ops.conv1d(fv(tmp16), f2, residual=fv(cur), res_neg_scale=1.0 / slope, out2=fv(dst),
                                   act2=ACT_LRELU, act2_slope=slope)
                    else:
                        ops.conv1d(tmp16, pw2, residual=cur, res_neg_scale=1.0 / slope, out2=dst,
                                   act2=ACT_LRELU, act2_slope=slope)
                    cur = dst
            # x = (sum_j resblock_j(x)) / 3 (models.py:108-112), then LeakyReLU of models.py:104 / :113 (default slope)
            cur16 = ops.mrf_combine(branch, slope, 1.0 / nk, 0.01 if last else 0.1, out=tmp16)
        wav = torch.empty(b, t, device=dev, dtype=torch.float32)
        ops.single_channel_conv(cur16, pk["conv_post"], b, 1, t, act=ACT_TANH, out=wav)  # models.py:113-115
        return wav

    def forward(self, x):
        """Reference signature: x [B, num_mels, T] fp32 -> [B, 1, T_out]."""
        mel16 = x.transpose(1, 2).contiguous().to(ops.OPERAND_DTYPE)
        return self.forward_btc(mel16).unsqueeze(1)


class Decoder(PackedModule):
    """LDM-style conv decoder 8x256x16 -> 1x1024x64, modules.py:546-683."""

    def __init__(self, **ddconfig):
        super().__init__()
        cfg = {"embed_dim": weights.VAE_CONFIG["embed_dim"], "ddconfig": dict(weights.VAE_CONFIG["ddconfig"])}
        cfg["ddconfig"].update({k: v for k, v in ddconfig.items() if k in cfg["ddconfig"]})
        if cfg["ddconfig"] != weights.VAE_CONFIG["ddconfig"]:
            raise ValueError("only the audioldm-s-full first-stage decoder is implemented")
        sch = {k[len("decoder."):]: v for k, v in weights.vae_decoder_schema(cfg).items() if k.startswith("decoder.")}
        register_tree(self, sch)
        # levels whose residual stream is held in the 16-bit operand type (0 = 128 ch at 1024 x 64, 1 = 256 ch at 512 x 32):
        # there the stream IS the HBM traffic (fp32: 2.1 / 1.1 GB per tensor at B = 64).  CTTA_VAE_STREAM16=0: all fp32.
        self.stream16_levels = () if os.environ.get("CTTA_VAE_STREAM16") == "0" else (0, 1)

    def _pack(self, sd, dev):
        pk = {}
        sd = {k: v.to(dev) for k, v in sd.items()}
        for k in sd:
            if k.endswith(".weight"):
                name = k[:-7]
                w = sd[k]
                if w.dim() == 4 and w.shape[0] == 1:   # conv_out 128 -> 1: pointwise GEMM over the 9 taps + shifted sum
                    pk[name] = ops.pack_single_channel_conv(w, sd[name + ".bias"])
                elif w.dim() == 4:
                    pk[name] = ops.pack_conv2d(w, sd[name + ".bias"])
                    if name.endswith(".upsample.conv"):
                        pk[name + ".up2x"] = ops.pack_upsample2x_conv2d(w, sd[name + ".bias"])
                else:
                    pk[name] = (w.float().contiguous(), sd[name + ".bias"].float().contiguous())
        return pk

    def _resnet(self, pk, p, x, x_stats, stream16=False):
        """ResnetBlock.forward (temb None), modules.py:155-175.  x_stats: GroupNorm moments of x emitted by the kernel
        that wrote x.  Returns (out, moments of out).  `stream16`: the block's output (the residual stream) is written in
        the 16-bit operand type instead of fp32 — used at the two high-resolution levels, where the stream is the HBM
        traffic (x may arrive in either type; the accumulation itself stays fp32 in the epilogue)."""
        b, h, w, cin = x.shape
        dev = x.device
        has_sc = (p + ".nin_shortcut") in pk
        raw = torch.empty(b, h, w, cin, device=dev, dtype=ops.OPERAND_DTYPE) if has_sc else None
        a = ops.groupnorm_apply(x, GROUPS, x_stats, *pk[p + ".norm1"], eps=EPS, act=ACT_SILU, raw_out=raw)
        cout = pk[p + ".conv1"].n
        # 16-bit hidden tensor (only GroupNorm 2 reads it); its moments come out of the conv epilogue
        hid = torch.empty(b, h, w, cout, device=dev, dtype=ops.OPERAND_DTYPE)
        st2 = torch.empty(b, GROUPS, 2, device=dev, dtype=torch.float32)
        ops.conv2d(a, pk[p + ".conv1"], out=hid, stats=st2, stats_groups=GROUPS)
        del a
        a2 = ops.groupnorm_apply(hid, GROUPS, st2, *pk[p + ".norm2"], eps=EPS, act=ACT_SILU)
        del hid
        sdt = ops.OPERAND_DTYPE if stream16 else torch.float32
        if has_sc:
            res = torch.empty(b, h, w, cout, device=dev, dtype=sdt)
            ops.conv2d(raw, pk[p + ".nin_shortcut"], out=res)
        else:
            res = x
        out = torch.empty(b, h, w, cout, device=dev, dtype=sdt)
        out_stats = torch.empty(b, GROUPS, 2, device=dev, dtype=torch.float32)
        ops.conv2d(a2, pk[p + ".conv2"], out=out, residual=res, stats=out_stats, stats_groups=GROUPS)
        return out, out_stats

    def _attn(self, pk, p, x, x_stats):
        """AttnBlock.forward, modules.py:204-230: single head over H*W tokens, d = C = 512.
        scores and P.V run on the tcgen05 GEMM per sample; V's bias is added after P.V (softmax rows sum to 1)."""
        b, h, w, c = x.shape
        dev = x.device
        f16 = ops.OPERAND_DTYPE
        n = h * w
        a = ops.groupnorm_apply(x, GROUPS, x_stats, *pk[p + ".norm"], eps=EPS, act=ACT_NONE).view(b, n, c)
        q = torch.empty(b, n, c, device=dev, dtype=f16)
        k = torch.empty(b, n, c, device=dev, dtype=f16)
        ops.linear(a.view(b * n, c), pk[p + ".q"], out=q.view(b * n, c))
        ops.linear(a.view(b * n, c), pk[p + ".k"], out=k.view(b * n, c))
        wv = pk[p + ".v"]
        # all samples at once (batched GEMMs: image i multiplies its own K-major operand):
        #   V^T[i] = W_v . a[i]^T,   S[i] = q[i] . k[i]^T (fp32),   P = softmax(S / sqrt(c)),   o[i] = P[i] . V[i] + b_v
        # W_v replicated per sample: the A operand of the batched V^T GEMM.  One entry PER BATCH SIZE, never replaced:
        # captured CUDA graphs hold the raw pointer, and a freed entry would be recycled by the caching allocator
        wv_rep = pk.get((p, "v_rep", b))
        if wv_rep is None:
            wv_rep = pk[(p, "v_rep", b)] = wv.w.unsqueeze(0).expand(b, c, c).contiguous()
        vt = torch.empty(b, c, n, device=dev, dtype=f16)
        ops.bmm_nt(wv_rep, a, out=vt)
        s = torch.empty(b, n, n, device=dev, dtype=torch.float32)
        ops.bmm_nt(q, k, out=s)
        pr = torch.empty(b, n, n, device=dev, dtype=f16)
        ops.softmax_rows(s.view(b * n, n), float(c) ** -0.5, out=pr.view(b * n, n))
        del s
        o = torch.empty(b, n, c, device=dev, dtype=f16)
        ops.bmm_nt(pr, vt, bias=wv.bias, out=o)
        del pr
        out = torch.empty(b, h, w, c, device=dev, dtype=torch.float32)
        out_stats = torch.empty(b, GROUPS, 2, device=dev, dtype=torch.float32)
        ops.linear(o.view(b * n, c), pk[p + ".proj_out"], out=out.view(b * n, c), residual=x.view(b * n, c),
                   stats=out_stats, stats_groups=GROUPS, stats_rows_per_img=n)
        return out, out_stats

    def forward_nhwc(self, z16, out16=None):
        """z16: 16-bit channels-last [B, 256, 16, 8] (after post_quant_conv) -> fp32 [B, 1024, 64, 1]."""
        pk = self.packed()
        dev = z16.device
        b, h, w, _ = z16.shape
        x = torch.empty(b, h, w, pk["conv_in"].n, device=dev, dtype=torch.float32)
        st = torch.empty(b, GROUPS, 2, device=dev, dtype=torch.float32)
        ops.conv2d(z16, pk["conv_in"], out=x, stats=st, stats_groups=GROUPS)
        x, st = self._resnet(pk, "mid.block_1", x, st)
        x, st = self._attn(pk, "mid.attn_1", x, st)
        x, st = self._resnet(pk, "mid.block_2", x, st)
        for lvl in (2, 1, 0):
            for blk in range(3):
                x, st = self._resnet(pk, "up.%d.block.%d" % (lvl, blk), x, st, stream16=lvl in self.stream16_levels)
            if lvl != 0:  # Upsample, modules.py:53-57
                bb, hh, ww, cc = x.shape
                # nearest 2x + conv3x3 as four 2x2 phase convs on the low-resolution tensor (4/9 of the FLOPs)
                x16 = x if x.dtype != torch.float32 else ops.groupnorm_apply(x, 1, None, None, None, act=ACT_NONE)
                del x
                x = torch.empty(bb, 2 * hh, 2 * ww, cc, device=dev, dtype=torch.float32)
                st = torch.empty(b, GROUPS, 2, device=dev, dtype=torch.float32)
                ops.conv2d_upsample2x(x16, pk["up.%d.upsample.conv.up2x" % lvl], x, stats=st, stats_groups=GROUPS)
                del x16
        a = ops.groupnorm_apply(x, GROUPS, st, *pk["norm_out"], eps=EPS, act=ACT_SILU)
        bb, hh, ww, _ = x.shape
        del x
        out = torch.empty(bb, hh, ww, 1, device=dev, dtype=torch.float32)
        ops.single_channel_conv(a, pk["conv_out"], bb, hh, ww, out=out, out16=out16)  # modules.py:680
        return out

    def forward(self, z):
        """Reference signature: z fp32 NCHW [B, 8, 256, 16] -> [B, 1, 1024, 64]."""
        z16 = ops.nchw_to_nhwc(z.float(), dtype=ops.OPERAND_DTYPE)
        out = self.forward_nhwc(z16)
        b, h, w, _ = out.shape
        return out.view(b, 1, h, w)


class AutoencoderKL(nn.Module):
    """autoencoder.py:9-111, decode side."""

    def __init__(self, ddconfig=None, lossconfig=None, image_key="fbank", embed_dim=None, time_shuffle=1, subband=1,
                 ckpt_path=None, reload_from_ckpt=None, ignore_keys=(), colorize_nlabels=None, monitor=None,
                 base_learning_rate=1e-5, scale_factor=1):
        super().__init__()
        ddconfig = ddconfig or weights.VAE_CONFIG["ddconfig"]
        embed_dim = embed_dim or weights.VAE_CONFIG["embed_dim"]
        self.decoder = Decoder(**ddconfig)
        self.ema_decoder = None
        self.subband = int(subband)
        if self.subband != 1:
            raise NotImplementedError("subband decomposition is not used by ConsistencyTTA")
        self.post_quant_conv = _Node()
        z = ddconfig["z_channels"]
        self.post_quant_conv.register_parameter("weight", nn.Parameter(torch.zeros(z, embed_dim, 1, 1), False))
        self.post_quant_conv.register_parameter("bias", nn.Parameter(torch.zeros(z), False))
        self.ema_post_quant_conv = None
        self.vocoder = Generator()
        self.embed_dim = embed_dim
        self.image_key = image_key
        self.time_shuffle = time_shuffle
        self.scale_factor = scale_factor
        self._pq = {}
        self._pq_version = next_pack_version()
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module._drop_post_quant())

    @property
    def device(self):
        return next(self.parameters()).device

    def load_state_dict(self, state_dict, strict=True, **kw):
        """Accepts the reference's 398-key VAE state_dict; encode-side tensors are dropped (out of scope).  A state_dict
        that carries `ema_decoder.*` / `ema_post_quant_conv.*` (a VAE saved after fine-tuning with
        models/audio_consistency_model_ftvae.py:52-65) creates the EMA modules before loading."""
        sd = {k: v for k, v in state_dict.items() if not k.startswith(("encoder.", "quant_conv.", "loss."))}
        if self.ema_decoder is None and any(k.startswith(("ema_decoder.", "ema_post_quant_conv.")) for k in sd):
            self.enable_ema_modules()
        self._drop_post_quant()
        return super().load_state_dict(sd, strict=strict, **kw)

    def enable_ema_modules(self):
        """EMA copies of the decoder and post_quant_conv (audio_consistency_model_ftvae.py:52-65), selected by
        `decode_first_stage(..., use_ema=True)` (autoencoder.py:91-97)."""
        from copy import deepcopy
        dev = self.device
        self.ema_decoder = deepcopy(self.decoder)
        self.ema_decoder._invalidate()
        self.ema_decoder.requires_grad_(False).eval()
        self.ema_post_quant_conv = deepcopy(self.post_quant_conv)
        self.ema_post_quant_conv.requires_grad_(False)
        self._drop_post_quant()
        return self.to(dev)

    @classmethod
    def from_vae_state_dict_file(cls, path_or_obj, **kw):
        """`consistencytta_clapft_ckpt/vae_state_dict.pt` = {"state_dict", "scale_factor"} (consistencytta.py:32-44)."""
        raw = torch.load(path_or_obj, map_location="cpu") if isinstance(path_or_obj, (str, bytes)) or hasattr(
            path_or_obj, "read") else path_or_obj
        sf = raw["scale_factor"]
        vae = cls(scale_factor=float(sf.item() if torch.is_tensor(sf) else sf), **kw)
        vae.load_state_dict(raw["state_dict"])
        vae.eval().requires_grad_(False)
        return vae

    @classmethod
    def from_audioldm_checkpoint(cls, path_or_obj, **kw):
        """An AudioLDM checkpoint {"state_dict": {"first_stage_model.*", "scale_factor", ...}}: keeps the first-stage
        tensors with the 18-character prefix removed and reads the latent scale (tools/build_pretrained.py:9-22)."""
        ckpt = torch.load(path_or_obj, map_location="cpu") if isinstance(path_or_obj, (str, bytes)) or hasattr(
            path_or_obj, "read") else path_or_obj
        sd = ckpt["state_dict"]
        sf = sd["scale_factor"]
        vae_sd = {k[len("first_stage_model."):]: v for k, v in sd.items() if "first_stage_model." in k}
        vae = cls(scale_factor=float(sf.item() if torch.is_tensor(sf) else sf), **kw)
        vae.load_state_dict(vae_sd)
        vae.eval().requires_grad_(False)
        return vae

    def _drop_post_quant(self):
        self._pq = {}
        self._pq_version = next_pack_version()

    def _apply(self, fn, *a, **k):
        self._drop_post_quant()
        return super()._apply(fn, *a, **k)

    @property
    def pack_version(self):
        """Changes whenever any packed operand of the decode path may have been rebuilt (see PackedModule)."""
        mods = [self.decoder, self.vocoder] + ([self.ema_decoder] if self.ema_decoder is not None else [])
        return (self._pq_version,) + tuple(m.pack_version for m in mods)

    def to_prepacked(self, device):
        """`.to(device)` with the decoder / vocoder operands packed on the host first (PackedModule.to_prepacked)."""
        self.to(device)
        for m in (self.decoder, self.vocoder, self.ema_decoder):
            if m is not None:
                m.to_prepacked(device)
        return self

    def encode(self, x):
        raise NotImplementedError("the VAE encode side is training/evaluation only (SURVEY.md 8a V1)")

    encode_first_stage = encode

    def _post_quant(self, use_ema, z_scale):
        """post_quant_conv 1x1 (autoencoder.py:99) with `z / scale_factor` (:105) folded into its weights."""
        mod = self.post_quant_conv
        if use_ema and self.ema_post_quant_conv is not None:
            mod = self.ema_post_quant_conv
        key = ("ema" if mod is self.ema_post_quant_conv else "raw", float(z_scale), str(mod.weight.device), ops.OPERAND_DTYPE)
        if key not in self._pq:   # one entry per key, kept alive: captured graphs hold the pointer
            self._pq[key] = ops.pack_conv2d(mod.weight.detach().float() * float(z_scale), mod.bias.detach())
        return self._pq[key]

    def decode_nhwc(self, z_nhwc, use_ema=False, z_scale=1.0, mel16=None):
        """z_nhwc: fp32 channels-last latent [B, 256, 16, 8] -> fp32 [B, 1024, 64, 1] (== NCHW [B,1,1024,64]).
        Computes Decoder(post_quant_conv(z * z_scale)); mel16 (optional, [B,1024,64,1] 16-bit) gets a 16-bit copy."""
        if use_ema and self.ema_decoder is None:
            print("VAE does not have EMA modules, but specified use_ema. Using the none-EMA modules instead.")
        dec = self.ema_decoder if (use_ema and self.ema_decoder is not None) else self.decoder
        b, h, w, c = z_nhwc.shape
        z16 = ops.groupnorm_apply(z_nhwc, 1, None, None, None, act=ACT_NONE)
        zq = torch.empty(b, h, w, c, device=z_nhwc.device, dtype=ops.OPERAND_DTYPE)
        ops.conv2d(z16, self._post_quant(use_ema, z_scale), out=zq)
        return dec.forward_nhwc(zq, out16=mel16)

    def decode(self, z, use_ema=False):
        zn = ops.nchw_to_nhwc(z.float(), dtype=torch.float32)
        out = self.decode_nhwc(zn, use_ema)
        b, h, w, _ = out.shape
        return out.view(b, 1, h, w)

    def decode_first_stage(self, z, allow_grad=False, use_ema=False):
        """autoencoder.py:103-106."""
        if allow_grad:
            raise NotImplementedError("allow_grad=True (training losses) is out of scope (SURVEY.md 8f N4)")
        with torch.no_grad():
            zn = ops.nchw_to_nhwc(z.float(), dtype=torch.float32)
            out = self.decode_nhwc(zn, use_ema, z_scale=1.0 / float(self.scale_factor))
            b, h, w, _ = out.shape
            return out.view(b, 1, h, w)

    def waveform_from_mel_nhwc(self, mel16):
        """mel16 16-bit [B, T, 64] -> (int16 [B, T_out] tensor on device, fp32 waveform)."""
        wav = self.vocoder.forward_btc(mel16)
        i16, _ = ops.wave_to_int16(wav)
        return i16, wav

    def decode_to_waveform(self, dec, allow_grad=False):
        """autoencoder.py:108-111 + vocoder_infer (hifigan/utilities.py:76-91) -> numpy int16 [B, 163872]."""
        if allow_grad:
            raise NotImplementedError("allow_grad=True (training losses) is out of scope (SURVEY.md 8f N4)")
        with torch.no_grad():
            b, _, t, m = dec.shape
            # [B,1,T,mel].squeeze(1).permute(0,2,1) is [B,mel,T]; channels-last [B,T,mel] is dec's own memory layout
            mel16 = ops.groupnorm_apply(dec.float().contiguous().view(b, t, 1, m), 1, None, None, None, act=ACT_NONE)
            i16, _ = self.waveform_from_mel_nhwc(mel16.view(b, t, m))
            return i16.cpu().numpy()
