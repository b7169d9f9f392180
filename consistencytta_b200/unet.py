"""Drop-in UNet2DConditionGuidedModel (tango_diffusion_light) running on libctta kernels.

Mirrors the reference's module API — diffusers/models/unet_2d_condition_guided.py:137-302 (constructor/config),
:716-945 (forward) — and its 691-tensor state_dict layout, but the forward is a flat sequence of C-ABI kernel
calls over channels-last activations:

  * residual stream fp32 [B, H, W, C] == token matrix [B*H*W, C]  (the reference's NCHW<->token permutes vanish)
  * every conv / linear is ctta_gemm (tcgen05) with fp16 operands and fp32 accumulation; bias, time-embedding add,
    residual add and GEGLU are GEMM epilogues
  * inner dims 255/510/1020 are zero-padded to 256/512/1024 and the 51-wide heads to 64 at weight-pack time
  * text K/V of all 16 cross-attention sites come from ONE GEMM over the prompt embeddings; the 22 time_emb_proj
    Linear layers are ONE GEMM
"""
import json
from collections import OrderedDict

import torch

from . import ops, weights
from .module_base import PackedModule, register_tree
from .ops import ACT_GEGLU, ACT_NONE, ACT_SILU

HEAD_PAD = 64


def prefix_mask_lengths(mask):
    """bool / 0-1 mask [B, L] with True only on a prefix of each row -> int32 [B] counts; anything else raises
    (the attention kernels take a key COUNT per row; an interior hole would be silently mis-applied)."""
    m = mask.to(torch.bool)
    if m.dim() != 2:
        raise ValueError("encoder_attention_mask must be [batch, text_len], got %s" % (tuple(m.shape),))
    if m.shape[1] > 1 and bool((~m[:, :-1] & m[:, 1:]).any()):
        raise ValueError("encoder_attention_mask must be a prefix mask (trailing padding only): a True after a False "
                         "cannot be expressed as a key count")
    return m.sum(dim=1).to(torch.int32)


class FrozenConfig(dict):
    """dict with attribute access, like diffusers' FrozenDict (configuration_utils.py)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


class UNet2DConditionOutput:
    """diffusers/models/unet_2d_condition_guided.py:42-50."""

    def __init__(self, sample):
        self.sample = sample

    def __getitem__(self, i):
        return (self.sample,)[i]


def _pad2(w, n, k):
    out = torch.zeros(n, k, device=w.device, dtype=torch.float32)
    out[: w.shape[0], : w.shape[1]] = w
    return out


def _pad1(b, n):
    out = torch.zeros(n, device=b.device, dtype=torch.float32)
    out[: b.shape[0]] = b
    return out


def _head_rows(w, heads, d, k_pad):
    """[heads*d, k] -> [heads*64, k_pad]: head h occupies rows [64h, 64h + d)."""
    k = w.shape[1]
    out = torch.zeros(heads, HEAD_PAD, k_pad, device=w.device, dtype=torch.float32)
    out[:, :d, :k] = w.reshape(heads, d, k)
    return out.reshape(heads * HEAD_PAD, k_pad)


def _head_cols(w, heads, d, n_pad):
    """[n, heads*d] -> [n_pad, heads*64]."""
    n = w.shape[0]
    out = torch.zeros(n_pad, heads, HEAD_PAD, device=w.device, dtype=torch.float32)
    out[:n, :, :d] = w.reshape(n, heads, d)
    return out.reshape(n_pad, heads * HEAD_PAD)


class UNet2DConditionGuidedModel(PackedModule):
    config_name = "config.json"

    def __init__(self, **config):
        super().__init__()
        cfg = dict(weights.UNET_CONFIG)
        cfg.update({k: v for k, v in config.items() if not k.startswith("_")})
        for k in ("block_out_channels", "attention_head_dim", "down_block_types", "up_block_types"):
            if list(cfg[k]) != list(weights.UNET_CONFIG[k]):
                # same failure mode as the reference's config validation (unet_2d_condition_guided.py:209-249)
                raise ValueError("only the tango_diffusion_light architecture is implemented (got %s=%r)" % (k, cfg[k]))
        self.config = FrozenConfig(cfg)
        register_tree(self, weights.unet_schema(cfg))

    # ------------------------------------------------------------------ reference ConfigMixin surface
    @classmethod
    def load_config(cls, path, **kwargs):
        """configuration_utils.py:256 — reads the JSON config (local path only; no hub I/O)."""
        import os
        if os.path.isdir(path):
            sub = kwargs.get("subfolder")
            path = os.path.join(path, sub, cls.config_name) if sub else os.path.join(path, cls.config_name)
        with open(path) as f:
            return json.load(f)

    @classmethod
    def from_config(cls, config, **kwargs):
        """configuration_utils.py:161 — `subfolder` and other hub kwargs are accepted and ignored."""
        return cls(**{k: v for k, v in dict(config).items() if not k.startswith("_")})

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    # ------------------------------------------------------------------ weight packing
    def _pack(self, sd, dev):
        pk = {}
        cfg = self.config
        boc = cfg["block_out_channels"]

        def conv(name):
            pk[name] = ops.pack_conv2d(sd[name + ".weight"], sd[name + ".bias"])

        def conv_s2(name):
            pk[name] = ops.pack_conv2d_im2col(sd[name + ".weight"], sd[name + ".bias"])

        def norm(name, pad=None):
            g, b = sd[name + ".weight"].float(), sd[name + ".bias"].float()
            if pad:
                g, b = _pad1(g, pad), _pad1(b, pad)
            pk[name] = (g.contiguous(), b.contiguous())

        temb_w, temb_b, temb_off = [], [], {}
        kv_w, kv_off = [], {}

        def resnet(p):
            norm(p + ".norm1")
            conv(p + ".conv1")
            norm(p + ".norm2")
            conv(p + ".conv2")
            if p + ".conv_shortcut.weight" in sd:
                conv(p + ".conv_shortcut")
            temb_off[p] = sum(w.shape[0] for w in temb_w)
            temb_w.append(sd[p + ".time_emb_proj.weight"].float())
            temb_b.append(sd[p + ".time_emb_proj.bias"].float())

        def transformer(p, c, heads):
            d = c // heads
            inner = heads * d
            dp = ops.round_up(inner, 64)
            hp = heads * HEAD_PAD
            norm(p + ".norm")
            pk[p + ".proj_in"] = ops.pack_linear(_pad2(sd[p + ".proj_in.weight"].float(), dp, c),
                                                 _pad1(sd[p + ".proj_in.bias"].float(), dp))
            t = p + ".transformer_blocks.0"
            for n in ("norm1", "norm2", "norm3"):
                norm(t + "." + n, pad=dp)
            qkv = torch.cat([_head_rows(sd[t + ".attn1.to_%s.weight" % x].float(), heads, d, dp) for x in "qkv"], 0)
            pk[t + ".attn1.qkv"] = ops.pack_linear(qkv)
            pk[t + ".attn1.to_out"] = ops.pack_linear(_head_cols(sd[t + ".attn1.to_out.0.weight"].float(), heads, d, dp),
                                                      _pad1(sd[t + ".attn1.to_out.0.bias"].float(), dp))
            pk[t + ".attn2.to_q"] = ops.pack_linear(_head_rows(sd[t + ".attn2.to_q.weight"].float(), heads, d, dp))
            pk[t + ".attn2.to_out"] = ops.pack_linear(_head_cols(sd[t + ".attn2.to_out.0.weight"].float(), heads, d, dp),
                                                      _pad1(sd[t + ".attn2.to_out.0.bias"].float(), dp))
            kv_off[t] = sum(w.shape[0] for w in kv_w)
            cross = cfg["cross_attention_dim"]
            kv_w.append(_head_rows(sd[t + ".attn2.to_k.weight"].float(), heads, d, cross))
            kv_w.append(_head_rows(sd[t + ".attn2.to_v.weight"].float(), heads, d, cross))
            # GEGLU: interleave (value_i, gate_i) rows so the epilogue sees the pair in adjacent columns
            w1 = sd[t + ".ff.net.0.proj.weight"].float()
            b1 = sd[t + ".ff.net.0.proj.bias"].float()
            f = 4 * inner
            fp = 4 * dp
            wi = torch.zeros(fp, 2, dp, device=dev)
            bi = torch.zeros(fp, 2, device=dev)
            wi[:f, 0, :inner] = w1[:f]
            wi[:f, 1, :inner] = w1[f:]
            bi[:f, 0] = b1[:f]
            bi[:f, 1] = b1[f:]
            pk[t + ".ff1"] = ops.pack_linear(wi.reshape(2 * fp, dp), bi.reshape(2 * fp))
            pk[t + ".ff2"] = ops.pack_linear(_pad2(sd[t + ".ff.net.2.weight"].float(), dp, fp),
                                             _pad1(sd[t + ".ff.net.2.bias"].float(), dp))
            pk[p + ".proj_out"] = ops.pack_linear(_pad2(sd[p + ".proj_out.weight"].float(), c, dp),
                                                  sd[p + ".proj_out.bias"].float())
            pk[p + ".meta"] = (heads, d, inner, dp)

        sd = {k: v.to(dev) for k, v in sd.items()}
        conv("conv_in")
        for n in ("time_embedding", "guidance_embedding"):
            for l in ("linear_1", "linear_2"):
                pk[n + "." + l] = (sd[n + "." + l + ".weight"].float().contiguous(),
                                   sd[n + "." + l + ".bias"].float().contiguous())
        pk["guidance_proj"] = sd["guidance_proj.weight"].float().contiguous()
        heads = cfg["attention_head_dim"]
        for i in range(4):
            for j in range(2):
                resnet("down_blocks.%d.resnets.%d" % (i, j))
                if i < 3:
                    transformer("down_blocks.%d.attentions.%d" % (i, j), boc[i], heads[i])
            if i < 3:
                conv_s2("down_blocks.%d.downsamplers.0.conv" % i)
        resnet("mid_block.resnets.0")
        transformer("mid_block.attentions.0", boc[3], heads[3])
        resnet("mid_block.resnets.1")
        rboc, rheads = boc[::-1], heads[::-1]
        for i in range(4):
            for j in range(3):
                resnet("up_blocks.%d.resnets.%d" % (i, j))
                if i > 0:
                    transformer("up_blocks.%d.attentions.%d" % (i, j), rboc[i], rheads[i])
            if i < 3:
                conv("up_blocks.%d.upsamplers.0.conv" % i)
                n_ = "up_blocks.%d.upsamplers.0.conv" % i
                pk[n_ + ".up2x"] = ops.pack_upsample2x_conv2d(sd[n_ + ".weight"], sd[n_ + ".bias"])
        norm("conv_norm_out")
        conv("conv_out")
        pk["temb_all"] = ops.pack_linear(torch.cat(temb_w, 0), torch.cat(temb_b, 0))
        pk["temb_off"] = temb_off
        pk["kv_all"] = ops.pack_linear(torch.cat(kv_w, 0))
        pk["kv_off"] = kv_off
        return pk

    # ------------------------------------------------------------------ building blocks (channels-last)
    def _resnet(self, pk, p, x, skip, temb_all, eps, x_stats=None, emit_stats=False):
        """ResnetBlock2D.forward, resnet.py:549-597. x fp32 [B,H,W,C1], skip fp32 [B,H,W,C2] or None (torch.cat).
        x_stats: GroupNorm moments of x already produced by the kernel that wrote x (fused statistics pass).
        Returns (out, moments of out or None)."""
        b, h, w, _ = x.shape
        g = self.config["norm_num_groups"]
        has_sc = (p + ".conv_shortcut") in pk
        c_in = x.shape[-1] + (skip.shape[-1] if skip is not None else 0)
        st = x_stats if (x_stats is not None and skip is None) else ops.groupnorm_stats(x, g, x2=skip)
        raw = torch.empty(b, h, w, c_in, device=x.device, dtype=ops.OPERAND_DTYPE) if has_sc else None
        a = ops.groupnorm_apply(x, g, st, *pk[p + ".norm1"], eps=eps, act=ACT_SILU, x2=skip, raw_out=raw)
        c_out = pk[p + ".conv1"].n
        off = pk["temb_off"][p]
        # conv1 + time embedding -> 16-bit hidden tensor (only GroupNorm 2 reads it) + its moments from the epilogue
        hid = torch.empty(b, h, w, c_out, device=x.device, dtype=ops.OPERAND_DTYPE)
        st2 = torch.empty(b, g, 2, device=x.device, dtype=torch.float32)
        ops.conv2d(a, pk[p + ".conv1"], out=hid, rowadd=temb_all[:, off:off + c_out], rowadd_rows=h * w, stats=st2,
                   stats_groups=g)
        a2 = ops.groupnorm_apply(hid, g, st2, *pk[p + ".norm2"], eps=eps, act=ACT_SILU)
        if has_sc:
            res = torch.empty(b, h, w, c_out, device=x.device, dtype=torch.float32)
            ops.conv2d(raw, pk[p + ".conv_shortcut"], out=res)
        else:
            res = x
        out = torch.empty(b, h, w, c_out, device=x.device, dtype=torch.float32)
        out_stats = torch.empty(b, g, 2, device=x.device, dtype=torch.float32) if emit_stats else None
        ops.conv2d(a2, pk[p + ".conv2"], out=out, residual=res, stats=out_stats, stats_groups=g)
        return out, out_stats

    def _transformer(self, pk, p, x, enc_kv, kv_len, n_text, x_stats=None, emit_stats=False):
        """Transformer2DModel + BasicTransformerBlock, transformer_2d.py:255-299, attention.py:276-334.
        Returns (out, GroupNorm moments of out or None)."""
        b, h, w, c = x.shape
        heads, d, inner, dp = pk[p + ".meta"]
        hp = heads * HEAD_PAD
        m = b * h * w
        dev = x.device
        f16 = ops.OPERAND_DTYPE
        t = p + ".transformer_blocks.0"
        g = self.config["norm_num_groups"]
        scale = d ** -0.5
        st = x_stats if x_stats is not None else ops.groupnorm_stats(x, g)
        a = ops.groupnorm_apply(x, g, st, *pk[p + ".norm"], eps=1e-6, act=ACT_NONE)
        y = torch.empty(m, dp, device=dev, dtype=torch.float32)
        ops.linear(a.view(m, c), pk[p + ".proj_in"], out=y)
        # --- self attention
        n1 = ops.layernorm(y, inner, *pk[t + ".norm1"], 1e-5)
        qkv = torch.empty(m, 3 * hp, device=dev, dtype=f16)
        ops.linear(n1, pk[t + ".attn1.qkv"], out=qkv)
        qkv5 = qkv.view(b, h * w, 3, heads, HEAD_PAD)
        o = torch.empty(b, h * w, heads, HEAD_PAD, device=dev, dtype=f16)
        ops.attention(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], scale, out=o)
        ops.linear(o.view(m, hp), pk[t + ".attn1.to_out"], out=y, residual=y)
        # --- cross attention (K/V of the prompt were projected once for all sites)
        n2 = ops.layernorm(y, inner, *pk[t + ".norm2"], 1e-5, out=n1)
        q = torch.empty(m, hp, device=dev, dtype=f16)
        ops.linear(n2, pk[t + ".attn2.to_q"], out=q)
        off = pk["kv_off"][t]
        kview = enc_kv[:, :, off:off + hp].unflatten(2, (heads, HEAD_PAD))
        vview = enc_kv[:, :, off + hp:off + 2 * hp].unflatten(2, (heads, HEAD_PAD))
        ops.attention(q.view(b, h * w, heads, HEAD_PAD), kview, vview, scale, kv_len=kv_len, out=o)
        ops.linear(o.view(m, hp), pk[t + ".attn2.to_out"], out=y, residual=y)
        # --- GEGLU feed-forward
        n3 = ops.layernorm(y, inner, *pk[t + ".norm3"], 1e-5, out=n1)
        ff = torch.empty(m, 4 * dp, device=dev, dtype=f16)
        ops.linear(n3, pk[t + ".ff1"], out=ff, act=ACT_GEGLU)
        y16 = n1  # reuse: [m, dp] 16-bit
        ops.linear(ff, pk[t + ".ff2"], out=y16, residual=y)
        out = torch.empty(b, h, w, c, device=dev, dtype=torch.float32)
        out_stats = torch.empty(b, g, 2, device=dev, dtype=torch.float32) if emit_stats else None
        ops.linear(y16, pk[p + ".proj_out"], out=out.view(m, c), residual=x.view(m, c), stats=out_stats, stats_groups=g,
                   stats_rows_per_img=h * w)
        return out, out_stats

    # ------------------------------------------------------------------ forward
    def forward(self, sample, timestep, guidance, encoder_hidden_states, class_labels=None, timestep_cond=None,
                guidance_cond=None, attention_mask=None, cross_attention_kwargs=None,
                down_block_additional_residuals=None, mid_block_additional_residual=None,
                encoder_attention_mask=None, return_dict=True):
        for name, val in (("class_labels", class_labels), ("timestep_cond", timestep_cond),
                          ("guidance_cond", guidance_cond), ("attention_mask", attention_mask),
                          ("cross_attention_kwargs", cross_attention_kwargs),
                          ("down_block_additional_residuals", down_block_additional_residuals),
                          ("mid_block_additional_residual", mid_block_additional_residual)):
            if val is not None:
                raise NotImplementedError("%s is not used on the ConsistencyTTA inference path" % name)
        out = self.forward_nhwc(sample, timestep, guidance, encoder_hidden_states, encoder_attention_mask)
        b, h, w, c = out.shape
        res = ops.nhwc_to_nchw(out)
        if torch.is_autocast_enabled():
            # the reference returns the autocast dtype when called under torch.autocast
            res = res.to(torch.get_autocast_gpu_dtype())
        if not return_dict:
            return (res,)
        return UNet2DConditionOutput(sample=res)

    def _batch_scalar(self, v, b, dev):
        """_prepare_tensor + expand, unet_2d_condition_guided.py:699-714,803,810."""
        if not torch.is_tensor(v):
            return torch.full((b,), float(v), device=dev, dtype=torch.float32)
        v = v.to(device=dev, dtype=torch.float32).reshape(-1)
        return v.expand(b).contiguous()

    def kv_lengths(self, encoder_attention_mask, b, n_text, dev):
        """encoder_attention_mask (bool [B, L]) -> per-row key count.  The kernels SKIP keys >= kv_len instead of adding
        the reference's -10000 bias (attention_processor.py:374-397), which is the same thing for the trailing-padding
        masks a tokenizer produces; a mask with an interior False cannot be expressed as a count and is refused."""
        if encoder_attention_mask is None:
            return None
        return prefix_mask_lengths(encoder_attention_mask).to(dev).contiguous()

    def project_text(self, enc, out=None):
        """K and V of every cross-attention site for prompt embeddings `enc` [B, L, 1024] (attention_processor.py:
        1117-1118): ONE GEMM against the 32 stacked to_k / to_v matrices -> 16-bit [B, L, 32 * heads * 64].  The result
        depends on the prompt only, so callers cache it across num_samples / multi-step queries (`enc_kv=`)."""
        pk = self.packed()
        dev = self.device
        b, n_text, _ = enc.shape
        enc16 = ops.groupnorm_apply(enc.to(dev).float().contiguous().view(b, n_text, 1, -1), 1, None, None, None,
                                    act=ACT_NONE)
        if out is None:
            out = torch.empty(b, n_text, pk["kv_all"].n, device=dev, dtype=ops.OPERAND_DTYPE)
        ops.linear(enc16.view(b * n_text, -1), pk["kv_all"], out=out.view(b * n_text, -1))
        return out

    def forward_nhwc(self, sample, timestep, guidance, enc, enc_mask=None, kv_len=None, sample_is_nhwc=False,
                     enc_kv=None):
        """Channels-last core: returns fp32 [B, 256, 16, 8].  `kv_len` int32 [B] may be given instead of a mask;
        `enc_kv` (from `project_text`) replaces `enc` when the prompt's K/V were already projected."""
        pk = self.packed()
        dev = self.device
        f16 = ops.OPERAND_DTYPE
        cfg = self.config
        eps = cfg["norm_eps"]
        if sample_is_nhwc:
            x0 = sample
            b = x0.shape[0]
        else:
            b = sample.shape[0]
            x0 = ops.nchw_to_nhwc(sample.to(dev).float(), dtype=f16)
        n_text = enc_kv.shape[1] if enc_kv is not None else enc.shape[1]
        if kv_len is None:
            # trailing padding only (T5 tokenizer); masked keys == truncated keys
            kv_len = self.kv_lengths(enc_mask, b, n_text, dev)

        # 1. time + guidance embeddings (unet_2d_condition_guided.py:803-816)
        t = self._batch_scalar(timestep, b, dev)
        gw = self._batch_scalar(guidance, b, dev)
        tf, gf = ops.time_features(t, gw, pk["guidance_proj"])
        h1 = ops.small_linear(tf, *pk["time_embedding.linear_1"], act_out=ACT_SILU)
        emb = ops.small_linear(h1, *pk["time_embedding.linear_2"])
        h2 = ops.small_linear(gf, *pk["guidance_embedding.linear_1"], act_out=ACT_SILU)
        ops.small_linear(h2, *pk["guidance_embedding.linear_2"], accumulate=True, out=emb)
        # all 22 time_emb_proj(SiLU(emb)) at once (resnet.py:573)
        emb16 = ops.groupnorm_apply(emb.view(b, 1, 1, -1), 1, None, None, None, act=ACT_SILU)
        temb_all = torch.empty(b, pk["temb_all"].n, device=dev, dtype=torch.float32)
        ops.linear(emb16.view(b, -1), pk["temb_all"], out=temb_all)

        # 2. text K/V for every cross-attention site (attention_processor.py:1117-1118), one GEMM
        if enc_kv is None:
            enc_kv = self.project_text(enc)

        # 3. down path.  `st` carries the GroupNorm moments of x whenever the kernel that produced x could emit them
        _, hh, ww, _ = x0.shape
        g = cfg["norm_num_groups"]
        x = torch.empty(b, hh, ww, cfg["block_out_channels"][0], device=dev, dtype=torch.float32)
        st = torch.empty(b, g, 2, device=dev, dtype=torch.float32)
        ops.conv2d(x0, pk["conv_in"], out=x, stats=st, stats_groups=g)
        skips = [x]
        for i in range(4):
            for j in range(2):
                x, st = self._resnet(pk, "down_blocks.%d.resnets.%d" % (i, j), x, None, temb_all, eps, st, True)
                if i < 3:
                    x, st = self._transformer(pk, "down_blocks.%d.attentions.%d" % (i, j), x, enc_kv, kv_len, n_text,
                                              st, True)
                skips.append(x)
            if i < 3:
                p = "down_blocks.%d.downsamplers.0.conv" % i
                bb, h_, w_, c_ = x.shape
                x16 = ops.groupnorm_apply(x, 1, None, None, None, act=ACT_NONE)
                cols = ops.im2col_s2(x16)
                x = torch.empty(bb, h_ // 2, w_ // 2, c_, device=dev, dtype=torch.float32)
                st = torch.empty(b, g, 2, device=dev, dtype=torch.float32)
                ops.linear(cols, pk[p], out=x.view(-1, c_), stats=st, stats_groups=g,
                           stats_rows_per_img=(h_ // 2) * (w_ // 2))
                skips.append(x)
        # 4. mid
        x, st = self._resnet(pk, "mid_block.resnets.0", x, None, temb_all, eps, st, True)
        x, st = self._transformer(pk, "mid_block.attentions.0", x, enc_kv, kv_len, n_text, st, True)
        x, st = self._resnet(pk, "mid_block.resnets.1", x, None, temb_all, eps, st, False)
        # 5. up path (the resnets normalise torch.cat([x, skip]): their moments come from the stand-alone kernel)
        for i in range(4):
            for j in range(3):
                last = i == 3 and j == 2
                x, st = self._resnet(pk, "up_blocks.%d.resnets.%d" % (i, j), x, skips.pop(), temb_all, eps, None, i > 0)
                if i > 0:
                    x, st = self._transformer(pk, "up_blocks.%d.attentions.%d" % (i, j), x, enc_kv, kv_len, n_text, st,
                                              last)
            if i < 3:
                p = "up_blocks.%d.upsamplers.0.conv" % i
                bb, h_, w_, c_ = x.shape
                if ops.upsample2x_conv_supported(h_, w_):
                    # Upsample2D (resnet.py:126-161): nearest 2x + conv3x3 as four 2x2 phase convs on the low-res tensor
                    x16 = ops.groupnorm_apply(x, 1, None, None, None, act=ACT_NONE)
                    x = torch.empty(bb, 2 * h_, 2 * w_, c_, device=dev, dtype=torch.float32)
                    ops.conv2d_upsample2x(x16, pk[p + ".up2x"], x)
                else:
                    up = ops.groupnorm_apply(x, 1, None, None, None, act=ACT_NONE, upsample=True)
                    x = torch.empty(bb, 2 * h_, 2 * w_, c_, device=dev, dtype=torch.float32)
                    ops.conv2d(up, pk[p], out=x)
        # 6. out
        if st is None:
            st = ops.groupnorm_stats(x, g)
        a = ops.groupnorm_apply(x, g, st, *pk["conv_norm_out"], eps=eps, act=ACT_SILU)
        out = torch.empty(b, x.shape[1], x.shape[2], cfg["out_channels"], device=dev, dtype=torch.float32)
        ops.conv2d(a, pk["conv_out"], out=out)
        return out
