"""UNet2DConditionGuidedModel.forward restated functionally (fp32, NCHW) over the reference state_dict.

Follows diffusers/models/unet_2d_condition_guided.py:716-945 and the blocks it instantiates for
configs/tango_diffusion_light.json.
"""
import math

import torch
import torch.nn.functional as F

BLOCK_OUT = [256, 512, 1024, 1024]
HEADS = [5, 10, 20, 20]
GROUPS = 32


def timestep_embedding(t, dim=256):
    """diffusers/models/embeddings.py:25-65 with flip_sin_to_cos=True, downscale_freq_shift=0."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def fourier_embedding(w, weight):
    """GaussianFourierProjection, embeddings.py:240-248 (log=False, flip_sin_to_cos=True).
    A python-float guidance becomes a float64 tensor (unet_2d_condition_guided.py:703-708)."""
    x_proj = w[:, None] * weight[None, :].to(w.dtype) * 2 * math.pi
    return torch.cat([torch.cos(x_proj), torch.sin(x_proj)], dim=-1)


def mlp_embedding(sd, p, x):
    """TimestepEmbedding.forward, embeddings.py:189-202."""
    x = F.linear(x, sd[p + ".linear_1.weight"], sd[p + ".linear_1.bias"])
    x = F.silu(x)
    return F.linear(x, sd[p + ".linear_2.weight"], sd[p + ".linear_2.bias"])


def resnet(sd, p, x, emb, eps=1e-5):
    """ResnetBlock2D.forward, diffusers/models/resnet.py:549-597 (time_embedding_norm default, scale 1)."""
    h = F.group_norm(x, GROUPS, sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    t = F.linear(F.silu(emb), sd[p + ".time_emb_proj.weight"], sd[p + ".time_emb_proj.bias"])
    h = h + t[:, :, None, None]
    h = F.group_norm(h, GROUPS, sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], eps)
    h = F.silu(h)
    h = F.conv2d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def attention(sd, p, x, ctx, bias, heads):
    """Attention + AttnProcessor2_0, attention_processor.py:1077-1147 (to_q/k/v no bias, to_out.0 with bias)."""
    b, lq, inner = x.shape
    ctx = x if ctx is None else ctx
    q = F.linear(x, sd[p + ".to_q.weight"])
    k = F.linear(ctx, sd[p + ".to_k.weight"])
    v = F.linear(ctx, sd[p + ".to_v.weight"])
    d = inner // heads
    q = q.view(b, -1, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    if bias is not None:
        # prepare_attention_mask repeats over heads (attention_processor.py:380-415) -> [B, heads, 1, L]
        bias = bias[:, None, :, :].expand(b, heads, 1, bias.shape[-1])
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias)
    o = o.transpose(1, 2).reshape(b, -1, inner)
    return F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])


def transformer(sd, p, x, enc, enc_bias, heads):
    """Transformer2DModel.forward (use_linear_projection), transformer_2d.py:255-299 +
    BasicTransformerBlock.forward, attention.py:276-334 + GEGLU FeedForward, attention.py:383-386,430-432."""
    b, c, h, w = x.shape
    res = x
    y = F.group_norm(x, GROUPS, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6)
    y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
    y = F.linear(y, sd[p + ".proj_in.weight"], sd[p + ".proj_in.bias"])
    t = p + ".transformer_blocks.0"
    inner = y.shape[-1]

    def ln(name, z):
        return F.layer_norm(z, (inner,), sd[t + name + ".weight"], sd[t + name + ".bias"], 1e-5)

    y = attention(sd, t + ".attn1", ln(".norm1", y), None, None, heads) + y
    y = attention(sd, t + ".attn2", ln(".norm2", y), enc, enc_bias, heads) + y
    z = F.linear(ln(".norm3", y), sd[t + ".ff.net.0.proj.weight"], sd[t + ".ff.net.0.proj.bias"])
    hidden, gate = z.chunk(2, dim=-1)
    z = hidden * F.gelu(gate)
    y = F.linear(z, sd[t + ".ff.net.2.weight"], sd[t + ".ff.net.2.bias"]) + y
    y = F.linear(y, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    y = y.reshape(b, h, w, c).permute(0, 3, 1, 2)
    return y + res


def unet_forward(sd, sample, timestep, guidance, enc, enc_mask=None):
    """unet_2d_condition_guided.py:716-945.  sample [B,8,256,16]; timestep scalar / [B]; guidance float / [B];
    enc [B,L,1024]; enc_mask bool [B,L] (True = keep).  Returns [B,8,256,16]."""
    b = sample.shape[0]
    enc_bias = None
    if enc_mask is not None:  # :793-795
        enc_bias = ((1 - enc_mask.to(sample.dtype)) * -10000.0).unsqueeze(1)
    t = torch.as_tensor(timestep).reshape(-1).expand(b).to(sample.device)
    if torch.is_tensor(guidance):
        g = guidance.reshape(-1).expand(b).to(sample.device)
    else:
        g = torch.tensor([float(guidance)], dtype=torch.float64).expand(b).to(sample.device)
    t_emb = mlp_embedding(sd, "time_embedding", timestep_embedding(t).to(sample.dtype))  # :803-808
    g_emb = mlp_embedding(sd, "guidance_embedding", fourier_embedding(g, sd["guidance_proj.weight"]).to(sample.dtype))
    emb = t_emb + g_emb  # :816

    x = F.conv2d(sample, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)  # :863
    skips = [x]
    for i in range(4):  # :866-880, unet_2d_blocks.py:912-972,1027-1058
        for j in range(2):
            x = resnet(sd, "down_blocks.%d.resnets.%d" % (i, j), x, emb)
            if i < 3:
                x = transformer(sd, "down_blocks.%d.attentions.%d" % (i, j), x, enc, enc_bias, HEADS[i])
            skips.append(x)
        if i < 3:  # Downsample2D, resnet.py:199-208
            p = "down_blocks.%d.downsamplers.0.conv" % i
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=2, padding=1)
            skips.append(x)
    x = resnet(sd, "mid_block.resnets.0", x, emb)  # unet_2d_blocks.py:588-609
    x = transformer(sd, "mid_block.attentions.0", x, enc, enc_bias, HEADS[3])
    x = resnet(sd, "mid_block.resnets.1", x, emb)
    rheads = HEADS[::-1]
    for i in range(4):  # :908-934, unet_2d_blocks.py:2017-2078,2129-2159
        for j in range(3):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet(sd, "up_blocks.%d.resnets.%d" % (i, j), x, emb)
            if i > 0:
                x = transformer(sd, "up_blocks.%d.attentions.%d" % (i, j), x, enc, enc_bias, rheads[i])
        if i < 3:  # Upsample2D, resnet.py:126-161
            p = "up_blocks.%d.upsamplers.0.conv" % i
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)
    x = F.group_norm(x, GROUPS, sd["conv_norm_out.weight"], sd["conv_norm_out.bias"], 1e-5)  # :937-940
    x = F.silu(x)
    return F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
