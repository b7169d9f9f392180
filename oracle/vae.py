"""AutoencoderKL decode side restated functionally (fp32, NCHW).

Follows audioldm/variational_autoencoder/autoencoder.py:91-106 and modules.py:38-57,155-175,204-230,650-683
for the audioldm-s-full first-stage config (ch 128, ch_mult [1,2,4], 2 res blocks, z 8, no attn_resolutions).
"""
import torch
import torch.nn.functional as F


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)  # Normalize, modules.py:38-41


def _swish(x):
    return x * torch.sigmoid(x)  # modules.py:33-35


def _conv(sd, p, x, pad):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def resnet_block(sd, p, x):
    """ResnetBlock.forward with temb=None, modules.py:155-175."""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), 1)
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(sd, p + ".nin_shortcut", x, 0)
    return x + h


def attn_block(sd, p, x):
    """AttnBlock.forward, modules.py:204-230 (single head over h*w tokens, scale c^-0.5)."""
    h_ = _gn(sd, p + ".norm", x)
    q = _conv(sd, p + ".q", h_, 0)
    k = _conv(sd, p + ".k", h_, 0)
    v = _conv(sd, p + ".v", h_, 0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, p + ".proj_out", h_, 0)


def decoder_forward(sd, z, p="decoder"):
    """Decoder.forward, modules.py:650-683."""
    h = _conv(sd, p + ".conv_in", z, 1)
    h = resnet_block(sd, p + ".mid.block_1", h)
    h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for lvl in (2, 1, 0):
        for b in range(3):
            h = resnet_block(sd, p + ".up.%d.block.%d" % (lvl, b), h)
        if lvl != 0:  # Upsample, modules.py:53-57
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, p + ".up.%d.upsample.conv" % lvl, h, 1)
    h = _swish(_gn(sd, p + ".norm_out", h))
    return _conv(sd, p + ".conv_out", h, 1)


def decode_first_stage(sd, z, scale_factor):
    """AutoencoderKL.decode_first_stage / decode, autoencoder.py:91-106 (subband 1, no EMA modules)."""
    z = z / scale_factor
    z = _conv(sd, "post_quant_conv", z, 0)
    return decoder_forward(sd, z)
