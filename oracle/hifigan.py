"""HiFi-GAN Generator + vocoder_infer restated functionally (fp32).

Follows audioldm/hifigan/models.py:56-63,101-117 and audioldm/hifigan/utilities.py:76-91 for HIFIGAN_16K_64.
"""
import numpy as np
import torch
import torch.nn.functional as F

UPS = [(16, 5), (16, 4), (8, 2), (4, 2), (4, 2)]  # (kernel, stride), padding (k - s) // 2, models.py:84-89
KERNELS = [3, 7, 11]
DILATIONS = [1, 3, 5]


def resblock(sd, p, x, k):
    """ResBlock.forward, models.py:56-63."""
    for m, d in enumerate(DILATIONS):
        xt = F.leaky_relu(x, 0.1)
        xt = F.conv1d(xt, sd[p + ".convs1.%d.weight" % m], sd[p + ".convs1.%d.bias" % m], dilation=d,
                      padding=(k * d - d) // 2)
        xt = F.leaky_relu(xt, 0.1)
        xt = F.conv1d(xt, sd[p + ".convs2.%d.weight" % m], sd[p + ".convs2.%d.bias" % m], padding=(k - 1) // 2)
        x = xt + x
    return x


def generator_forward(sd, mel, p="vocoder."):
    """Generator.forward, models.py:101-117. mel [B, 64, T] -> [B, 1, T * 160 + 32]."""
    x = F.conv1d(mel, sd[p + "conv_pre.weight"], sd[p + "conv_pre.bias"], padding=3)
    for i, (k, s) in enumerate(UPS):
        x = F.leaky_relu(x, 0.1)
        x = F.conv_transpose1d(x, sd[p + "ups.%d.weight" % i], sd[p + "ups.%d.bias" % i], stride=s,
                               padding=(k - s) // 2)
        xs = None
        for j, ks in enumerate(KERNELS):
            r = resblock(sd, p + "resblocks.%d" % (i * 3 + j), x, ks)
            xs = r if xs is None else xs + r
        x = xs / 3
    x = F.leaky_relu(x)  # default slope 0.01, models.py:113
    x = F.conv1d(x, sd[p + "conv_post.weight"], sd[p + "conv_post.bias"], padding=3)
    return torch.tanh(x)


def decode_to_waveform(sd, dec, return_float=False):
    """AutoencoderKL.decode_to_waveform (autoencoder.py:108-111) + vocoder_infer (utilities.py:76-91)."""
    mel = dec.squeeze(1).permute(0, 2, 1)
    wavs = generator_forward(sd, mel).squeeze(1).float()
    if return_float:
        return wavs
    wavs = wavs - (wavs.max() + wavs.min()) / 2
    return (wavs.cpu().numpy() * 32768).astype("int16")


def to_int16(wavs):
    wavs = wavs - (wavs.max() + wavs.min()) / 2
    return (wavs.cpu().numpy() * 32768).astype("int16")
