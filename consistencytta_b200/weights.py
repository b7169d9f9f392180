"""State-dict schemas of the hot-path networks and a deterministic synthetic-weight generator.

No checkpoints exist in this environment (and none can be downloaded), so parity tests, smoke() and bench.py use
random-init weights.  The generator mimics the reference modules' default initialisation (PyTorch kaiming-uniform
U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv/linear weights and biases; HiFi-GAN's N(0, 0.01) `init_weights`,
audioldm/hifigan/models.py:10-13,96-97) and seeds every tensor from its key, so the same state_dict is produced on
any machine without shipping 650 M parameters.  Key names / shapes are exactly the reference's
(fixture tests/golden/state_dict_schema.json was dumped from the reference modules).
"""
import hashlib
import math

import torch

# configs/tango_diffusion_light.json
UNET_CONFIG = {
    "in_channels": 8,
    "out_channels": 8,
    "block_out_channels": [256, 512, 1024, 1024],
    "attention_head_dim": [5, 10, 20, 20],  # used as the number of heads (unet_2d_condition_guided.py:206)
    "down_block_types": ["CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"],
    "up_block_types": ["UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D"],
    "layers_per_block": 2,
    "cross_attention_dim": 1024,
    "norm_num_groups": 32,
    "norm_eps": 1e-5,
    "flip_sin_to_cos": True,
    "freq_shift": 0,
    "use_linear_projection": True,
    "upcast_attention": True,
    "act_fn": "silu",
    "sample_size": [32, 2],
    "center_input_sample": False,
    "downsample_padding": 1,
    "dual_cross_attention": False,
    "mid_block_scale_factor": 1,
    "num_class_embeds": None,
    "only_cross_attention": False,
}

# audioldm/utils.py:160-182 (first_stage_config of audioldm-s-full)
VAE_CONFIG = {
    "embed_dim": 8,
    "ddconfig": {"double_z": True, "z_channels": 8, "resolution": 256, "downsample_time": False, "in_channels": 1,
                 "out_ch": 1, "ch": 128, "ch_mult": [1, 2, 4], "num_res_blocks": 2, "attn_resolutions": [],
                 "dropout": 0.0},
}

# audioldm/hifigan/utilities.py:9-39
HIFIGAN_CONFIG = {
    "upsample_rates": [5, 4, 2, 2, 2],
    "upsample_kernel_sizes": [16, 16, 8, 4, 4],
    "upsample_initial_channel": 1024,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "num_mels": 64,
}


# ------------------------------------------------------------------------------------------------ schemas
def _conv(d, name, cout, cin, *k):
    d[name + ".weight"] = ((cout, cin) + tuple(k), "w")
    d[name + ".bias"] = ((cout,), "b:%d" % (cin * int(math.prod(k))))


def _lin(d, name, n, k, bias=True):
    d[name + ".weight"] = ((n, k), "w")
    if bias:
        d[name + ".bias"] = ((n,), "b:%d" % k)


def _norm(d, name, c):
    d[name + ".weight"] = ((c,), "g")
    d[name + ".bias"] = ((c,), "z")


def _unet_resnet(d, p, cin, cout, temb=1024):
    _norm(d, p + ".norm1", cin)
    _conv(d, p + ".conv1", cout, cin, 3, 3)
    _lin(d, p + ".time_emb_proj", cout, temb)
    _norm(d, p + ".norm2", cout)
    _conv(d, p + ".conv2", cout, cout, 3, 3)
    if cin != cout:
        _conv(d, p + ".conv_shortcut", cout, cin, 1, 1)


def _unet_transformer(d, p, c, heads, cross):
    inner = heads * (c // heads)  # unet_2d_blocks.py:873-876 -> 255 / 510 / 1020
    _norm(d, p + ".norm", c)
    _lin(d, p + ".proj_in", inner, c)
    t = p + ".transformer_blocks.0"
    for a, kv in (("attn1", inner), ("attn2", cross)):
        _lin(d, t + "." + a + ".to_q", inner, inner, bias=False)
        _lin(d, t + "." + a + ".to_k", inner, kv, bias=False)
        _lin(d, t + "." + a + ".to_v", inner, kv, bias=False)
        _lin(d, t + "." + a + ".to_out.0", inner, inner)
    for n in ("norm1", "norm2", "norm3"):
        _norm(d, t + "." + n, inner)
    _lin(d, t + ".ff.net.0.proj", 8 * inner, inner)
    _lin(d, t + ".ff.net.2", inner, 4 * inner)
    _lin(d, p + ".proj_out", c, inner)


def unet_schema(cfg=UNET_CONFIG):
    """{key: (shape, kind)} of UNet2DConditionGuidedModel.state_dict() (691 tensors, 559 209 676 parameters)."""
    d = {}
    boc = cfg["block_out_channels"]
    heads = cfg["attention_head_dim"]
    cross = cfg["cross_attention_dim"]
    lpb = cfg["layers_per_block"]
    temb = boc[0] * 4
    _conv(d, "conv_in", boc[0], cfg["in_channels"], 3, 3)
    d["guidance_proj.weight"] = ((temb // 2,), "n")
    _lin(d, "time_embedding.linear_1", temb, boc[0])
    _lin(d, "time_embedding.linear_2", temb, temb)
    _lin(d, "guidance_embedding.linear_1", temb, temb)
    _lin(d, "guidance_embedding.linear_2", temb, temb)
    out_ch = boc[0]
    for i, bt in enumerate(cfg["down_block_types"]):
        in_ch, out_ch = out_ch, boc[i]
        for j in range(lpb):
            _unet_resnet(d, "down_blocks.%d.resnets.%d" % (i, j), in_ch if j == 0 else out_ch, out_ch, temb)
            if bt.startswith("CrossAttn"):
                _unet_transformer(d, "down_blocks.%d.attentions.%d" % (i, j), out_ch, heads[i], cross)
        if i != len(boc) - 1:
            _conv(d, "down_blocks.%d.downsamplers.0.conv" % i, out_ch, out_ch, 3, 3)
    _unet_resnet(d, "mid_block.resnets.0", boc[-1], boc[-1], temb)
    _unet_transformer(d, "mid_block.attentions.0", boc[-1], heads[-1], cross)
    _unet_resnet(d, "mid_block.resnets.1", boc[-1], boc[-1], temb)
    rev = list(reversed(boc))
    rheads = list(reversed(heads))
    out_ch = rev[0]
    for i, bt in enumerate(cfg["up_block_types"]):
        prev, out_ch = out_ch, rev[i]
        in_ch = rev[min(i + 1, len(boc) - 1)]
        for j in range(lpb + 1):
            skip = in_ch if j == lpb else out_ch
            rin = prev if j == 0 else out_ch
            _unet_resnet(d, "up_blocks.%d.resnets.%d" % (i, j), rin + skip, out_ch, temb)
            if bt.startswith("CrossAttn"):
                _unet_transformer(d, "up_blocks.%d.attentions.%d" % (i, j), out_ch, rheads[i], cross)
        if i != len(boc) - 1:
            _conv(d, "up_blocks.%d.upsamplers.0.conv" % i, out_ch, out_ch, 3, 3)
    _norm(d, "conv_norm_out", boc[0])
    _conv(d, "conv_out", cfg["out_channels"], boc[0], 3, 3)
    return d


def _vae_resnet(d, p, cin, cout):
    _norm(d, p + ".norm1", cin)
    _conv(d, p + ".conv1", cout, cin, 3, 3)
    _norm(d, p + ".norm2", cout)
    _conv(d, p + ".conv2", cout, cout, 3, 3)
    if cin != cout:
        _conv(d, p + ".nin_shortcut", cout, cin, 1, 1)


def vae_decoder_schema(cfg=VAE_CONFIG):
    """Decode-side keys of AutoencoderKL.state_dict(): post_quant_conv.* and decoder.* (encoder/quant_conv unused)."""
    d = {}
    dd = cfg["ddconfig"]
    ch, mult, nrb, z = dd["ch"], dd["ch_mult"], dd["num_res_blocks"], dd["z_channels"]
    _conv(d, "post_quant_conv", z, cfg["embed_dim"], 1, 1)
    block_in = ch * mult[-1]
    _conv(d, "decoder.conv_in", block_in, z, 3, 3)
    _vae_resnet(d, "decoder.mid.block_1", block_in, block_in)
    _norm(d, "decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(d, "decoder.mid.attn_1." + n, block_in, block_in, 1, 1)
    _vae_resnet(d, "decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(len(mult))):
        block_out = ch * mult[lvl]
        for b in range(nrb + 1):
            _vae_resnet(d, "decoder.up.%d.block.%d" % (lvl, b), block_in, block_out)
            block_in = block_out
        if lvl != 0:
            _conv(d, "decoder.up.%d.upsample.conv" % lvl, block_in, block_in, 3, 3)
    _norm(d, "decoder.norm_out", block_in)
    _conv(d, "decoder.conv_out", dd["out_ch"], block_in, 3, 3)
    return d


def vocoder_schema(cfg=HIFIGAN_CONFIG, prefix="vocoder."):
    """HiFi-GAN Generator keys (weight norm already removed, hifigan/utilities.py:71): 194 tensors."""
    d = {}
    c0 = cfg["upsample_initial_channel"]
    d[prefix + "conv_pre.weight"] = ((c0, cfg["num_mels"], 7), "w")
    d[prefix + "conv_pre.bias"] = ((c0,), "b:%d" % (cfg["num_mels"] * 7))
    nk = len(cfg["resblock_kernel_sizes"])
    ch = c0
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        cin, ch = c0 // 2 ** i, c0 // 2 ** (i + 1)
        d[prefix + "ups.%d.weight" % i] = ((cin, ch, k), "h")
        # ConvTranspose1d bias bound uses fan_in computed from weight.size(1) * k (torch's _calculate_fan_in_and_fan_out)
        d[prefix + "ups.%d.bias" % i] = ((ch,), "b:%d" % (ch * k))
        for j, ks in enumerate(cfg["resblock_kernel_sizes"]):
            for m in range(3):
                for cname in ("convs1", "convs2"):
                    p = prefix + "resblocks.%d.%s.%d" % (i * nk + j, cname, m)
                    d[p + ".weight"] = ((ch, ch, ks), "h")
                    d[p + ".bias"] = ((ch,), "b:%d" % (ch * ks))
    d[prefix + "conv_post.weight"] = ((1, ch, 7), "h")
    d[prefix + "conv_post.bias"] = ((1,), "b:%d" % (ch * 7))
    return d


def vae_schema():
    d = vae_decoder_schema()
    d.update(vocoder_schema())
    return d


# ------------------------------------------------------------------------------------------------ generator
def _seed_for(key, seed):
    h = hashlib.sha256(("%d:%s" % (seed, key)).encode()).digest()
    return int.from_bytes(h[:7], "little")


def _gen(key, shape, kind, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(_seed_for(key, seed))
    if kind == "w":  # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
        fan_in = int(math.prod(shape[1:]))
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    if kind.startswith("b:"):
        b = 1.0 / math.sqrt(int(kind[2:]))
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    if kind == "h":  # HiFi-GAN init_weights
        return torch.randn(shape, generator=g) * 0.01
    if kind == "n":
        return torch.randn(shape, generator=g)
    if kind == "g":  # norm scale: 1 + small perturbation so the affine path is exercised
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if kind == "z":
        return 0.1 * torch.randn(shape, generator=g)
    raise ValueError(kind)


def make_state_dict(schema, seed=0):
    return {k: _gen(k, shape, kind, seed) for k, (shape, kind) in schema.items()}


def make_unet_state_dict(seed=0):
    return make_state_dict(unet_schema(), seed)


def make_vae_state_dict(seed=1):
    return make_state_dict(vae_schema(), seed)


SCALE_FACTOR = 0.9227  # arbitrary fixed latent scale (the real value lives in the AudioLDM checkpoint)


def synthetic_inputs(batch, length, seed=1234):
    """Deterministic synthetic inputs shared by the golden script, the tests and bench.py:
    noise [B,8,256,16] ~ N(0,1), text embeddings [B,L,1024] ~ N(0,1), bool mask with ragged trailing padding."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    noise = torch.randn(batch, 8, 256, 16, generator=g)
    enc = torch.randn(batch, length, 1024, generator=g)
    mask = torch.ones(batch, length, dtype=torch.bool)
    for b in range(1, batch):
        mask[b, length - min(length - 1, (4 * b + 8) % length):] = False  # trailing padding, different per row
    return noise, enc, mask
