"""Generation entry points with the reference's call signatures, over one fused single-step engine.

  * ConsistencyTTA           — easy_inference/consistencytta.py:12-200 (`forward(prompt, cfg_scale_input, ...)`)
  * AudioLCM                 — models/audio_consistency_model.py:21,160-204,429-548 (`inference(...)`, `load_pretrained`)
  * SingleStepEngine         — the hot path itself: scheduler prologue -> UNet -> (post-CFG) -> VAE decode -> HiFi-GAN
                               -> centring/int16, all on channels-last device buffers, captured in a CUDA graph per
                               (batch, text length, cfg) bucket so a replay costs one launch on the host.

The FLAN-T5 text encoder stays stock PyTorch outside the hot path; without its weights the
engine is driven with prompt embeddings directly (`generate_from_embeddings`).
"""
from collections import OrderedDict
from time import time

import torch
from torch import nn

from . import ops, weights
from .scheduler import HeunDiscreteScheduler
from .unet import UNet2DConditionGuidedModel, prefix_mask_lengths
from .vae import AutoencoderKL

LATENT_SHAPE = (8, 256, 16)
WAVE_SAMPLES = 163872  # 1024 mel frames through ups 5*4*2*2*2 with odd (k - s) in the first stage


def slice_request_rows(t, lo, hi, b, cf):
    """Rows [lo, hi) of a per-request tensor of a b-clip request.  Scalars and broadcast (1-row) tensors pass through;
    with post-CFG (`cf`) text tensors hold 2b rows laid out [unconditional ; conditional] (consistencytta.py:171,
    audio_consistency_model.py:441) and the slice keeps that layout."""
    if t is None or not torch.is_tensor(t) or t.dim() == 0 or t.shape[0] == 1:
        return t
    if cf and t.shape[0] == 2 * b:
        return torch.cat([t[lo:hi], t[b + lo:b + hi]])
    return t[lo:hi]


def scheduler_input_scale(scheduler, timestep):
    """The scalar `scheduler.scale_model_input(x, t)` multiplies x by.  Our schedulers expose it as `input_scale`; for any
    other object with the diffusers scheduler surface (e.g. the reference's own HeunDiscreteScheduler / DDIMScheduler
    passed to `AudioLCM.inference`) it is read off a probe tensor."""
    if hasattr(scheduler, "input_scale"):
        return float(scheduler.input_scale(timestep))
    probe = torch.ones(1, 1, 1, 1)
    t = timestep if torch.is_tensor(timestep) else torch.tensor(timestep)
    return float(scheduler.scale_model_input(probe, t).reshape(-1)[0])


class SingleStepEngine:
    """The hot path: scheduler prologue -> UNet -> (post-CFG) -> VAE decode -> HiFi-GAN -> centring / int16.

    Static device buffers live in *buckets* keyed by (clips, text-length bucket, post-CFG, device); text lengths are
    rounded up to `text_bucket` tokens (padding keys are skipped through the per-row key count, so the result does not
    depend on the bucket).  Each bucket owns one captured CUDA graph per (guidance_post, use_ema, stages, reuse_text)
    variant; sigma / timestep / guidance are DATA (device scalars), so every step of the multi-step sampler replays the
    same graph.  At most `max_buckets` buckets are kept (least recently used first out) and all graphs capture into one
    shared memory pool, so a stream of different prompt lengths costs max-over-buckets memory, not the sum.

    Captured graphs hold raw pointers into the packed weights: the engine compares the modules' `pack_version` on every
    call and drops its graphs when a checkpoint was (re)loaded or the modules were moved / cast.

    `max_batch` bounds the clips resident in one pass (about 0.47 GB of activations per clip at the widest point, so 64
    clips = 30 GB): larger requests run as micro-batches through the same graphs and are stitched together, with the
    batch-GLOBAL waveform centring of `vocoder_infer` (hifigan/utilities.py:84-86) applied over the whole request, so
    the int16 result does not depend on how the request was split.

    Aliasing: the tensors `run()` returns are the bucket's static / graph-pool buffers and are overwritten by the next
    `run()` of the engine — pass `clone=True` (the public ConsistencyTTA / AudioLCM entry points do) to keep them."""

    def __init__(self, unet, vae, scheduler=None, use_graphs=True, max_batch=64, text_bucket=16, max_buckets=6,
                 validate_mask=True):
        self.unet = unet
        self.vae = vae
        self.scheduler = scheduler or HeunDiscreteScheduler.from_pretrained()
        self.use_graphs = use_graphs
        self.max_batch = max_batch
        self.text_bucket = max(int(text_bucket), 1)
        self.max_buckets = max(int(max_buckets), 1)
        self.validate_mask = validate_mask
        # 16-bit operand type per stage: fp16 (what keeps UNet / VAE within the 1e-2 gate) until a stage's output came back
        # non-finite, then bf16 for that stage from then on (`escalate`; run(check_overflow=True) does it automatically)
        self.stage_dtype = {"unet": torch.float16, "vae": torch.float16, "vocoder": torch.float16}
        self.escalations = []
        self._buckets = OrderedDict()
        self._pool = None
        self._versions = None
        self.text_projections = 0     # number of prompt K/V projections (kv_all GEMM) issued or replayed
        self.captures = 0
        self.last = None              # (bucket entry, graph-variant entry) of the last single-pass run()

    def reset(self):
        """Drops every captured graph and static buffer."""
        self._buckets.clear()
        self.last = None

    def escalate(self, stage):
        """Switches one stage ('unet' / 'vae' / 'vocoder') to bf16 operands; its buckets are re-created on the next run."""
        if self.stage_dtype[stage] != torch.bfloat16:
            self.stage_dtype[stage] = torch.bfloat16
            self.escalations.append(stage)
            self.reset()

    def _check_versions(self):
        v = (self.unet.pack_version, self.vae.pack_version if self.vae is not None else None)
        if v != self._versions:
            self.reset()
            self._versions = v

    # -------------------------------------------------------------------------------------------- hot path
    def _compute(self, io, guidance_post, use_ema, stages, reuse_text, refs):
        """Launches every kernel of the path on the current stream. io: dict of static device buffers."""
        b = io["z"].shape[0]
        cf = guidance_post > 1.0
        # z_N = noise * sigma_max and scale_model_input (consistencytta.py:160,173; heun:151-172) are one device scalar
        x0 = io["x0"]
        with ops.operand_dtype(self.stage_dtype["unet"]):
            ops.nchw_to_nhwc(io["z"], scale_dev=io["scale"], out=x0[:b])
            if cf:
                ops.nchw_to_nhwc(io["z"], scale_dev=io["scale"], out=x0[b:])  # torch.cat([z_n] * 2), consistencytta.py:171
            if not reuse_text:   # K/V of the prompt for all 16 cross-attention sites: once per prompt set
                self.unet.project_text(io["enc"], out=io["enc_kv"])
            lat = self.unet.forward_nhwc(x0, io["t"], io["w"], None, kv_len=io["kv_len"], sample_is_nhwc=True,
                                         enc_kv=io["enc_kv"])
        if cf:  # consistencytta.py:182-184
            lat = ops.cfg_mix(lat, float(guidance_post))
        ops.nhwc_to_nchw(lat, out=io["latent"])
        if stages == "unet":
            return
        mel16 = io["mel16"]      # the vocoder's operand: its dtype is the vocoder stage's
        with ops.operand_dtype(self.stage_dtype["vae"]):
            mel = self.vae.decode_nhwc(lat, use_ema=use_ema, z_scale=1.0 / float(self.vae.scale_factor), mel16=mel16)
        refs["mel"] = mel  # fp32 [B,1024,64,1], same memory layout as NCHW [B,1,1024,64]
        if stages == "vae":
            return
        with ops.operand_dtype(self.stage_dtype["vocoder"]):
            wav = self.vae.vocoder.forward_btc(mel16.view(b, mel16.shape[1], mel16.shape[2]))
        refs["wav"] = wav
        ops.wave_to_int16(wav, out=io["i16"])

    def _make_io(self, b, n_text, cf, dev):
        bu = 2 * b if cf else b
        f16 = self.stage_dtype["unet"]
        with ops.operand_dtype(f16):
            kv_cols = self.unet.packed()["kv_all"].n
        return {
            "z": torch.zeros((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32),
            "scale": torch.ones(1, device=dev, dtype=torch.float32),
            "x0": torch.zeros(bu, LATENT_SHAPE[1], LATENT_SHAPE[2], LATENT_SHAPE[0], device=dev, dtype=f16),
            "enc": torch.zeros(bu, n_text, 1024, device=dev, dtype=torch.float32),
            "enc_kv": torch.zeros(bu, n_text, kv_cols, device=dev, dtype=f16),
            "kv_len": torch.full((bu,), n_text, device=dev, dtype=torch.int32),
            "t": torch.zeros(bu, device=dev, dtype=torch.float32),
            "w": torch.zeros(bu, device=dev, dtype=torch.float32),
            "latent": torch.zeros((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32),
            "mel16": torch.zeros(b, 1024, 64, 1, device=dev, dtype=self.stage_dtype["vocoder"]),
            "i16": torch.zeros(b, WAVE_SAMPLES, device=dev, dtype=torch.int16),
        }

    def text_bucket_len(self, n_text):
        return (int(n_text) + self.text_bucket - 1) // self.text_bucket * self.text_bucket

    def _bucket(self, b, n_text, cf, dev):
        key = (b, self.text_bucket_len(n_text), bool(cf), str(dev))
        ent = self._buckets.get(key)
        if ent is None:
            while len(self._buckets) >= self.max_buckets:
                self._buckets.popitem(last=False)      # least recently used: its graphs and buffers are released
            ent = {"io": self._make_io(key[0], key[1], cf, dev), "graphs": {}, "text_id": None}
            self._buckets[key] = ent
        else:
            self._buckets.move_to_end(key)
        return ent

    def fill_inputs(self, io, z, enc, mask, guidance, timestep, in_scale, reuse_text=False):
        """Host->device (or device->device) copies of one batch into the static buffers of a bucket."""
        b = io["z"].shape[0]
        bu = io["enc"].shape[0]
        io["z"].copy_(z, non_blocking=True)
        if not reuse_text:
            n_text = enc.shape[1]
            if enc.shape[0] != bu:
                raise ValueError("text embeddings have %d rows, expected %d (clips%s)"
                                 % (enc.shape[0], bu, " x 2 [uncond ; cond] with post-CFG" if bu != b else ""))
            io["enc"][:, :n_text].copy_(enc, non_blocking=True)
            if mask is not None:
                if self.validate_mask:
                    kv = prefix_mask_lengths(mask)
                else:
                    kv = mask.to(torch.int32).sum(dim=1).to(torch.int32)
                io["kv_len"].copy_(kv, non_blocking=True)
            else:
                io["kv_len"].fill_(n_text)
        if torch.is_tensor(guidance):
            g = guidance.reshape(-1).float()
            io["w"].copy_(g.expand(bu) if g.numel() == 1 else (torch.cat([g, g]) if bu == 2 * g.numel() and bu != b else g))
        else:
            io["w"].fill_(float(guidance))
        io["t"].fill_(float(timestep))
        io["scale"].fill_(float(in_scale))

    def first_step(self):
        """(timesteps[0], input scale of z_N = noise * init_noise_sigma) of the 18-step training schedule the consistency
        model is queried at first (consistencytta.py:159-160,186; audio_consistency_model.py:489-496)."""
        self.scheduler.set_timesteps(18)
        t0 = self.scheduler.timesteps[0]
        return float(t0), float(self.scheduler.init_noise_sigma) * scheduler_input_scale(self.scheduler, t0)

    def run(self, noise, enc, mask, guidance, guidance_post=1.0, timestep=None, sigma=None, use_ema=False,
            stages="all", in_scale=None, reuse_text=False, clone=False, check_overflow=False):
        """`_run_once` + the fp16 overflow guard: with check_overflow=True the outputs are tested for inf / NaN (one host
        synchronisation); the first stage whose output is not finite is escalated to bf16 operands — sticky for this
        engine — and the request is run again.  fp16's range (65504) is the known weak spot of SD-2.x-style UNets (the
        reference config carries upcast_attention for it); bf16 trades 3 mantissa bits for fp32's range."""
        args = (noise, enc, mask, guidance, guidance_post, timestep, sigma, use_ema, stages, in_scale, reuse_text, clone)
        out = self._run_once(*args)
        if not check_overflow:
            return out
        for _ in range(3):
            bad = None
            for stage, key in (("unet", "latent"), ("vae", "mel"), ("vocoder", "wav")):
                if key in out and not bool(torch.isfinite(out[key]).all()):
                    bad = stage
                    break
            if bad is None or self.stage_dtype[bad] == torch.bfloat16:
                break
            if reuse_text:
                raise RuntimeError("fp16 overflow in a re-query (reuse_text=True): re-run the whole request")
            self.escalate(bad)
            out = self._run_once(*args)
        return out

    def _run_once(self, noise, enc, mask, guidance, guidance_post=1.0, timestep=None, sigma=None, use_ema=False,
                  stages="all", in_scale=None, reuse_text=False, clone=False):
        """noise [B,8,256,16] fp32 (host or device); enc [B or 2B, L, 1024]; mask bool [B or 2B, L] or None (prefix masks
        only).  The UNet input is `noise * in_scale` at `timestep`; by default the first query of the sampler, N(0,1)
        noise * sigma_max / sqrt(sigma_max^2 + 1) at timesteps[0] (`sigma=` is the Heun shorthand for
        in_scale = sigma / sqrt(sigma^2 + 1)).  `reuse_text=True` re-queries with the prompt K/V the previous run() of
        this bucket projected (multi-step sampling: one projection per prompt set).
        Returns dict(latent=[B,8,256,16] fp32, mel=[B,1,1024,64] fp32, wav=[B,163872] fp32, int16=[B,163872])."""
        dev = self.unet.device
        if dev.type != "cuda":
            raise RuntimeError("SingleStepEngine needs the models on a CUDA device (no CPU fallback)")
        if stages not in ("unet", "vae", "all"):
            raise ValueError("stages must be 'unet', 'vae' or 'all'")
        if stages != "unet" and self.vae is None:
            raise RuntimeError("this engine has no VAE attached: only stages='unet' is available")
        if in_scale is None and sigma is not None:
            in_scale = float(sigma) / float((float(sigma) ** 2 + 1) ** 0.5)
        if timestep is None or in_scale is None:
            t0, s0 = self.first_step()
            timestep = t0 if timestep is None else timestep
            in_scale = s0 if in_scale is None else in_scale
        b = noise.shape[0]
        if b == 0 or enc.shape[1] == 0:
            raise ValueError("empty request: need at least one clip and one text token (got %d clips, %d tokens)"
                             % (b, enc.shape[1]))
        if tuple(noise.shape[1:]) != LATENT_SHAPE or enc.shape[2] != 1024:
            raise ValueError("noise must be [B, 8, 256, 16] and text embeddings [B, L, 1024], got %s / %s"
                             % (tuple(noise.shape), tuple(enc.shape)))
        if self.max_batch and b > self.max_batch:
            return self._run_micro_batches(noise, enc, mask, guidance, guidance_post, timestep, in_scale, use_ema,
                                           stages, reuse_text)
        with torch.cuda.device(dev):
            self._check_versions()
            cf = guidance_post > 1.0
            ent = self._bucket(b, enc.shape[1], cf, dev)
            io = ent["io"]
            if reuse_text and ent["text_id"] is None:
                raise RuntimeError("reuse_text=True but this bucket holds no projected prompt yet")
            self.fill_inputs(io, noise, enc, mask, guidance, timestep, in_scale, reuse_text)
            if not reuse_text:
                ent["text_id"] = self.text_projections = self.text_projections + 1
            vkey = (float(guidance_post), bool(use_ema), stages, bool(reuse_text))
            var = ent["graphs"].get(vkey)
            if var is None:
                var = ent["graphs"][vkey] = {"graph": None, "launches": None, "refs": {}}
            if not self.use_graphs:
                n0 = ops.launch_count()
                self._compute(io, guidance_post, use_ema, stages, reuse_text, var["refs"])
                var["launches"] = ops.launch_count() - n0
            elif var["graph"] is None:
                # one eager pass (packs weights, sets function attributes), then capture into the shared pool
                self._compute(io, guidance_post, use_ema, stages, reuse_text, var["refs"])
                torch.cuda.synchronize(dev)
                if self._pool is None:
                    self._pool = torch.cuda.graph_pool_handle()
                g = torch.cuda.CUDAGraph()
                n0 = ops.launch_count()
                with torch.cuda.graph(g, pool=self._pool):
                    self._compute(io, guidance_post, use_ema, stages, reuse_text, var["refs"])
                var["launches"] = ops.launch_count() - n0
                var["graph"] = g
                self.captures += 1
                g.replay()
            else:
                var["graph"].replay()
            self.last = (ent, var)
            out = {"latent": io["latent"]}
            if stages in ("vae", "all"):
                out["mel"] = var["refs"]["mel"].view(b, 1, 1024, 64)
            if stages == "all":
                out["wav"] = var["refs"]["wav"]
                out["int16"] = io["i16"]
            if clone:
                out = {k: v.clone() for k, v in out.items()}
            out["launches"] = var["launches"]
            return out

    def _run_micro_batches(self, noise, enc, mask, guidance, guidance_post, timestep, in_scale, use_ema, stages,
                           reuse_text):
        """run() for more clips than `max_batch`: slices of at most max_batch clips reuse the per-size graphs."""
        if reuse_text:
            raise ValueError("reuse_text is per bucket: requests larger than max_batch re-project their prompts")
        dev = self.unet.device
        b = noise.shape[0]
        cf = guidance_post > 1.0   # enc / mask / guidance rows are [unconditional ; conditional] then
        out = {"latent": torch.empty((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32), "launches": 0}
        if stages in ("vae", "all"):
            out["mel"] = torch.empty(b, 1, 1024, 64, device=dev, dtype=torch.float32)
        if stages == "all":
            out["wav"] = torch.empty(b, WAVE_SAMPLES, device=dev, dtype=torch.float32)

        def rows(t, lo, hi):
            return slice_request_rows(t, lo, hi, b, cf)

        for lo in range(0, b, self.max_batch):
            hi = min(lo + self.max_batch, b)
            g = guidance.reshape(-1) if torch.is_tensor(guidance) else guidance
            part = self.run(noise[lo:hi], rows(enc, lo, hi), rows(mask, lo, hi), rows(g, lo, hi), guidance_post,
                            timestep, None, use_ema, stages, in_scale=in_scale)
            out["latent"][lo:hi].copy_(part["latent"])
            if "mel" in out:
                out["mel"][lo:hi].copy_(part["mel"])
            if "wav" in out:
                out["wav"][lo:hi].copy_(part["wav"])
            out["launches"] += part.get("launches") or 0
        if stages == "all":
            with torch.cuda.device(dev):
                out["int16"], _ = ops.wave_to_int16(out["wav"])   # centring over the WHOLE request
        return out


def build_random_init_models(device="cuda", unet_seed=0, vae_seed=1):
    """The synthetic-weight models used by tests, smoke() and bench.py (no checkpoints exist offline)."""
    unet = UNet2DConditionGuidedModel()
    unet.load_state_dict(weights.make_unet_state_dict(unet_seed))
    vae = AutoencoderKL(scale_factor=weights.SCALE_FACTOR)
    vae.load_state_dict(weights.make_vae_state_dict(vae_seed))
    unet.eval().requires_grad_(False)
    vae.eval().requires_grad_(False)
    # operands are packed on the host: the first kernels launched on the GPU are the hot path's own
    return unet.to_prepacked(device), vae.to_prepacked(device)


class TextFrontEnd:
    """Prompt -> FLAN-T5 embeddings + masks (consistencytta.py:84-132 == audio_consistency_model.py:207-263), shared by
    ConsistencyTTA and AudioLCM.  The encoder stays stock PyTorch outside the hot path; this adds a prompt-keyed cache."""

    def _init_text(self, text_encoder, tokenizer, text_cache_size):
        self.text_encoder = text_encoder
        self.tokenizer = tokenizer
        self.text_cache_size = int(text_cache_size)   # prompts whose encoder rows are kept (0: reference behaviour)
        self._text_cache = OrderedDict()
        self.text_encoder_calls = 0

    @torch.no_grad()
    def encode_text(self, prompt, max_length=None, padding=True):
        """consistencytta.py:84-101.  With the prompt cache on (`text_cache_size` > 0) the encoder only sees prompts it
        has not seen before: FLAN-T5's output rows of a prompt's own tokens do not depend on the padding or on the
        other prompts of the batch (padding keys are masked), so each prompt's [n_tokens, 1024] rows are kept and a
        batch is re-assembled from them, zero padded (the padding rows are masked out downstream in either case)."""
        if self.text_encoder is None or self.tokenizer is None:
            raise RuntimeError("no text encoder attached: pass prompt embeddings to generate_from_embeddings(), or "
                               "construct with text_encoder= / tokenizer= (FLAN-T5 weights are not bundled)")
        device = self.text_encoder.device
        if max_length is None:
            max_length = self.tokenizer.model_max_length
        if not self.text_cache_size:
            batch = self.tokenizer(prompt, max_length=max_length, padding=padding, truncation=True, return_tensors="pt")
            input_ids = batch.input_ids.to(device)
            attention_mask = batch.attention_mask.to(device)
            prompt_embeds = self.text_encoder(input_ids=input_ids, attention_mask=attention_mask)[0]
            return prompt_embeds, (attention_mask == 1).to(device)
        prompt = [prompt] if isinstance(prompt, str) else list(prompt)
        cache = self._text_cache
        misses = [p for p in dict.fromkeys(prompt) if (p, max_length) not in cache]
        if misses:
            batch = self.tokenizer(misses, max_length=max_length, padding=True, truncation=True, return_tensors="pt")
            am = batch.attention_mask.to(device)
            emb = self.text_encoder(input_ids=batch.input_ids.to(device), attention_mask=am)[0]
            self.text_encoder_calls += 1
            for i, p in enumerate(misses):
                cache[(p, max_length)] = emb[i, : int(am[i].sum())].clone()
        rows = []
        for p in prompt:
            cache.move_to_end((p, max_length))
            rows.append(cache[(p, max_length)])
        while len(cache) > self.text_cache_size:
            cache.popitem(last=False)
        n = max_length if padding == "max_length" else max(r.shape[0] for r in rows)
        prompt_embeds = rows[0].new_zeros(len(rows), n, rows[0].shape[1])
        mask = torch.zeros(len(rows), n, dtype=torch.bool, device=device)
        for i, r in enumerate(rows):
            prompt_embeds[i, : r.shape[0]] = r
            mask[i, : r.shape[0]] = True
        return prompt_embeds, mask

    @torch.no_grad()
    def encode_text_classifier_free(self, prompt, num_samples_per_prompt):
        """consistencytta.py:104-132."""
        cond_embeds, cond_mask = self.encode_text(prompt)
        cond_embeds = cond_embeds.repeat_interleave(num_samples_per_prompt, 0)
        cond_mask = cond_mask.repeat_interleave(num_samples_per_prompt, 0)
        neg_embeds, neg_mask = self.encode_text([""] * len(prompt), max_length=cond_embeds.shape[1],
                                                padding="max_length")
        neg_embeds = neg_embeds.repeat_interleave(num_samples_per_prompt, 0)
        neg_mask = neg_mask.repeat_interleave(num_samples_per_prompt, 0)
        return torch.cat([neg_embeds, cond_embeds]), torch.cat([neg_mask, cond_mask]), cond_embeds, cond_mask


class ConsistencyTTA(TextFrontEnd, nn.Module):
    """easy_inference/consistencytta.py:12-200.  Construct with already-built models (checkpoint and hub I/O is the
    caller's; `from_checkpoints` reproduces the reference constructor when the files are present)."""

    def __init__(self, unet=None, vae=None, text_encoder=None, tokenizer=None, scheduler=None, use_graphs=True,
                 text_cache_size=1024):
        super().__init__()
        self.unet = unet if unet is not None else UNet2DConditionGuidedModel()
        self.vae = vae if vae is not None else AutoencoderKL(scale_factor=1.0)
        self._init_text(text_encoder, tokenizer, text_cache_size)
        self.scheduler = scheduler or HeunDiscreteScheduler.from_pretrained(
            pretrained_model_name_or_path="stabilityai/stable-diffusion-2-1", subfolder="scheduler")
        self.engine = SingleStepEngine(self.unet, self.vae, self.scheduler, use_graphs=use_graphs)
        self.check_overflow = True    # fp16 -> bf16 escalation of a stage whose output came back non-finite (one host sync)

    @classmethod
    def from_checkpoints(cls, unet_weight_path="consistencytta_clapft_ckpt/unet_state_dict.pt",
                         vae_weight_path="consistencytta_clapft_ckpt/vae_state_dict.pt",
                         unet_config_path="tango_diffusion_light.json", text_encoder_name="google/flan-t5-large"):
        """consistencytta.py:14-58 — needs the checkpoint files and a local FLAN-T5 (not available offline)."""
        from transformers import AutoTokenizer, T5EncoderModel
        cfg = UNet2DConditionGuidedModel.load_config(unet_config_path)
        unet = UNet2DConditionGuidedModel.from_config(cfg, subfolder="unet")
        unet.load_state_dict(torch.load(unet_weight_path, map_location="cpu"))
        vae = AutoencoderKL.from_vae_state_dict_file(vae_weight_path)
        tok = AutoTokenizer.from_pretrained(text_encoder_name)
        enc = T5EncoderModel.from_pretrained(text_encoder_name)
        enc.eval().requires_grad_(False)
        unet.eval().requires_grad_(False)
        vae.eval().requires_grad_(False)
        return cls(unet, vae, enc, tok)

    def train(self, mode=True):
        self.unet.train(mode)
        for m in (self.text_encoder, self.vae):
            if m is not None:
                m.eval()
        return self

    def eval(self):
        return self.train(mode=False)

    def check_eval_mode(self):
        for model, name in zip([self.text_encoder, self.vae, self.unet], ["text_encoder", "vae", "unet"]):
            if model is None:
                continue
            assert model.training is False, f"The {name} is not in eval mode."
            for param in model.parameters():
                assert param.requires_grad is False, f"The {name} is not frozen."

    @torch.no_grad()
    def generate_from_embeddings(self, prompt_embeds, prompt_mask=None, cfg_scale_input=3.0, cfg_scale_post=1.0,
                                 num_steps=1, noise=None, uncond_embeds=None, uncond_mask=None, return_all=False,
                                 step_noises=None):
        """The hot path of forward() for given text-encoder outputs.  prompt_embeds [B, L, 1024].
        `step_noises` (optional list of [B,8,256,16] tensors) replaces the `torch.randn_like` draws of the multi-step
        re-noising (consistencytta.py:193) so that a run can be reproduced against the oracle.
        Returns the int16 clips as a device tensor [B, 163872] (`return_all`: dict with latent / mel / wav / int16)."""
        self.check_eval_mode()
        dev = self.unet.device
        b = prompt_embeds.shape[0]
        use_cf = cfg_scale_post > 1.0
        if use_cf:
            if uncond_embeds is None:
                raise ValueError("cfg_scale_post > 1 needs the unconditional embeddings")
            enc = torch.cat([uncond_embeds, prompt_embeds])
            mask = None if prompt_mask is None else torch.cat([uncond_mask, prompt_mask])
        else:
            enc, mask = prompt_embeds, prompt_mask
        if noise is None:
            noise = torch.randn((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32)  # randn_tensor, :153-155
        eng, sched = self.engine, self.scheduler
        t0, s0 = eng.first_step()
        # multi-step consistency sampling (consistencytta.py:192-197): re-noise zhat_0 and re-query the SAME captured
        # graph (sigma and timestep are device scalars) with the prompt K/V projected once; the last query carries the
        # VAE decode and the vocoder
        sched.set_timesteps(num_steps)
        later = list(sched.timesteps[1::sched.order]) if num_steps > 1 else []
        if b > (eng.max_batch or b):
            later_reuse = False   # micro-batched requests re-project per slice
        else:
            later_reuse = True
        for attempt in range(4):
            out = eng.run(noise, enc, mask, cfg_scale_input, cfg_scale_post, t0, in_scale=s0,
                          stages="unet" if later else "all", check_overflow=self.check_overflow)
            for i, t in enumerate(later):
                zhat = out["latent"]
                n_i = step_noises[i].to(zhat) if step_noises is not None else torch.randn_like(zhat)
                zn = sched.add_noise(zhat, n_i, t)
                out = eng.run(zn, enc, mask, cfg_scale_input, cfg_scale_post, float(t), in_scale=scheduler_input_scale(sched, t),
                              stages="all" if i == len(later) - 1 else "unet", reuse_text=later_reuse)
            if not (self.check_overflow and later):
                break
            # a re-query cannot escalate by itself (its prompt K/V belong to the old operand type): test the final outputs
            bad = next((st for st, key in (("unet", "latent"), ("vae", "mel"), ("vocoder", "wav"))
                        if key in out and not bool(torch.isfinite(out[key]).all())), None)
            if bad is None or eng.stage_dtype[bad] == torch.bfloat16:
                break
            eng.escalate(bad)
        out = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in out.items()}   # engine buffers are reused
        return out if return_all else out["int16"]

    def forward(self, prompt, cfg_scale_input=3.0, cfg_scale_post=1.0, num_steps=1, num_samples=1, sr=16000):
        """consistencytta.py:135-200 -> numpy int16 [B, 9.5 * sr]."""
        embeds_cf, mask_cf, embeds, mask = self.encode_text_classifier_free(prompt, num_samples)
        nb = embeds.shape[0]
        res = self.generate_from_embeddings(embeds.float(), mask, cfg_scale_input, cfg_scale_post, num_steps,
                                            uncond_embeds=embeds_cf[:nb].float(), uncond_mask=mask_cf[:nb])
        wav = res.cpu().numpy() if torch.is_tensor(res) else res
        return wav[:, : int(sr * 9.5)]


class AudioLCM(TextFrontEnd, nn.Module):
    """Inference-side mirror of models/audio_consistency_model.py: student / EMA UNet bookkeeping,
    `load_pretrained` key remap (:160-204) and `inference()` (:429-548).  Training forward() is out of scope."""

    def __init__(self, text_encoder=None, tokenizer=None, unet_model_config_path=None, use_edm=True,
                 text_cache_size=1024, vae=None, **unused):
        super().__init__()
        cfg = UNet2DConditionGuidedModel.load_config(unet_model_config_path) if unet_model_config_path else {}
        self.student_target_unet = UNet2DConditionGuidedModel.from_config(cfg, subfolder="unet")
        self.student_ema_unet = UNet2DConditionGuidedModel.from_config(cfg, subfolder="unet")
        self._init_text(text_encoder, tokenizer, text_cache_size)
        self.use_edm = use_edm
        self._engines = {}
        self.vae = vae
        self.eval().requires_grad_(False)

    def load_pretrained(self, state_dict, strict=True):
        new_sd = OrderedDict()
        for key, val in state_dict.items():
            if "consistency_unet" in key or "diffusion_unet" in key or "vae." in key or "loss." in key:
                continue  # student (trainable copy), teacher, VAE and loss modules are not used at inference
            if "consistency_ema_" in key:
                aft = key.split("consistency_ema_")[-1]
                new_sd["student_target_" + aft] = val
                new_sd.setdefault("student_ema_" + aft, val)
            elif "consistency_slow_ema_" in key:
                new_sd["student_ema_" + key.split("consistency_slow_ema_")[-1]] = val
            elif key.startswith(("student_target_unet.", "student_ema_unet.")):
                new_sd[key] = val
        for u in (self.student_target_unet, self.student_ema_unet):
            u._invalidate()
        return self.load_state_dict(new_sd, strict=False)

    def check_eval_mode(self):
        for u in (self.student_target_unet, self.student_ema_unet):
            assert u.training is False, "The unet is not in eval mode."

    def _engine(self, use_ema, scheduler):
        unet = self.student_ema_unet if use_ema else self.student_target_unet
        key = (bool(use_ema), id(scheduler))
        if key not in self._engines:
            self._engines[key] = SingleStepEngine(unet, self.vae, scheduler)
        return self._engines[key]

    @torch.no_grad()
    def inference_from_embeddings(self, prompt_embeds, prompt_mask, inference_scheduler, guidance_scale_input=3,
                                  guidance_scale_post=1, num_steps=1, use_ema=True, uncond_embeds=None,
                                  uncond_mask=None, noise=None, step_noises=None):
        """`inference()` after the text encoder (audio_consistency_model.py:476-507): returns zhat_0 [B,8,256,16].
        Works with either scheduler `inference.py:159-162` builds: HeunDiscreteScheduler (EDM: z_N = noise * sigma_max,
        input / sqrt(sigma^2 + 1), every 2nd timestep) or DDIMScheduler (z_N = noise, identity input scaling, t0 = 935,
        alpha-product re-noising, every timestep) — the stride follows `self.use_edm` exactly like the reference."""
        self.check_eval_mode()
        eng = self._engine(use_ema, inference_scheduler)
        b = prompt_embeds.shape[0]
        dev = eng.unet.device
        use_cf = guidance_scale_post > 1.0
        enc = torch.cat([uncond_embeds, prompt_embeds]) if use_cf else prompt_embeds
        mask = (torch.cat([uncond_mask, prompt_mask]) if use_cf else prompt_mask) if prompt_mask is not None else None
        if noise is None:
            noise = torch.randn((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32)
        t0, s0 = eng.first_step()
        out = eng.run(noise, enc, mask, guidance_scale_input, guidance_scale_post, t0, in_scale=s0, stages="unet")
        inference_scheduler.set_timesteps(num_steps)
        order = 2 if self.use_edm else 1
        reuse = b <= (eng.max_batch or b)
        for i, t in enumerate(inference_scheduler.timesteps[1::order]):
            zhat = out["latent"]
            n_i = step_noises[i].to(zhat) if step_noises is not None else torch.randn_like(zhat)
            zn = inference_scheduler.add_noise(zhat, n_i, t)
            out = eng.run(zn, enc, mask, guidance_scale_input, guidance_scale_post, float(t),
                          in_scale=scheduler_input_scale(inference_scheduler, t), stages="unet", reuse_text=reuse)
        return out["latent"].clone()

    @torch.no_grad()
    def inference(self, prompt, inference_scheduler, guidance_scale_input=3, guidance_scale_post=1, num_steps=20,
                  use_edm=False, num_samples=1, use_ema=True, query_teacher=False, num_teacher_steps=18,
                  return_all=False):
        if query_teacher:
            raise NotImplementedError("the diffusion teacher loop (demo.py) is outside the single-step hot path")
        # like the reference, the `use_edm` ARGUMENT is not read (audio_consistency_model.py:429-507 only consults
        # self.use_edm for the timestep stride); the scheduler object passed in decides the prologue
        t_start = time()
        embeds_cf, mask_cf, embeds, mask = self.encode_text_classifier_free(prompt, num_samples)
        nb = embeds.shape[0]
        z = self.inference_from_embeddings(embeds.float(), mask, inference_scheduler, guidance_scale_input,
                                           guidance_scale_post, num_steps, use_ema, embeds_cf[:nb].float(),
                                           mask_cf[:nb])
        if return_all:
            torch.cuda.synchronize()
            return z, None, time() - t_start, None
        return z


class AudioLCM_FTVAE(AudioLCM):
    """Inference-side mirror of models/audio_consistency_model_ftvae.py: the checkpoint additionally carries the
    fine-tuned VAE decoder / post_quant_conv and their EMA copies, which `inference.py:204-206` selects with
    `vae.decode_first_stage(latents, use_ema=True)`."""

    def __init__(self, *args, vae=None, **kw):
        super().__init__(*args, vae=vae, **kw)
        if self.vae is None:
            raise ValueError("AudioLCM_FTVAE needs the VAE whose decoder was fine-tuned (vae=)")
        self.vae.enable_ema_modules()
        self.ema_vae_decoder = self.vae.ema_decoder
        self.ema_vae_pqconv = self.vae.ema_post_quant_conv

    def load_pretrained(self, state_dict, strict=True):
        """audio_consistency_model_ftvae.py:69-91: the four VAE modules are filled from the keys that contain their name;
        a `loss.`-prefixed duplicate only fills what the plain key did not."""
        info = super().load_pretrained(state_dict, strict)
        for module, name in zip([self.vae.decoder, self.vae.post_quant_conv, self.ema_vae_decoder, self.ema_vae_pqconv],
                                ["vae.decoder", "vae.post_quant_conv", "ema_vae_decoder", "ema_vae_pqconv"]):
            off = len(name) + 1
            module_sd = {}
            for key, val in state_dict.items():
                if name in key:
                    if "loss" in key:
                        new_key = key[5:]
                        module_sd.setdefault(new_key[off:], val)
                    else:
                        module_sd[key[off:]] = val
            module.load_state_dict(module_sd)
        self.vae._drop_post_quant()
        self.vae.ema_decoder = self.ema_vae_decoder
        self.vae.ema_post_quant_conv = self.ema_vae_pqconv
        return info
