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
from .unet import UNet2DConditionGuidedModel
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


class SingleStepEngine:
    """One bucket of static buffers + one CUDA graph per (batch, text length, sigma, post-CFG, stage) key.

    `max_batch` bounds the clips resident in one pass (about 0.47 GB of activations per clip at the widest point, so 64
    clips = 30 GB): larger requests run as micro-batches through the same graphs and are stitched together, with the
    batch-GLOBAL waveform centring of `vocoder_infer` (hifigan/utilities.py:84-86) applied over the whole request, so
    the int16 result does not depend on how the request was split."""

    def __init__(self, unet, vae, scheduler=None, use_graphs=True, max_batch=64):
        self.unet = unet
        self.vae = vae
        self.scheduler = scheduler or HeunDiscreteScheduler.from_pretrained()
        self.use_graphs = use_graphs
        self.max_batch = max_batch
        self._graphs = {}

    # -------------------------------------------------------------------------------------------- hot path
    def _compute(self, io, sigma, guidance_post, use_ema, stages):
        """Launches every kernel of the path on the current stream. io: dict of static device buffers."""
        b = io["noise"].shape[0]
        cf = guidance_post > 1.0
        f16 = ops.OPERAND_DTYPE
        # z_N = noise * sigma_max, scale_model_input: / sqrt(sigma^2 + 1)  (consistencytta.py:160,173; heun:151-172)
        scale = float(sigma) / float((sigma ** 2 + 1) ** 0.5)
        bu = 2 * b if cf else b
        x0 = io["x0"]
        ops.nchw_to_nhwc(io["noise"], scale=scale, out=x0[:b])
        if cf:
            ops.nchw_to_nhwc(io["noise"], scale=scale, out=x0[b:])  # torch.cat([z_n] * 2), consistencytta.py:171
        lat = self.unet.forward_nhwc(x0, io["t"], io["w"], io["enc"], kv_len=io["kv_len"], sample_is_nhwc=True)
        if cf:  # consistencytta.py:182-184
            lat = ops.cfg_mix(lat, float(guidance_post))
        ops.nhwc_to_nchw(lat, out=io["latent"])
        if stages == "unet":
            return
        mel16 = io["mel16"]
        mel = self.vae.decode_nhwc(lat, use_ema=use_ema, z_scale=1.0 / float(self.vae.scale_factor), mel16=mel16)
        io["mel_ref"] = mel  # fp32 [B,1024,64,1], same memory layout as NCHW [B,1,1024,64]
        if stages == "vae":
            return
        wav = self.vae.vocoder.forward_btc(mel16.view(b, mel16.shape[1], mel16.shape[2]))
        io["wav_ref"] = wav
        ops.wave_to_int16(wav, out=io["i16"])

    def _make_io(self, b, n_text, cf, dev):
        bu = 2 * b if cf else b
        f16 = ops.OPERAND_DTYPE
        return {
            "noise": torch.zeros((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32),
            "x0": torch.zeros(bu, LATENT_SHAPE[1], LATENT_SHAPE[2], LATENT_SHAPE[0], device=dev, dtype=f16),
            "enc": torch.zeros(bu, n_text, 1024, device=dev, dtype=torch.float32),
            "kv_len": torch.full((bu,), n_text, device=dev, dtype=torch.int32),
            "t": torch.zeros(bu, device=dev, dtype=torch.float32),
            "w": torch.zeros(bu, device=dev, dtype=torch.float32),
            "latent": torch.zeros((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32),
            "mel16": torch.zeros(b, 1024, 64, 1, device=dev, dtype=f16),
            "i16": torch.zeros(b, WAVE_SAMPLES, device=dev, dtype=torch.int16),
        }

    def _bucket(self, b, n_text, sigma, guidance_post, use_ema, stages, dev):
        key = (b, n_text, float(sigma), float(guidance_post), bool(use_ema), stages, str(dev))
        ent = self._graphs.get(key)
        if ent is None:
            io = self._make_io(b, n_text, guidance_post > 1.0, dev)
            ent = {"io": io, "graph": None, "warm": 0}
            self._graphs[key] = ent
        return ent

    def fill_inputs(self, io, noise, enc, mask, guidance, timestep):
        """Host->device (or device->device) copies of one batch into the static buffers of a bucket."""
        b = io["noise"].shape[0]
        bu = io["enc"].shape[0]
        io["noise"].copy_(noise, non_blocking=True)
        io["enc"].copy_(enc, non_blocking=True)
        if mask is not None:
            io["kv_len"].copy_(mask.to(torch.int32).sum(dim=1), non_blocking=True)
        else:
            io["kv_len"].fill_(io["enc"].shape[1])
        if torch.is_tensor(guidance):
            g = guidance.reshape(-1).float()
            io["w"].copy_(g.expand(bu) if g.numel() == 1 else (torch.cat([g, g]) if bu == 2 * g.numel() and bu != b else g))
        else:
            io["w"].fill_(float(guidance))
        io["t"].fill_(float(timestep))

    def run(self, noise, enc, mask, guidance, guidance_post=1.0, timestep=None, sigma=None, use_ema=False,
            stages="all"):
        """noise [B,8,256,16] fp32 (N(0,1), host or device); enc [B or 2B, L, 1024]; mask bool [B or 2B, L] or None.
        Returns dict(latent=[B,8,256,16] fp32, mel=[B,1,1024,64] fp32, wav=[B,163872] fp32, int16=[B,163872])."""
        dev = self.unet.device
        if dev.type != "cuda":
            raise RuntimeError("SingleStepEngine needs the models on a CUDA device (no CPU fallback)")
        if timestep is None or sigma is None:
            self.scheduler.set_timesteps(18)
            timestep = float(self.scheduler.timesteps[0])
            sigma = float(self.scheduler.init_noise_sigma)
        b = noise.shape[0]
        if self.max_batch and b > self.max_batch:
            return self._run_micro_batches(noise, enc, mask, guidance, guidance_post, timestep, sigma, use_ema, stages)
        ent = self._bucket(b, enc.shape[1], sigma, guidance_post, use_ema, stages, dev)
        io = ent["io"]
        self.fill_inputs(io, noise, enc, mask, guidance, timestep)
        if not self.use_graphs:
            self._compute(io, sigma, guidance_post, use_ema, stages)
        elif ent["graph"] is None:
            # one eager pass (packs weights, sets function attributes), then capture
            self._compute(io, sigma, guidance_post, use_ema, stages)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                self._compute(io, sigma, guidance_post, use_ema, stages)
            ent["launches"] = ops.launch_count() - n0
            ent["graph"] = g
            g.replay()
        else:
            ent["graph"].replay()
        out = {"latent": io["latent"]}
        if stages in ("vae", "all"):
            out["mel"] = io["mel_ref"].view(b, 1, 1024, 64)
        if stages == "all":
            out["wav"] = io["wav_ref"]
            out["int16"] = io["i16"]
        out["launches"] = ent.get("launches")
        return out


    def _run_micro_batches(self, noise, enc, mask, guidance, guidance_post, timestep, sigma, use_ema, stages):
        """run() for more clips than `max_batch`: slices of at most max_batch clips reuse the per-size graphs."""
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
                            timestep, sigma, use_ema, stages)
            out["latent"][lo:hi].copy_(part["latent"])
            if "mel" in out:
                out["mel"][lo:hi].copy_(part["mel"])
            if "wav" in out:
                out["wav"][lo:hi].copy_(part["wav"])
            out["launches"] += part.get("launches") or 0
        if stages == "all":
            out["int16"], _ = ops.wave_to_int16(out["wav"])   # centring over the WHOLE request
        return out


def _random_init(module, schema_fn, seed):
    module.load_state_dict(weights.make_state_dict(schema_fn(), seed), strict=True)
    return module


def build_random_init_models(device="cuda", unet_seed=0, vae_seed=1):
    """The synthetic-weight models used by tests, smoke() and bench.py (no checkpoints exist offline)."""
    unet = UNet2DConditionGuidedModel()
    unet.load_state_dict(weights.make_unet_state_dict(unet_seed))
    vae = AutoencoderKL(scale_factor=weights.SCALE_FACTOR)
    vae.load_state_dict(weights.make_vae_state_dict(vae_seed))
    unet.eval().requires_grad_(False)
    vae.eval().requires_grad_(False)
    return unet.to(device), vae.to(device)


class ConsistencyTTA(nn.Module):
    """easy_inference/consistencytta.py:12-200.  Construct with already-built models (checkpoint and hub I/O is the
    caller's; `from_checkpoints` reproduces the reference constructor when the files are present)."""

    def __init__(self, unet=None, vae=None, text_encoder=None, tokenizer=None, scheduler=None, use_graphs=True):
        super().__init__()
        self.unet = unet if unet is not None else UNet2DConditionGuidedModel()
        self.vae = vae if vae is not None else AutoencoderKL(scale_factor=1.0)
        self.text_encoder = text_encoder
        self.tokenizer = tokenizer
        self.scheduler = scheduler or HeunDiscreteScheduler.from_pretrained(
            pretrained_model_name_or_path="stabilityai/stable-diffusion-2-1", subfolder="scheduler")
        self.engine = SingleStepEngine(self.unet, self.vae, self.scheduler, use_graphs=use_graphs)

    @classmethod
    def from_checkpoints(cls, unet_weight_path="consistencytta_clapft_ckpt/unet_state_dict.pt",
                         vae_weight_path="consistencytta_clapft_ckpt/vae_state_dict.pt",
                         unet_config_path="tango_diffusion_light.json", text_encoder_name="google/flan-t5-large"):
        """consistencytta.py:14-58 — needs the checkpoint files and a local FLAN-T5 (not available offline)."""
        from transformers import AutoTokenizer, T5EncoderModel
        cfg = UNet2DConditionGuidedModel.load_config(unet_config_path)
        unet = UNet2DConditionGuidedModel.from_config(cfg, subfolder="unet")
        unet.load_state_dict(torch.load(unet_weight_path, map_location="cpu"))
        raw = torch.load(vae_weight_path, map_location="cpu")
        vae = AutoencoderKL(scale_factor=raw["scale_factor"])
        vae.load_state_dict(raw["state_dict"])
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
    def encode_text(self, prompt, max_length=None, padding=True):
        """consistencytta.py:84-101."""
        if self.text_encoder is None or self.tokenizer is None:
            raise RuntimeError("no text encoder attached: pass prompt embeddings to generate_from_embeddings(), or "
                               "construct with text_encoder= / tokenizer= (FLAN-T5 weights are not bundled)")
        device = self.text_encoder.device
        if max_length is None:
            max_length = self.tokenizer.model_max_length
        batch = self.tokenizer(prompt, max_length=max_length, padding=padding, truncation=True, return_tensors="pt")
        input_ids = batch.input_ids.to(device)
        attention_mask = batch.attention_mask.to(device)
        prompt_embeds = self.text_encoder(input_ids=input_ids, attention_mask=attention_mask)[0]
        return prompt_embeds, (attention_mask == 1).to(device)

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

    @torch.no_grad()
    def generate_from_embeddings(self, prompt_embeds, prompt_mask=None, cfg_scale_input=3.0, cfg_scale_post=1.0,
                                 num_steps=1, noise=None, uncond_embeds=None, uncond_mask=None, return_all=False,
                                 step_noises=None):
        """The hot path of forward() for given text-encoder outputs.  prompt_embeds [B, L, 1024].
        `step_noises` (optional list of [B,8,256,16] tensors) replaces the `torch.randn_like` draws of the multi-step
        re-noising (consistencytta.py:193) so that a run can be reproduced against the oracle."""
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
        self.scheduler.set_timesteps(18)
        t0 = float(self.scheduler.timesteps[0])
        sigma = float(self.scheduler.init_noise_sigma)
        if num_steps == 1:
            out = self.engine.run(noise, enc, mask, cfg_scale_input, cfg_scale_post, t0, sigma)
            return out if return_all else out["int16"]
        # multi-step consistency sampling (consistencytta.py:192-197): re-noise and re-query the same UNet graph
        out = self.engine.run(noise, enc, mask, cfg_scale_input, cfg_scale_post, t0, sigma, stages="unet")
        zhat = out["latent"].clone()
        self.scheduler.set_timesteps(num_steps)
        for i, t in enumerate(self.scheduler.timesteps[1::2]):
            sig_t = float(self.scheduler.sigma_for_timestep(t))
            # add_noise then scale_model_input: (zhat + n * sig_t) / sqrt(sig_t^2 + 1) == noise' * sig' form of run()
            n_i = step_noises[i].to(zhat) if step_noises is not None else torch.randn_like(zhat)
            zn = self.scheduler.add_noise(zhat, n_i, t)
            out = self.engine.run(zn / sig_t if sig_t > 0 else zn, enc, mask, cfg_scale_input, cfg_scale_post,
                                  float(t), sig_t if sig_t > 0 else 1.0, stages="unet")
            zhat = out["latent"].clone()
        mel = self.vae.decode_first_stage(zhat)
        wav = self.vae.decode_to_waveform(mel)
        return {"latent": zhat, "mel": mel, "int16": wav} if return_all else wav

    def forward(self, prompt, cfg_scale_input=3.0, cfg_scale_post=1.0, num_steps=1, num_samples=1, sr=16000):
        """consistencytta.py:135-200 -> numpy int16 [B, 9.5 * sr]."""
        embeds_cf, mask_cf, embeds, mask = self.encode_text_classifier_free(prompt, num_samples)
        nb = embeds.shape[0]
        res = self.generate_from_embeddings(embeds.float(), mask, cfg_scale_input, cfg_scale_post, num_steps,
                                            uncond_embeds=embeds_cf[:nb].float(), uncond_mask=mask_cf[:nb])
        wav = res.cpu().numpy() if torch.is_tensor(res) else res
        return wav[:, : int(sr * 9.5)]


class AudioLCM(nn.Module):
    """Inference-side mirror of models/audio_consistency_model.py: student / EMA UNet bookkeeping,
    `load_pretrained` key remap (:160-204) and `inference()` (:429-548).  Training forward() is out of scope."""

    def __init__(self, text_encoder=None, tokenizer=None, unet_model_config_path=None, use_edm=True, **unused):
        super().__init__()
        cfg = UNet2DConditionGuidedModel.load_config(unet_model_config_path) if unet_model_config_path else {}
        self.student_target_unet = UNet2DConditionGuidedModel.from_config(cfg, subfolder="unet")
        self.student_ema_unet = UNet2DConditionGuidedModel.from_config(cfg, subfolder="unet")
        self.text_encoder = text_encoder
        self.tokenizer = tokenizer
        self.use_edm = use_edm
        self._engines = {}
        self.vae = None
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
                                  uncond_mask=None, noise=None):
        """`inference()` after the text encoder: returns zhat_0 [B,8,256,16]."""
        self.check_eval_mode()
        eng = self._engine(use_ema, inference_scheduler)
        b = prompt_embeds.shape[0]
        dev = eng.unet.device
        use_cf = guidance_scale_post > 1.0
        enc = torch.cat([uncond_embeds, prompt_embeds]) if use_cf else prompt_embeds
        mask = (torch.cat([uncond_mask, prompt_mask]) if use_cf else prompt_mask) if prompt_mask is not None else None
        if noise is None:
            noise = torch.randn((b,) + LATENT_SHAPE, device=dev, dtype=torch.float32)
        inference_scheduler.set_timesteps(18)
        t0 = float(inference_scheduler.timesteps[0])
        sigma = float(inference_scheduler.init_noise_sigma)
        out = eng.run(noise, enc, mask, guidance_scale_input, guidance_scale_post, t0, sigma, stages="unet")
        zhat = out["latent"].clone()
        inference_scheduler.set_timesteps(num_steps)
        order = 2 if self.use_edm else 1
        for t in inference_scheduler.timesteps[1::order]:
            sig_t = float(inference_scheduler.sigma_for_timestep(t))
            zn = inference_scheduler.add_noise(zhat, torch.randn_like(zhat), t)
            out = eng.run(zn / sig_t, enc, mask, guidance_scale_input, guidance_scale_post, float(t), sig_t,
                          stages="unet")
            zhat = out["latent"].clone()
        return zhat

    @torch.no_grad()
    def inference(self, prompt, inference_scheduler, guidance_scale_input=3, guidance_scale_post=1, num_steps=20,
                  use_edm=False, num_samples=1, use_ema=True, query_teacher=False, num_teacher_steps=18,
                  return_all=False):
        if query_teacher:
            raise NotImplementedError("the diffusion teacher loop (demo.py) is outside the single-step hot path")
        helper = ConsistencyTTA.__new__(ConsistencyTTA)
        nn.Module.__init__(helper)
        helper.text_encoder, helper.tokenizer = self.text_encoder, self.tokenizer
        t_start = time()
        embeds_cf, mask_cf, embeds, mask = helper.encode_text_classifier_free(prompt, num_samples)
        nb = embeds.shape[0]
        z = self.inference_from_embeddings(embeds.float(), mask, inference_scheduler, guidance_scale_input,
                                           guidance_scale_post, num_steps, use_ema, embeds_cf[:nb].float(),
                                           mask_cf[:nb])
        if return_all:
            torch.cuda.synchronize()
            return z, None, time() - t_start, None
        return z
