"""Single-step ConsistencyTTA generation restated: scheduler prologue -> UNet -> VAE decode -> vocoder.

Follows easy_inference/consistencytta.py:135-200 and models/audio_consistency_model.py:430-507, with the text
encoder replaced by given embeddings (FLAN-T5 is outside the hot path).
"""
import numpy as np
import torch

from . import hifigan, unet, vae


def heun_sigmas(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    """HeunDiscreteScheduler.__init__ + set_timesteps (scaled_linear), scheduling_heun_discrete.py:100-227.
    Returns float64 sigma per training timestep."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    return (((1 - alphas_cumprod) / alphas_cumprod) ** 0.5).numpy()


def heun_first_step(num_inference_steps=18, num_train_timesteps=1000):
    """(timesteps[0], init_noise_sigma) after set_timesteps(18): (999.0, 14.6146...)."""
    sig = heun_sigmas(num_train_timesteps)
    timesteps = np.linspace(0, num_train_timesteps - 1, num_inference_steps, dtype=float)[::-1].copy()
    sigmas = np.interp(timesteps, np.arange(0, len(sig)), sig).astype(np.float32)
    return float(timesteps[0]), float(sigmas.max())


def generate(unet_sd, vae_sd, scale_factor, noise, enc, enc_mask, guidance, guidance_post=1.0):
    """Returns (latent [B,8,256,16], mel [B,1,1024,64], float waveform [B,163872])."""
    t0, sigma = heun_first_step()
    z = noise * sigma  # consistencytta.py:160
    use_cf = guidance_post > 1.0
    z_in = torch.cat([z] * 2) if use_cf else z
    z_in = z_in / ((sigma ** 2 + 1) ** 0.5)  # scale_model_input, scheduling_heun_discrete.py:151-172
    out = unet.unet_forward(unet_sd, z_in, torch.tensor(t0, dtype=torch.float64), guidance, enc, enc_mask)
    if use_cf:
        u, c = out.chunk(2)
        out = (1 - guidance_post) * u + guidance_post * c
    mel = vae.decode_first_stage(vae_sd, out.float(), scale_factor)
    wav = hifigan.decode_to_waveform(vae_sd, mel, return_float=True)
    return out, mel, wav


def heun_schedule(num_inference_steps, num_train_timesteps=1000):
    """(timesteps, sigmas) after set_timesteps(n), interleaved as the second-order scheduler stores them
    (scheduling_heun_discrete.py:174-227): timesteps [t0, t1, t1, t2, t2, ...], sigmas [s0, s1, s1, ..., 0]."""
    sig = heun_sigmas(num_train_timesteps)
    ts = np.linspace(0, num_train_timesteps - 1, num_inference_steps, dtype=float)[::-1].copy()
    sg = np.interp(ts, np.arange(0, len(sig)), sig).astype(np.float32)
    sg = np.concatenate([sg, [0.0]]).astype(np.float32)
    timesteps = np.concatenate([ts[:1], np.repeat(ts[1:], 2)])
    sigmas = np.concatenate([sg[:1], np.repeat(sg[1:-1], 2), sg[-1:]])
    return timesteps, sigmas


def heun_index_for_timestep(timesteps, t):
    """scheduling_heun_discrete.py:137-149 with the scheduler in first-order state: `(mask * arange).argmax()`, i.e. the
    LAST position where timesteps == t (position 0 when only the first entry matches)."""
    hits = np.nonzero(timesteps == t)[0]
    assert len(hits) > 0, t
    return int(hits[-1])


def generate_latent_multistep(unet_sd, noise, step_noises, enc, enc_mask, guidance, num_steps, guidance_post=1.0,
                              enc_uncond=None, mask_uncond=None):
    """Multi-step consistency sampling, easy_inference/consistencytta.py:186-197: query at timesteps[0] of the 18-step
    schedule, then for every t in set_timesteps(num_steps).timesteps[1::2]: re-noise zhat_0 with `add_noise`
    (z + n * sigma_t, scheduling_heun_discrete.py:364-385), `scale_model_input`, query again."""
    use_cf = guidance_post > 1.0
    if use_cf:
        enc = torch.cat([enc_uncond, enc])
        enc_mask = torch.cat([mask_uncond, enc_mask])

    def query(z_n, sig, t):
        z_in = torch.cat([z_n] * 2) if use_cf else z_n
        z_in = z_in / ((sig ** 2 + 1) ** 0.5)
        out = unet.unet_forward(unet_sd, z_in, torch.tensor(float(t), dtype=torch.float64), guidance, enc, enc_mask)
        if use_cf:
            u, c = out.chunk(2)
            out = (1 - guidance_post) * u + guidance_post * c
        return out

    t0, sigma = heun_first_step()
    zhat = query(noise * sigma, sigma, t0)
    timesteps, sigmas = heun_schedule(num_steps)
    for i, t in enumerate(timesteps[1::2]):
        sig_t = float(sigmas[heun_index_for_timestep(timesteps, t)])
        zhat = query(zhat + step_noises[i] * sig_t, sig_t, t)
    return zhat


def ddim_alphas_cumprod(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(num_inference_steps, num_train_timesteps=1000):
    """scheduling_ddim.py:119-142."""
    ratio = num_train_timesteps // num_inference_steps
    return (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)


def generate_latent_multistep_ddim(unet_sd, noise, step_noises, enc, enc_mask, guidance, num_steps):
    """models/audio_consistency_model.py:486-507 with the DDIM scheduler `inference.py:159-162` picks without --use_edm
    (the model's use_edm False -> stride 1): z_N = noise (init_noise_sigma 1), identity scale_model_input, first query
    at set_timesteps(18).timesteps[0] = 935, re-noising sqrt(abar_t) zhat + sqrt(1 - abar_t) n
    (scheduling_ddim.py:75,83-95,372-393)."""
    ac = ddim_alphas_cumprod()
    t0 = int(ddim_timesteps(18)[0])
    zhat = unet.unet_forward(unet_sd, noise, torch.tensor(float(t0), dtype=torch.float64), guidance, enc, enc_mask)
    for i, t in enumerate(ddim_timesteps(num_steps)[1::1]):
        a = float(ac[int(t)])
        zn = (a ** 0.5) * zhat + ((1 - a) ** 0.5) * step_noises[i]
        zhat = unet.unet_forward(unet_sd, zn, torch.tensor(float(t), dtype=torch.float64), guidance, enc, enc_mask)
    return zhat
