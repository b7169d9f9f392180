"""Minimal HeunDiscreteScheduler / DDIMScheduler with the surface the generation path uses.

Mirrors diffusers/schedulers/scheduling_heun_discrete.py:100-227,364-385 (constructor, set_timesteps,
init_noise_sigma, timesteps, sigmas, scale_model_input, add_noise, index_for_timestep).  DDIMScheduler is the
non-EDM variant `inference.py:159-162` selects without --use_edm: diffusers/schedulers/scheduling_ddim.py:34-95,119-142,
372-393 (init_noise_sigma = 1, identity scale_model_input, t0 = 935 after set_timesteps(18), alpha-product add_noise).
Scalar host math (numpy); the per-sample scaling itself is folded into the first kernel of the UNet by the engine
(`input_scale`).
"""
import numpy as np
import torch

SD21_SCHEDULER_CONFIG = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                             beta_schedule="scaled_linear", prediction_type="v_prediction")


class HeunDiscreteScheduler:
    order = 2

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="linear",
                 trained_betas=None, prediction_type="epsilon", use_karras_sigmas=False):
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        if use_karras_sigmas:
            raise NotImplementedError("use_karras_sigmas is not used by ConsistencyTTA")
        self.config = type("Config", (), dict(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                               beta_end=beta_end, beta_schedule=beta_schedule,
                                               prediction_type=prediction_type))()
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.set_timesteps(num_train_timesteps, None, num_train_timesteps)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, subfolder=None, **kw):
        """No hub access here: 'stabilityai/stable-diffusion-2-1' scheduler constants are built in (train.sh:37)."""
        return cls(**SD21_SCHEDULER_CONFIG)

    @property
    def state_in_first_order(self):
        return True

    def set_timesteps(self, num_inference_steps, device=None, num_train_timesteps=None):
        self.num_inference_steps = num_inference_steps
        num_train_timesteps = num_train_timesteps or self.config.num_train_timesteps
        timesteps = np.linspace(0, num_train_timesteps - 1, num_inference_steps, dtype=float)[::-1].copy()
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        sigmas = np.concatenate([sigmas, [0.0]]).astype(np.float32)
        sigmas = torch.from_numpy(sigmas).to(device=device)
        self.sigmas = torch.cat([sigmas[:1], sigmas[1:-1].repeat_interleave(2), sigmas[-1:]])
        self.init_noise_sigma = self.sigmas.max()
        timesteps = torch.from_numpy(timesteps)
        self.timesteps = torch.cat([timesteps[:1], timesteps[1:].repeat_interleave(2)]).to(device)

    def index_for_timestep(self, timestep):
        t = torch.as_tensor(timestep).reshape(-1).cpu().double()
        avail = self.timesteps.reshape(1, -1).cpu().double()
        mask = avail == t.reshape(-1, 1)
        assert (mask.sum(dim=1) != 0).all(), f"timestep: {t.tolist()}"
        return (mask * torch.arange(mask.shape[1]).reshape(1, -1)).argmax(dim=1).numpy()

    def sigma_for_timestep(self, timestep):
        return self.sigmas.cpu()[self.index_for_timestep(timestep)]

    def scale_model_input(self, sample, timestep):
        sigma = self.sigma_for_timestep(timestep).reshape(-1, 1, 1, 1).to(sample.device)
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def input_scale(self, timestep):
        """The scalar `scale_model_input` multiplies by at `timestep` (the engine folds it into its first kernel)."""
        sigma = float(self.sigma_for_timestep(timestep).reshape(-1)[0])
        return 1.0 / (sigma ** 2 + 1) ** 0.5

    def add_noise(self, original_samples, noise, timesteps):
        sigma = self.sigma_for_timestep(timesteps).reshape(-1, 1, 1, 1).to(original_samples.device)
        return original_samples + noise * sigma


class DDIMScheduler:
    """scheduling_ddim.py — what AudioLCM.inference touches when the model was not trained with EDM (use_edm False)."""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 trained_betas=None, prediction_type="epsilon", **unused):
        if trained_betas is not None:
            self.betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.config = type("Config", (), dict(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                               beta_end=beta_end, beta_schedule=beta_schedule,
                                               prediction_type=prediction_type))()
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self.num_train_timesteps = num_train_timesteps

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, subfolder=None, **kw):
        return cls(**SD21_SCHEDULER_CONFIG)

    def set_timesteps(self, num_inference_steps, device=None):
        if num_inference_steps > self.config.num_train_timesteps:
            raise ValueError(
                f"`num_inference_steps`: {num_inference_steps} cannot be larger than `self.config.train_timesteps`:"
                f" {self.config.num_train_timesteps} as the unet model trained with this scheduler can only handle"
                f" maximal {self.config.num_train_timesteps} timesteps.")
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy()
        self.timesteps = torch.from_numpy(timesteps.astype(np.int64)).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def input_scale(self, timestep=None):
        return 1.0

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        t = torch.as_tensor(timesteps).to(original_samples.device).long().reshape(-1)
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return (ac[t] ** 0.5).reshape(shape) * original_samples + ((1 - ac[t]) ** 0.5).reshape(shape) * noise
