"""Generates tests/golden/* by running the REFERENCE modules (imported from /root/reference/easy_inference, CPU,
fp32) on the deterministic synthetic weights and seeded inputs, and checks the oracle restatement against them.

Run in the build container only (the reference does not exist on the GPU box):
    python -m oracle.make_golden
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CTTA_REFERENCE", "/root/reference/easy_inference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """The two shims of SURVEY.md 8c, then the reference's own classes."""
    sys.path.insert(0, REF)
    import huggingface_hub

    class _HfFolder:
        @staticmethod
        def get_token():
            return None

    huggingface_hub.HfFolder = _HfFolder
    import huggingface_hub.constants as hc
    if not hasattr(hc, "hf_cache_home"):
        hc.hf_cache_home = os.path.expanduser("~/.cache/huggingface")
    from diffusers import HeunDiscreteScheduler, UNet2DConditionGuidedModel
    from audioldm.variational_autoencoder import AutoencoderKL
    from audioldm.utils import default_audioldm_config
    import_reference.DDIM = import_reference_ddim()
    return UNet2DConditionGuidedModel, HeunDiscreteScheduler, AutoencoderKL, default_audioldm_config


def import_reference_ddim():
    """The reference's DDIMScheduler lives only in the FULL tree (/root/reference/diffusers/schedulers/scheduling_ddim.py),
    whose package __init__ does not import under this image's transformers / huggingface_hub.  Its source file is
    executed unmodified, from where it lies, inside a stand-in package whose three relative imports (configuration_utils,
    utils.{BaseOutput, randn_tensor}, scheduling_utils) resolve to the same helpers of the trimmed easy_inference copy."""
    import importlib.util
    import types
    from diffusers.utils import configuration_utils, outputs, scheduling_utils, torch_utils
    pkg = types.ModuleType("refddim")
    pkg.__path__ = []
    sub = types.ModuleType("refddim.schedulers")
    sub.__path__ = []
    utils = types.ModuleType("refddim.utils")
    utils.BaseOutput = outputs.BaseOutput
    utils.randn_tensor = torch_utils.randn_tensor
    sys.modules.update({"refddim": pkg, "refddim.schedulers": sub, "refddim.utils": utils,
                        "refddim.configuration_utils": configuration_utils,
                        "refddim.schedulers.scheduling_utils": scheduling_utils})
    path = os.path.join(os.path.dirname(REF.rstrip("/")), "diffusers", "schedulers", "scheduling_ddim.py")
    spec = importlib.util.spec_from_file_location("refddim.schedulers.scheduling_ddim", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod.DDIMScheduler


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def main():
    from consistencytta_b200 import weights
    from oracle import hifigan as o_hifigan
    from oracle import pipeline as o_pipe
    from oracle import unet as o_unet
    from oracle import vae as o_vae

    torch.manual_seed(0)
    os.makedirs(GOLD, exist_ok=True)
    UNet, Heun, AutoencoderKL, default_cfg = import_reference()

    # ---------------- reference modules with the synthetic weights
    cfg = UNet.load_config(os.path.join(REF, "tango_diffusion_light.json"))
    unet = UNet.from_config(cfg, subfolder="unet").eval().requires_grad_(False)
    vae_cfg = default_cfg("audioldm-s-full")["model"]["params"]["first_stage_config"]["params"]
    vae_cfg["scale_factor"] = weights.SCALE_FACTOR
    vae = AutoencoderKL(**vae_cfg).eval().requires_grad_(False)

    ref_unet_sd = unet.state_dict()
    ref_vae_sd = vae.state_dict()
    schema = {"unet": {k: list(v.shape) for k, v in ref_unet_sd.items()},
              "vae": {k: list(v.shape) for k, v in ref_vae_sd.items()}}
    with open(os.path.join(GOLD, "state_dict_schema.json"), "w") as f:
        json.dump(schema, f, indent=0, sort_keys=True)

    unet_sd = weights.make_unet_state_dict(0)
    vae_sd = weights.make_vae_state_dict(1)
    unet.load_state_dict(unet_sd, strict=True)
    missing, unexpected = vae.load_state_dict(vae_sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("encoder.", "quant_conv.")) for k in missing), missing

    sched = Heun(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="v_prediction")
    sched.set_timesteps(18)
    t0 = sched.timesteps[0]
    sigma = sched.init_noise_sigma
    o_t0, o_sigma = o_pipe.heun_first_step()
    assert float(t0) == o_t0 and abs(float(sigma) - o_sigma) < 1e-6, (t0, sigma, o_t0, o_sigma)

    report = {"t0": float(t0), "sigma_max": float(sigma)}
    out = {}
    with torch.no_grad():
        # ---------------- case A: full chain, B = 1, L = 32, python-float guidance (easy_inference/consistencytta.py:135-200)
        noise, enc, mask = weights.synthetic_inputs(1, 32)
        tic = time.time()
        z_in = sched.scale_model_input(noise * sigma, t0)
        lat = unet(z_in, t0, guidance=4.0, encoder_hidden_states=enc, encoder_attention_mask=mask).sample
        t_unet = time.time() - tic
        tic = time.time()
        mel = vae.decode_first_stage(lat.float())
        t_vae = time.time() - tic
        tic = time.time()
        wav_i16 = vae.decode_to_waveform(mel)
        t_voc = time.time() - tic
        wav_f = vae.vocoder(mel.squeeze(1).permute(0, 2, 1)).squeeze(1).float()
        report["cpu_seconds"] = {"unet": t_unet, "vae": t_vae, "vocoder": t_voc, "threads": torch.get_num_threads()}
        out["a_latent"], out["a_mel"], out["a_wav"], out["a_wav_i16"] = lat, mel, wav_f, torch.from_numpy(wav_i16)

        o_lat, o_mel, o_wav = o_pipe.generate(unet_sd, vae_sd, weights.SCALE_FACTOR, noise, enc, mask, 4.0)
        report["oracle_vs_reference"] = {"latent": rel(o_lat, lat), "mel": rel(o_mel, mel), "wav": rel(o_wav, wav_f)}
        o_i16 = o_hifigan.to_int16(o_wav)
        report["oracle_vs_reference"]["int16_max_abs_diff"] = int(np.abs(o_i16.astype(np.int32) - wav_i16.astype(np.int32)).max())

        # ---------------- case B: UNet only, B = 2, ragged mask, per-sample guidance tensor
        noise, enc, mask = weights.synthetic_inputs(2, 24, seed=4321)
        w = torch.tensor([1.0, 5.0])
        z_in = sched.scale_model_input(noise * sigma, t0)
        lat_b = unet(z_in, t0, guidance=w, encoder_hidden_states=enc, encoder_attention_mask=mask).sample
        out["b_latent"] = lat_b
        o_lat_b = o_unet.unet_forward(unet_sd, z_in, t0, w, enc, mask)
        report["oracle_vs_reference"]["latent_b"] = rel(o_lat_b, lat_b)

        # ---------------- case C: stand-alone vocoder on a realistic mel distribution N(-4.63, 2.74) (audioldm/utils.py:120-121)
        g = torch.Generator().manual_seed(99)
        mel_c = torch.randn(1, 1, 256, 64, generator=g) * 2.74 - 4.63
        wav_c = vae.vocoder(mel_c.squeeze(1).permute(0, 2, 1)).squeeze(1).float()
        out["c_mel"], out["c_wav"] = mel_c, wav_c
        report["oracle_vs_reference"]["wav_c"] = rel(o_hifigan.decode_to_waveform(vae_sd, mel_c, True), wav_c)

        # ---------------- case D: multi-step consistency sampling with the REFERENCE scheduler objects
        # (easy_inference/consistencytta.py:159-197 verbatim, torch.randn_like replaced by fixed step noises): 4 steps,
        # B = 1, L = 8 -> queries at t = 999, 666, 333, 0
        noise, enc, mask = weights.synthetic_inputs(1, 8, seed=77)
        g = torch.Generator().manual_seed(78)
        step_noises = [torch.randn(1, 8, 256, 16, generator=g) for _ in range(3)]
        sched.set_timesteps(18)
        z_N = noise * sched.init_noise_sigma

        def calc_zhat_0(z_n, t):
            z_n_input = sched.scale_model_input(z_n, t)
            return unet(z_n_input, t, guidance=3.0, encoder_hidden_states=enc, encoder_attention_mask=mask).sample

        zhat_0 = calc_zhat_0(z_N, sched.timesteps[0])
        sched.set_timesteps(4)
        ts_d = [float(t) for t in sched.timesteps[1::2]]
        for i, t in enumerate(sched.timesteps[1::2]):
            zhat_n = sched.add_noise(zhat_0, step_noises[i], t)
            zhat_0 = calc_zhat_0(zhat_n, t)
        out["d_latent"] = zhat_0
        report["case_d_timesteps"] = ts_d
        o_lat_d = o_pipe.generate_latent_multistep(unet_sd, noise, step_noises, enc, mask, 3.0, 4)
        report["oracle_vs_reference"]["latent_d_multistep4"] = rel(o_lat_d, zhat_0)

        # ---------------- case E: the non-EDM variant (inference.py:159-162 without --use_edm): reference DDIMScheduler,
        # models/audio_consistency_model.py:486-507 with use_edm False (stride 1), 2 steps -> queries at t = 935, 0
        ddim = import_reference.DDIM(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                                     beta_schedule="scaled_linear", prediction_type="v_prediction")
        ddim.set_timesteps(18)
        z_N = noise * ddim.init_noise_sigma
        t0_e = ddim.timesteps[0]
        zhat_0 = unet(ddim.scale_model_input(z_N, t0_e), t0_e, guidance=3.0, encoder_hidden_states=enc,
                      encoder_attention_mask=mask).sample
        ddim.set_timesteps(2)
        ts_e = [int(t0_e)] + [int(t) for t in ddim.timesteps[1::1]]
        for i, t in enumerate(ddim.timesteps[1::1]):
            zhat_n = ddim.add_noise(zhat_0, step_noises[i], t)
            zhat_0 = unet(ddim.scale_model_input(zhat_n, t), t, guidance=3.0, encoder_hidden_states=enc,
                          encoder_attention_mask=mask).sample
        out["e_latent"] = zhat_0
        report["case_e_timesteps"] = ts_e
        o_lat_e = o_pipe.generate_latent_multistep_ddim(unet_sd, noise, step_noises, enc, mask, 3.0, 2)
        report["oracle_vs_reference"]["latent_e_ddim2"] = rel(o_lat_e, zhat_0)

    print(json.dumps(report, indent=1))
    for k, v in report["oracle_vs_reference"].items():
        if k != "int16_max_abs_diff":
            assert v < 1e-4, (k, v)
    torch.save({k: v.contiguous() for k, v in out.items()}, os.path.join(GOLD, "reference_outputs.pt"))
    with open(os.path.join(GOLD, "reference_report.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
