"""CPU-side tests of the caller-facing pieces either side of the hot path: multi-step / DDIM
oracles against reference goldens, scheduler mirrors, prefix-mask refusal, prompt cache, checkpoint formats, EMA
decoder plumbing, packed-operand versioning, WAV output."""
import io
import json
import os
import struct
import wave

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def _step_noises(n):
    g = torch.Generator().manual_seed(78)
    return [torch.randn(1, 8, 256, 16, generator=g) for _ in range(n)]


def test_oracle_multistep_matches_reference_golden():
    """Case D of oracle/make_golden.py: easy_inference/consistencytta.py:159-197 run with the REFERENCE scheduler and
    UNet objects, 4 steps (queries at t = 999, 666, 333, 0), fixed re-noising draws."""
    from consistencytta_b200 import weights
    from oracle import pipeline as op
    gold = torch.load(os.path.join(GOLD, "reference_outputs.pt"))
    rep = json.load(open(os.path.join(GOLD, "reference_report.json")))
    assert rep["case_d_timesteps"] == [666.0, 333.0, 0.0]
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    noise, enc, mask = weights.synthetic_inputs(1, 8, seed=77)
    with torch.no_grad():
        lat = op.generate_latent_multistep(weights.make_unet_state_dict(0), noise, _step_noises(3), enc, mask, 3.0, 4)
    assert rel(lat, gold["d_latent"]) < 1e-4


def test_oracle_ddim_matches_reference_golden_and_scheduler_mirror():
    """Case E: the non-EDM variant (reference DDIMScheduler, stride 1, queries at t = 935 then 0)."""
    from consistencytta_b200 import DDIMScheduler, weights
    from oracle import pipeline as op
    gold = torch.load(os.path.join(GOLD, "reference_outputs.pt"))
    rep = json.load(open(os.path.join(GOLD, "reference_report.json")))
    assert rep["case_e_timesteps"] == [935, 0]
    s = DDIMScheduler.from_pretrained("stabilityai/stable-diffusion-2-1", subfolder="scheduler")
    s.set_timesteps(18)
    assert int(s.timesteps[0]) == 935 and s.init_noise_sigma == 1.0 and s.input_scale(s.timesteps[0]) == 1.0
    assert s.timesteps.tolist() == op.ddim_timesteps(18).tolist()
    s.set_timesteps(2)
    assert s.timesteps.tolist() == [500, 0]
    x, n = torch.randn(2, 8, 4, 4), torch.randn(2, 8, 4, 4)
    ac = op.ddim_alphas_cumprod()
    assert torch.allclose(s.add_noise(x, n, s.timesteps[1]), ac[0] ** 0.5 * x + (1 - ac[0]) ** 0.5 * n)
    assert s.scale_model_input(x, 5) is x
    with pytest.raises(ValueError):
        s.set_timesteps(1001)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    noise, enc, mask = weights.synthetic_inputs(1, 8, seed=77)
    with torch.no_grad():
        lat = op.generate_latent_multistep_ddim(weights.make_unet_state_dict(0), noise, _step_noises(3), enc, mask, 3.0, 2)
    assert rel(lat, gold["e_latent"]) < 1e-4


def test_heun_index_for_timestep_takes_last_match():
    """scheduling_heun_discrete.py:137-149: (mask * arange).argmax() -> LAST matching position; input_scale is what
    scale_model_input multiplies by."""
    from consistencytta_b200 import HeunDiscreteScheduler
    from oracle import pipeline as op
    s = HeunDiscreteScheduler.from_pretrained()
    s.set_timesteps(4)
    assert s.timesteps.tolist() == [999.0, 666.0, 666.0, 333.0, 333.0, 0.0, 0.0]
    assert s.index_for_timestep(torch.tensor(666.0)).tolist() == [2]
    assert s.index_for_timestep(torch.tensor(999.0)).tolist() == [0]
    assert s.index_for_timestep(torch.tensor(0.0)).tolist() == [6]
    ts, sg = op.heun_schedule(4)
    for t in (999.0, 666.0, 333.0, 0.0):
        i = op.heun_index_for_timestep(ts, t)
        assert i == int(s.index_for_timestep(torch.tensor(t))[0])
        assert abs(s.input_scale(t) - 1.0 / (float(sg[i]) ** 2 + 1) ** 0.5) < 1e-7
    x = torch.ones(1, 8, 2, 2)
    assert torch.allclose(s.scale_model_input(x, 333.0), x * s.input_scale(333.0))


def test_prefix_mask_lengths_refuses_interior_holes():
    from consistencytta_b200.unet import prefix_mask_lengths
    m = torch.tensor([[1, 1, 1, 0], [1, 0, 0, 0], [1, 1, 1, 1], [0, 0, 0, 0]], dtype=torch.bool)
    assert prefix_mask_lengths(m).tolist() == [3, 1, 4, 0]
    assert prefix_mask_lengths(m.long()).dtype == torch.int32
    with pytest.raises(ValueError):
        prefix_mask_lengths(torch.tensor([[1, 0, 1, 0]], dtype=torch.bool))
    with pytest.raises(ValueError):
        prefix_mask_lengths(torch.tensor([[0, 1, 1, 1]], dtype=torch.bool))
    with pytest.raises(ValueError):
        prefix_mask_lengths(torch.ones(4, dtype=torch.bool))


class _Tok:
    model_max_length = 16

    def __init__(self):
        self.calls = []

    def __call__(self, prompts, max_length=None, padding=True, truncation=True, return_tensors="pt"):
        self.calls.append(list(prompts))
        ids = [[(sum(map(ord, w)) % 1000) + 2 for w in p.split()][: max_length - 1] + [1] for p in prompts]
        n = max_length if padding == "max_length" else max(len(i) for i in ids)
        batch = type("Batch", (), {})()
        batch.input_ids = torch.tensor([i + [0] * (n - len(i)) for i in ids])
        batch.attention_mask = torch.tensor([[1] * len(i) + [0] * (n - len(i)) for i in ids])
        return batch


class _Enc(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.emb = torch.nn.Embedding(1002, 1024)

    @property
    def device(self):
        return self.emb.weight.device

    def forward(self, input_ids=None, attention_mask=None):
        return (self.emb(input_ids),)


def test_prompt_cache_encodes_each_prompt_once_and_matches_uncached():
    """N1: the text encoder sees a prompt once; batches are re-assembled from the cached rows (consistencytta.py:84-132)."""
    from consistencytta_b200 import ConsistencyTTA
    torch.manual_seed(0)
    enc = _Enc().eval().requires_grad_(False)
    cached = ConsistencyTTA(text_encoder=enc, tokenizer=_Tok())
    plain = ConsistencyTTA(text_encoder=enc, tokenizer=_Tok(), text_cache_size=0)
    prompts = ["a dog barks twice", "rain", "a dog barks twice"]
    e1, m1 = cached.encode_text(prompts)
    e0, m0 = plain.encode_text(prompts)
    assert torch.equal(m1, m0) and m1.dtype == torch.bool
    assert torch.equal(e1 * m1[..., None], e0 * m0[..., None])      # rows of real tokens are identical; padding is masked
    assert cached.tokenizer.calls == [["a dog barks twice", "rain"]] and cached.text_encoder_calls == 1
    cached.encode_text(["rain", "a dog barks twice"])
    assert cached.text_encoder_calls == 1                            # nothing new to encode
    cf, mcf, c, mc = cached.encode_text_classifier_free(["rain", "wind in trees"], 2)
    cf0, mcf0, c0, mc0 = plain.encode_text_classifier_free(["rain", "wind in trees"], 2)
    assert cf.shape == cf0.shape == (8, 4, 1024) and torch.equal(mcf, mcf0) and torch.equal(mc, mc0)
    assert torch.equal(cf * mcf[..., None], cf0 * mcf0[..., None])
    assert mcf[:4].sum(1).tolist() == [1, 1, 1, 1]                   # "" is one EOS token, padded to max_length
    assert cached.text_encoder_calls == 3                            # "wind in trees" and "" were new
    small = ConsistencyTTA(text_encoder=enc, tokenizer=_Tok(), text_cache_size=2)
    small.encode_text(["a", "b", "c"])
    assert len(small._text_cache) == 2


def test_vae_checkpoint_formats(tmp_path):
    """N3: {"state_dict", "scale_factor"} files (consistencytta.py:32-44), AudioLDM checkpoints with the
    `first_stage_model.` prefix and a tensor scale_factor (tools/build_pretrained.py:9-22), EMA decoder keys."""
    from consistencytta_b200 import AutoencoderKL, weights
    sd = weights.make_vae_state_dict(1)
    ref_schema = json.load(open(os.path.join(GOLD, "state_dict_schema.json")))["vae"]
    full = dict(sd)
    for k, shape in ref_schema.items():          # encode-side tensors of the reference's 398-key file are accepted and dropped
        full.setdefault(k, torch.zeros(shape))
    assert len(full) == 398
    p1 = tmp_path / "vae_state_dict.pt"
    torch.save({"state_dict": full, "scale_factor": 0.9227914214134216}, p1)
    vae = AutoencoderKL.from_vae_state_dict_file(str(p1))
    assert abs(vae.scale_factor - 0.9227914214134216) < 1e-12 and not vae.training
    assert torch.equal(vae.decoder.conv_in.weight, sd["decoder.conv_in.weight"])
    assert torch.equal(vae.vocoder.conv_pre.weight, sd["vocoder.conv_pre.weight"])
    ckpt = {"state_dict": {"first_stage_model." + k: v for k, v in full.items()}}
    ckpt["state_dict"]["scale_factor"] = torch.tensor(0.5)
    ckpt["state_dict"]["model.diffusion_model.input_blocks.0.0.weight"] = torch.zeros(3)
    ckpt["state_dict"]["cond_stage_model.model.logit_scale_a"] = torch.zeros(())
    p2 = tmp_path / "audioldm-s-full.ckpt"
    torch.save(ckpt, p2)
    vae2 = AutoencoderKL.from_audioldm_checkpoint(str(p2))
    assert vae2.scale_factor == 0.5
    assert all(torch.equal(a, b) for a, b in zip(vae2.state_dict().values(), vae.state_dict().values()))
    # EMA modules: created on demand, distinct tensors, part of the state_dict, restored by load_state_dict
    assert vae.ema_decoder is None
    v0 = vae.pack_version
    vae.enable_ema_modules()
    assert vae.pack_version != v0
    assert vae.ema_decoder is not vae.decoder and torch.equal(vae.ema_decoder.conv_in.weight, vae.decoder.conv_in.weight)
    with torch.no_grad():
        vae.ema_decoder.conv_in.weight.add_(1.0)
        vae.ema_post_quant_conv.bias.add_(2.0)
    saved = vae.state_dict()
    assert "ema_decoder.conv_in.weight" in saved and "ema_post_quant_conv.bias" in saved
    vae3 = AutoencoderKL(scale_factor=1.0)
    vae3.load_state_dict(saved)
    assert vae3.ema_decoder is not None
    assert torch.equal(vae3.ema_decoder.conv_in.weight, vae3.decoder.conv_in.weight + 1.0)
    assert torch.equal(vae3.ema_post_quant_conv.bias, vae3.post_quant_conv.bias + 2.0)


def test_audiolcm_ftvae_load_pretrained():
    """models/audio_consistency_model_ftvae.py:69-91: vae.decoder / vae.post_quant_conv / ema_vae_decoder / ema_vae_pqconv
    come out of the training checkpoint; `loss.`-prefixed duplicates only fill gaps."""
    from consistencytta_b200 import AudioLCM_FTVAE, AutoencoderKL
    vae = AutoencoderKL(scale_factor=1.0)
    m = AudioLCM_FTVAE(vae=vae)
    dec_sd = vae.decoder.state_dict()
    ckpt = {}
    for k, v in dec_sd.items():
        ckpt["vae.decoder." + k] = torch.full_like(v, 1.0)
        ckpt["ema_vae_decoder." + k] = torch.full_like(v, 2.0)
        ckpt["loss.vae.decoder." + k] = torch.full_like(v, 7.0)        # duplicate: must not win
    for k, v in vae.post_quant_conv.state_dict().items():
        ckpt["loss.vae.post_quant_conv." + k] = torch.full_like(v, 3.0)  # only copy: fills the gap
        ckpt["ema_vae_pqconv." + k] = torch.full_like(v, 4.0)
    w = m.student_target_unet.conv_in.weight
    ckpt["consistency_ema_unet.conv_in.weight"] = torch.full_like(w, 5.0)
    v0 = vae.pack_version
    m.load_pretrained(ckpt)
    assert (vae.decoder.conv_in.weight == 1.0).all() and (vae.ema_decoder.conv_in.weight == 2.0).all()
    assert (vae.post_quant_conv.weight == 3.0).all() and (vae.ema_post_quant_conv.weight == 4.0).all()
    assert (m.student_target_unet.conv_in.weight == 5.0).all()
    assert vae.pack_version != v0
    assert m.vae.ema_decoder is m.ema_vae_decoder


def test_pack_version_changes_on_any_reload_path():
    """captured graphs are keyed on pack_version; it must move for load_state_dict on the module itself, on a
    PARENT module (nn.Module.load_state_dict does not call the children's), and for .to() / .half()."""
    from consistencytta_b200 import AutoencoderKL, ConsistencyTTA, UNet2DConditionGuidedModel
    unet, vae = UNet2DConditionGuidedModel(), AutoencoderKL()
    tta = ConsistencyTTA(unet, vae)
    seen = {unet.pack_version}
    unet.load_state_dict(unet.state_dict())
    assert unet.pack_version not in seen
    seen.add(unet.pack_version)
    v_vae = vae.pack_version
    tta.load_state_dict(tta.state_dict())
    assert unet.pack_version not in seen and vae.pack_version != v_vae
    seen.add(unet.pack_version)
    v_vae = vae.pack_version
    tta.float()
    assert unet.pack_version not in seen and vae.pack_version != v_vae
    v_dec = vae.decoder.pack_version
    vae.decoder.load_state_dict(vae.decoder.state_dict())
    assert vae.decoder.pack_version != v_dec and vae.pack_version != v_vae


def test_wav_writer_is_canonical_pcm16(tmp_path):
    """inference.py:221-222 writes int16 clips with soundfile -> canonical 44-byte RIFF header + little-endian samples."""
    from consistencytta_b200 import save_clips, wav_bytes_pcm16, write_wav_pcm16
    rng = np.random.default_rng(0)
    x = rng.integers(-32768, 32767, size=1600, dtype=np.int16)
    raw = wav_bytes_pcm16(x, 16000)
    expect = (b"RIFF" + struct.pack("<I", 36 + 3200) + b"WAVEfmt " + struct.pack("<I", 16) + struct.pack("<H", 1) +
              struct.pack("<H", 1) + struct.pack("<I", 16000) + struct.pack("<I", 32000) + struct.pack("<H", 2) +
              struct.pack("<H", 16) + b"data" + struct.pack("<I", 3200) + x.astype("<i2").tobytes())
    assert raw == expect and len(raw) == 44 + 3200
    ref = io.BytesIO()
    with wave.open(ref, "wb") as w:            # the standard library writes the same canonical header
        w.setnchannels(1)
        w.setsampwidth(2)
        w.setframerate(16000)
        w.writeframes(x.astype("<i2").tobytes())
    assert raw == ref.getvalue()
    p = write_wav_pcm16(str(tmp_path / "a.wav"), x)
    with wave.open(p, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, 1600)
        assert np.array_equal(np.frombuffer(w.readframes(1600), dtype="<i2"), x)
    with pytest.raises(TypeError):
        wav_bytes_pcm16(x.astype(np.float32))
    clips = rng.integers(-100, 100, size=(3, 163872), dtype=np.int16)
    paths = save_clips(clips, str(tmp_path / "out"), sr=16000)
    assert [os.path.basename(q) for q in paths] == ["output_0.wav", "output_1.wav", "output_2.wav"]
    with wave.open(paths[2], "rb") as w:       # truncated to 10 s (inference.py:208)
        assert w.getnframes() == 160000
        assert np.array_equal(np.frombuffer(w.readframes(160000), dtype="<i2"), clips[2, :160000])


def test_engine_text_bucket_and_lru_bookkeeping():
    """Host logic of the graph cache (no CUDA needed): text lengths round up to the bucket, least-recently-used out."""
    from consistencytta_b200 import SingleStepEngine

    class _U:
        pack_version = 1
        device = torch.device("cpu")

    eng = SingleStepEngine(_U(), None, text_bucket=16, max_buckets=2)
    assert [eng.text_bucket_len(n) for n in (1, 16, 17, 32, 33)] == [16, 16, 32, 32, 48]
    eng._make_io = lambda b, n, cf, dev: {"shape": (b, n, cf)}
    a = eng._bucket(4, 20, False, "cuda:0")
    assert eng._bucket(4, 30, False, "cuda:0") is a and a["io"]["shape"] == (4, 32, False)
    b = eng._bucket(4, 33, False, "cuda:0")
    assert eng._bucket(4, 20, False, "cuda:0") is a            # refreshes a
    c = eng._bucket(4, 20, True, "cuda:0")                     # evicts b (least recently used)
    assert len(eng._buckets) == 2 and (4, 48, False, "cuda:0") not in eng._buckets and c is not a
    with pytest.raises(RuntimeError):
        eng.run(torch.zeros(1, 8, 256, 16), torch.zeros(1, 4, 1024), None, 3.0)   # CPU models: no fallback


def test_scheduler_input_scale_works_for_foreign_scheduler_objects():
    """AudioLCM.inference(prompt, inference_scheduler, ...) may be handed the reference's own scheduler object, which has
    no `input_scale`: the factor is then read off `scale_model_input` with a probe tensor."""
    from consistencytta_b200 import DDIMScheduler, HeunDiscreteScheduler
    from consistencytta_b200.pipeline import scheduler_input_scale

    class Foreign:                       # diffusers surface only
        def __init__(self, inner):
            self._s = inner

        def scale_model_input(self, sample, timestep):
            return self._s.scale_model_input(sample, timestep)

    h = HeunDiscreteScheduler.from_pretrained()
    h.set_timesteps(18)
    for t in (h.timesteps[0], h.timesteps[5]):
        assert abs(scheduler_input_scale(Foreign(h), t) - h.input_scale(t)) < 1e-7
        assert abs(scheduler_input_scale(h, t) - h.input_scale(t)) == 0
    d = DDIMScheduler.from_pretrained()
    d.set_timesteps(18)
    assert scheduler_input_scale(Foreign(d), d.timesteps[0]) == 1.0


def test_fuse_pairs_switch_reads_zero_as_off(monkeypatch):
    """CTTA_FUSE_PAIRS is an A/B switch of the vocoder (fused ResBlock-pair kernel, default on): "0" and "" mean off."""
    from consistencytta_b200.vae import Generator
    for val, want in ((None, True), ("1", True), ("0", False), ("", False)):
        if val is None:
            monkeypatch.delenv("CTTA_FUSE_PAIRS", raising=False)
        else:
            monkeypatch.setenv("CTTA_FUSE_PAIRS", val)
        assert Generator().fuse_pairs is want
