"""CPU-side tests: oracle against the golden vectors of the reference, schemas, scheduler, and the C-ABI surface."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def test_schema_matches_reference_state_dicts():
    from consistencytta_b200 import weights
    ref = json.load(open(os.path.join(GOLD, "state_dict_schema.json")))
    ours = {k: list(s) for k, (s, _) in weights.unet_schema().items()}
    assert ours == ref["unet"] and len(ours) == 691
    assert sum(int(np.prod(s)) for s in ours.values()) == 559209676
    vae = {k: list(s) for k, (s, _) in weights.vae_schema().items()}
    ref_dec = {k: v for k, v in ref["vae"].items() if not k.startswith(("encoder.", "quant_conv."))}
    assert vae == ref_dec
    assert len([k for k in vae if k.startswith("vocoder.")]) == 194 and len(ref["vae"]) == 398


def test_synthetic_weights_deterministic():
    from consistencytta_b200 import weights
    a = weights.make_state_dict(weights.vocoder_schema(), 1)
    b = weights.make_state_dict(weights.vocoder_schema(), 1)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert abs(a["vocoder.resblocks.0.convs1.0.weight"].std().item() - 0.01) < 1e-3


def test_scheduler_matches_reference_constants():
    from consistencytta_b200 import HeunDiscreteScheduler
    from oracle import pipeline as op
    s = HeunDiscreteScheduler.from_pretrained("stabilityai/stable-diffusion-2-1", subfolder="scheduler")
    s.set_timesteps(18)
    rep = json.load(open(os.path.join(GOLD, "reference_report.json")))
    assert float(s.timesteps[0]) == rep["t0"] == 999.0
    assert abs(float(s.init_noise_sigma) - rep["sigma_max"]) < 1e-6
    assert len(s.timesteps) == 35 and len(s.sigmas) == 36
    t0, sig = op.heun_first_step()
    assert t0 == 999.0 and abs(sig - rep["sigma_max"]) < 1e-6
    x = torch.ones(2, 8, 4, 4)
    assert torch.allclose(s.scale_model_input(x, s.timesteps[0]), x / (sig ** 2 + 1) ** 0.5)


def test_oracle_vocoder_matches_reference_golden():
    """Cheap oracle leg (the UNet / VAE legs take seconds each and are covered by test_oracle_full_chain)."""
    from consistencytta_b200 import weights
    from oracle import hifigan
    gold = torch.load(os.path.join(GOLD, "reference_outputs.pt"))
    sd = weights.make_state_dict(weights.vocoder_schema(), 1)
    wav = hifigan.decode_to_waveform(sd, gold["c_mel"], return_float=True)
    assert rel(wav, gold["c_wav"]) < 1e-5


def test_oracle_full_chain_matches_reference_golden():
    from consistencytta_b200 import weights
    from oracle import hifigan, pipeline
    gold = torch.load(os.path.join(GOLD, "reference_outputs.pt"))
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    noise, enc, mask = weights.synthetic_inputs(1, 32)
    with torch.no_grad():
        lat, mel, wav = pipeline.generate(weights.make_unet_state_dict(0), weights.make_vae_state_dict(1),
                                          weights.SCALE_FACTOR, noise, enc, mask, 4.0)
    assert lat.shape == (1, 8, 256, 16) and mel.shape == (1, 1, 1024, 64) and wav.shape == (1, 163872)
    assert rel(lat, gold["a_latent"]) < 1e-4 and rel(mel, gold["a_mel"]) < 1e-4 and rel(wav, gold["a_wav"]) < 1e-4
    i16 = hifigan.to_int16(wav)
    assert np.abs(i16.astype(np.int32) - gold["a_wav_i16"].numpy().astype(np.int32)).max() <= 1


def test_numpy_int16_semantics():
    """hifigan/utilities.py:86 relies on numpy's float->int16 cast: truncation toward zero, wrap on overflow."""
    x = np.array([0.9, -1.7, 40000.7, -40000.7, 32767.9], dtype=np.float32)
    assert x.astype("int16").tolist() == [0, -1, -25536, 25536, 32767]


def test_c_abi_exports_every_declared_symbol():
    from consistencytta_b200 import _lib, build
    path = build.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "ctta.h")).read()
    declared = set(re.findall(r"\b(ctta_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib.ctta_version.restype = ctypes.c_int
    assert lib.ctta_version() >= 100
    # struct mirror must match the C layout
    assert ctypes.sizeof(_lib.GemmDesc) % 8 == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "consistencytta_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_modules_refuse_cpu():
    from consistencytta_b200 import UNet2DConditionGuidedModel
    u = UNet2DConditionGuidedModel()
    assert len(u.state_dict()) == 691
    with pytest.raises(RuntimeError):
        u(torch.zeros(1, 8, 256, 16), 999.0, guidance=4.0, encoder_hidden_states=torch.zeros(1, 4, 1024))


def test_oracle_heun_schedule_matches_scheduler():
    """oracle.heun_schedule (used by the multi-step oracle) against the product's HeunDiscreteScheduler mirror."""
    from consistencytta_b200 import HeunDiscreteScheduler
    from oracle import pipeline as op
    s = HeunDiscreteScheduler.from_pretrained()
    for n in (2, 4, 18):
        s.set_timesteps(n)
        ts, sg = op.heun_schedule(n)
        assert np.allclose(s.timesteps.numpy(), ts) and np.allclose(s.sigmas.numpy(), sg)
        for t in s.timesteps[1::2]:
            idx = int(np.argmax(ts == float(t)))
            assert abs(float(s.sigma_for_timestep(t)) - float(sg[idx])) < 1e-7


def test_audiolcm_load_pretrained_key_remap():
    """models/audio_consistency_model.py:160-204: consistency_ema_* -> student_target_* (and student_ema_* unless a
    consistency_slow_ema_* key exists); trainable student, teacher, VAE and loss tensors are not used at inference."""
    from consistencytta_b200 import AudioLCM
    m = AudioLCM()
    w = m.student_target_unet.conv_in.weight
    b = m.student_target_unet.conv_in.bias
    ckpt = {
        "consistency_ema_unet.conv_in.weight": torch.full_like(w, 1.0),
        "consistency_slow_ema_unet.conv_in.weight": torch.full_like(w, 2.0),
        "consistency_ema_unet.conv_in.bias": torch.full_like(b, 3.0),
        "consistency_unet.conv_in.weight": torch.full_like(w, 9.0),
        "diffusion_unet.conv_in.weight": torch.full_like(w, 9.0),
        "vae.decoder.conv_in.weight": torch.zeros(3),
        "loss.window": torch.zeros(3),
    }
    info = m.load_pretrained(ckpt)
    assert not info.unexpected_keys
    assert (m.student_target_unet.conv_in.weight == 1.0).all()
    assert (m.student_ema_unet.conv_in.weight == 2.0).all()          # slow EMA wins
    assert (m.student_target_unet.conv_in.bias == 3.0).all()
    assert (m.student_ema_unet.conv_in.bias == 3.0).all()            # no slow EMA -> copy of the EMA weights
    m.check_eval_mode()


def test_bench_reference_arm_line_and_nonzero_rank_exit():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores): rank 0 prints ONE JSON line with the
    contract's keys; under torchrun every other rank exits 0 without work."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--text-len", "8"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_micro_batch_row_slicing_keeps_cfg_layout():
    """Host logic of SingleStepEngine._run_micro_batches: per-request tensors are sliced per micro-batch; with post-CFG
    the text tensors are [unconditional ; conditional] (2b rows) and every slice must keep that layout."""
    from consistencytta_b200.pipeline import slice_request_rows
    b = 5
    enc = torch.arange(2 * b).float().view(2 * b, 1, 1).expand(2 * b, 3, 4)
    s = slice_request_rows(enc, 2, 4, b, True)
    assert s.shape == (4, 3, 4) and s[:, 0, 0].tolist() == [2.0, 3.0, 7.0, 8.0]
    assert slice_request_rows(enc[:b], 4, 5, b, False)[:, 0, 0].tolist() == [4.0]
    g = torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0])
    assert slice_request_rows(g, 0, 2, b, True).tolist() == [1.0, 2.0]          # per-clip guidance: b rows even with CFG
    assert slice_request_rows(torch.tensor([3.0]), 2, 4, b, True).tolist() == [3.0]   # broadcast value
    assert slice_request_rows(4.0, 0, 2, b, False) == 4.0 and slice_request_rows(None, 0, 2, b, False) is None
    # the slices of all micro-batches cover every row exactly once, in order
    got = torch.cat([slice_request_rows(enc, lo, min(lo + 2, b), b, True)[: min(lo + 2, b) - lo] for lo in range(0, b, 2)])
    assert got[:, 0, 0].tolist() == [0.0, 1.0, 2.0, 3.0, 4.0]
