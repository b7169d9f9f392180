#!/usr/bin/env python
"""Stock-PyTorch bar on the SAME B200 (SURVEY.md 8d "GPU reference baseline"): the fp32 oracle restatement of the
reference's modules (the reference itself is a Python tree that does not exist on the GPU box) run in eager mode on
cuda:0 — fp32 with TF32 matmuls/convs (torch's default for convs), and under bf16 autocast as `inference.py` can be run —
next to this repository's engine at the same batch.  Lives under tests/ because only tests/, smoke() and bench.py's
CPU leg may execute oracle/.  Collected by pytest (`-m gpu`) at batch 8: the parity leg uses the STRICT fp32 eager run
(TF32 off) as a second, GPU-side oracle — latent, mel AND waveform SNR — and the speed leg asserts the engine beats
stock bf16-autocast eager.  As a script it prints the JSON record kept under profiles/:

    python tests/test_torch_eager_gpu_baseline.py --batch 64 --iters 3
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from consistencytta_b200 import SingleStepEngine, build_random_init_models, weights  # noqa: E402
from oracle import hifigan, pipeline  # noqa: E402


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(batch=8, iters=3, text_len=32):
    a = argparse.Namespace(batch=batch, iters=iters, text_len=text_len)
    dev = torch.device("cuda", 0)
    usd = {k: v.to(dev) for k, v in weights.make_unet_state_dict(0).items()}
    vsd = {k: v.to(dev) for k, v in weights.make_vae_state_dict(1).items()}
    noise, enc, mask = (t.to(dev) for t in weights.synthetic_inputs(a.batch, a.text_len))
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    out = {"batch": a.batch, "text_len": a.text_len, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}

    def eager():
        with torch.no_grad():
            lat, mel, wav = pipeline.generate(usd, vsd, weights.SCALE_FACTOR, noise, enc, mask, 4.0)
            return lat, mel, wav

    def eager_bf16():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return pipeline.generate(usd, vsd, weights.SCALE_FACTOR, noise, enc, mask, 4.0)

    ms = timed(eager, a.iters)
    out["torch_eager_fp32_tf32"] = {"ms_per_step": ms, "clips_per_s": a.batch / ms * 1e3}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref_lat, ref_mel, ref_wav = eager()      # strict fp32: the parity reference
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    ms = timed(eager_bf16, a.iters)
    out["torch_eager_bf16_autocast"] = {"ms_per_step": ms, "clips_per_s": a.batch / ms * 1e3}
    del usd, vsd
    torch.cuda.empty_cache()
    unet, vae = build_random_init_models(dev)
    eng = SingleStepEngine(unet, vae, use_graphs=True)
    res = eng.run(noise, enc, mask, 4.0)
    ms = timed(lambda: eng.run(noise, enc, mask, 4.0), max(a.iters, 5))
    out["consistencytta_b200"] = {"ms_per_step": ms, "clips_per_s": a.batch / ms * 1e3}
    rel = lambda x, y: ((x.double() - y.double()).norm() / y.double().norm()).item()
    d = res["wav"].double() - ref_wav.double()
    snr = float(10 * torch.log10(ref_wav.double().pow(2).sum() / d.pow(2).sum()))
    out["parity_vs_torch_eager_fp32"] = {"latent_rel_l2": rel(res["latent"], ref_lat), "mel_rel_l2": rel(res["mel"], ref_mel),
                                         "wav_snr_db": snr}
    out["speedup_vs_fp32_tf32"] = out["consistencytta_b200"]["clips_per_s"] / out["torch_eager_fp32_tf32"]["clips_per_s"]
    out["speedup_vs_bf16_autocast"] = out["consistencytta_b200"]["clips_per_s"] / out["torch_eager_bf16_autocast"]["clips_per_s"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--text-len", type=int, default=32)
    a = ap.parse_args()
    print(json.dumps(run(a.batch, a.iters, a.text_len)))


try:
    import pytest
except ImportError:  # script use without pytest
    pytest = None

if pytest is not None:
    @pytest.mark.gpu
    def test_engine_matches_and_beats_stock_torch_eager_on_gpu():
        out = run(batch=8, iters=2)
        print(json.dumps(out))
        par = out["parity_vs_torch_eager_fp32"]
        assert par["latent_rel_l2"] <= 1e-2 and par["mel_rel_l2"] <= 1e-2 and par["wav_snr_db"] >= 35.0
        assert out["speedup_vs_bf16_autocast"] > 1.5 and out["speedup_vs_fp32_tf32"] > 2.0
        d = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(d):
            with open(os.path.join(d, "torch_eager_gpu_baseline_b8.json"), "w") as f:
                json.dump(out, f)


if __name__ == "__main__":
    main()
