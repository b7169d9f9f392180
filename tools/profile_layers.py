#!/usr/bin/env python
"""Per-launch timing of the hot path in eager mode: CUDA events around every libctta call, aggregated by kernel
family and (for ctta_gemm) by problem shape, with achieved TFLOP/s against padded and true FLOPs.

    python tools/profile_layers.py --batch 64 --out gpurun_out/layers_b64.json
"""
import argparse
import json
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from consistencytta_b200 import SingleStepEngine, build_random_init_models, ops, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--text-len", type=int, default=32)
    ap.add_argument("--out", default="gpurun_out/layers.json")
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    unet, vae = build_random_init_models(dev)
    eng = SingleStepEngine(unet, vae, use_graphs=False)
    noise, enc, mask = weights.synthetic_inputs(args.batch, args.text_len)
    eng.run(noise, enc, mask, 4.0)
    torch.cuda.synchronize()

    records = []
    stage = {"name": "unet"}

    def wrap(name, fn, describe):
        def inner(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            records.append((name, describe(*a, **k), s, e, stage["name"]))
            return r
        return inner

    def d_gemm(a, pw, mode=0, n_img=1, h=1, w=None, rows_per_img=None, **k):
        rows = n_img * (rows_per_img if rows_per_img is not None else h * (w or 1))
        flops = 2.0 * rows * pw.n * pw.ntaps * pw.c
        return {"mode": mode, "rows": rows, "n": pw.n, "c": pw.c, "taps": pw.ntaps, "flops": flops,
                "out": str(k.get("out").dtype) if k.get("out") is not None else None, "out2": k.get("out2") is not None,
                "res": k.get("residual") is not None}

    def d_generic(*a, **k):
        t = a[0][0] if isinstance(a[0], (list, tuple)) else a[0]
        shp = list(t.shape)
        if len(a) > 1 and torch.is_tensor(a[1]):
            shp = [shp, list(a[1].shape)]
        return {"numel": t.numel(), "shape": shp}

    orig = {}
    for name in ("gemm", "groupnorm_stats", "groupnorm_apply", "layernorm", "attention", "softmax_rows", "im2col_s2",
                 "small_linear", "wave_to_int16", "nchw_to_nhwc", "nhwc_to_nchw", "time_features", "tap_sum", "mrf_combine",
                 "resblock_pair"):
        orig[name] = getattr(ops, name)
        setattr(ops, name, wrap(name, orig[name], d_gemm if name == "gemm" else d_generic))
    # stage markers
    vdec, voc = vae.decode_nhwc, vae.vocoder.forward_btc

    def vdec_m(*a, **k):
        stage["name"] = "vae"
        return vdec(*a, **k)

    def voc_m(*a, **k):
        stage["name"] = "vocoder"
        return voc(*a, **k)
    vae.decode_nhwc, vae.vocoder.forward_btc = vdec_m, voc_m
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    eng.run(noise, enc, mask, 4.0)
    t1.record()
    torch.cuda.synchronize()
    total = t0.elapsed_time(t1)

    by_kernel = defaultdict(lambda: [0, 0.0])
    by_stage = defaultdict(float)
    shapes = defaultdict(lambda: {"count": 0, "ms": 0.0, "flops": 0.0})
    other = defaultdict(lambda: {"count": 0, "ms": 0.0})
    for name, desc, s, e, st in records:
        ms = s.elapsed_time(e)
        by_kernel[name][0] += 1
        by_kernel[name][1] += ms
        by_stage[st + ":" + name] += ms
        if name == "gemm":
            key = "%s mode%d rows=%d n=%d c=%d taps=%d out=%s out2=%s res=%s" % (
                st, desc["mode"], desc["rows"], desc["n"], desc["c"], desc["taps"], desc["out"], desc["out2"], desc["res"])
            shapes[key]["count"] += 1
            shapes[key]["ms"] += ms
            shapes[key]["flops"] += desc["flops"]
        else:
            key = "%s %s %s" % (st, name, desc["shape"])
            other[key]["count"] += 1
            other[key]["ms"] += ms
    rows = sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])
    print("total eager step %.2f ms (batch %d)" % (total, args.batch))
    for k, (n, ms) in sorted(by_kernel.items(), key=lambda kv: -kv[1][1]):
        print("%-18s %5d launches %9.3f ms %5.1f%%" % (k, n, ms, 100 * ms / total))
    for k, ms in sorted(by_stage.items(), key=lambda kv: -kv[1])[:12]:
        print("  %-28s %9.3f ms" % (k, ms))
    print("top gemm shapes (padded-dim TFLOP/s):")
    for k, v in rows[:args.top]:
        print("%8.3f ms x%-3d %7.1f TF/s  %s" % (v["ms"], v["count"], v["flops"] / v["ms"] / 1e9, k))
    print("other kernels by first-argument shape:")
    orows = sorted(other.items(), key=lambda kv: -kv[1]["ms"])
    for k, v in orows[:args.top]:
        print("%8.3f ms x%-3d  %s" % (v["ms"], v["count"], k))
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"batch": args.batch, "total_ms": total, "by_kernel": {k: v for k, v in by_kernel.items()},
               "by_stage": by_stage, "gemm_shapes": rows, "other_shapes": orows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
