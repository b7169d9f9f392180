#!/usr/bin/env python
"""bench.py — ConsistencyTTA single-step hot path on B200: 10-s clips/s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--text-len L] [--impl ours|reference]

One "step" = one pass of UNet -> VAE decode -> HiFi-GAN -> int16 over a batch of B synthetic prompts per GPU
(BASELINE.json configs[2]: full pipeline, batch 64, 160 000-sample 16 kHz output on 1 B200; random-init weights of
the reference architecture, synthetic N(0,1) latents and text embeddings — no checkpoints exist offline).
Prints ONE JSON line on rank 0 (contract in the task statement).  `--impl reference` times the CPU oracle port
(oracle/, a line-by-line restatement of the reference validated against it) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Algorithmic work per 10-s clip at L = 32 text tokens, true (un-padded) dims, from SURVEY.md 8(d):
GFLOP_PER_CLIP = {"unet": 530.6, "vae": 670.5, "vocoder": 1027.0}
# ... of which executed by the tcgen05 implicit-GEMM kernel (everything but the UNet's fused flash attention):
#   UNet conv 267.9 + linear 163.9, VAE conv 636.1 + AttnBlock (QK^T, PV on the GEMM) 34.4, HiFi-GAN 1027.0
GEMM_GFLOP_PER_CLIP = 267.9 + 163.9 + 636.1 + 34.4 + 1027.0
TOTAL_GFLOP_PER_CLIP = sum(GFLOP_PER_CLIP.values())
CLIP_SAMPLES = 160000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="prompts per GPU per step")
    ap.add_argument("--text-len", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unet-only", action="store_true", help="report the UNet forward ms at --batch (configs[1])")
    ap.add_argument("--no-extras", action="store_true", help="skip the guidance-sweep / small-batch / UNet-B32 loops")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            try:
                pw.append(float(parts[3]))
            except ValueError:
                pass
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_clips_per_s(text_len, steps, warmup, threads=None, batch=1, inputs=None, guidance=4.0):
    """The reference's CPU path restated (oracle/), fp32, all host threads, `batch` clips per step.
    Returns (clips/s, s/step, threads, outputs of the last step)."""
    import torch
    from consistencytta_b200 import weights
    from oracle import hifigan, pipeline
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    usd, vsd = weights.make_unet_state_dict(0), weights.make_vae_state_dict(1)
    noise, enc, mask = inputs if inputs is not None else weights.synthetic_inputs(batch, text_len)
    times, last = [], None
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            lat, mel, wav = pipeline.generate(usd, vsd, weights.SCALE_FACTOR, noise, enc, mask, guidance)
            hifigan.to_int16(wav)
            dt = time.perf_counter() - t0
            last = (lat, mel, wav)
            if i >= warmup:
                times.append(dt)
    sec = sum(times) / len(times)
    return noise.shape[0] / sec, sec, threads, last


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, sec, threads, _ = cpu_reference_clips_per_s(args.text_len, args.steps, args.warmup)
    sample = "1 clip (batch 1, L=%d) per step on the host CPU, fp32, torch threads=%d" % (args.text_len, threads)
    line = {
        "impl": "reference", "metric": "10-s audio clips/sec end-to-end", "value": val, "unit": "clips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "full pipeline UNet+VAE+HiFi-GAN, CPU oracle port of the reference, batch 1",
                   "text_len": args.text_len},
        "cpu_baseline": {"value": val, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def snr_db(x, ref):
    x, ref = x.double().cpu(), ref.double().cpu()
    return float(10 * (ref.pow(2).sum() / (x - ref).pow(2).sum()).log10())


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from consistencytta_b200 import SingleStepEngine, build_random_init_models, ops, weights
    from consistencytta_b200.distributed import WaveformGatherer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, L = args.batch, args.text_len
    W = max(args.warmup, 3)
    stages = "unet" if args.unet_only else "all"
    unet, vae = build_random_init_models(dev)
    eng = SingleStepEngine(unet, vae, use_graphs=True, max_buckets=8)
    # prompts shard by batch across ranks: rank r owns rows [r*B, (r+1)*B) of the global synthetic batch
    noise, enc, mask = weights.synthetic_inputs(B, L, seed=1234 + rank)
    noise_h, enc_h = noise.pin_memory(), enc.pin_memory()
    out_h = torch.empty(B * (world if rank == 0 else 1), CLIP_SAMPLES, dtype=torch.int16).pin_memory()
    lat_h = torch.empty(B, 8, 256, 16, dtype=torch.float32).pin_memory()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def time_replays(graph, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(steps):
            graph.replay()
        e1.record()
        sync_all()
        return e0.elapsed_time(e1)

    def capture(b, n_text, guidance, post=1.0, stg="all", seed=99):
        """Warm-up + capture of one more bucket; returns (graph, launches per replay)."""
        n_, e_, m_ = weights.synthetic_inputs(b, n_text, seed=seed + rank)
        if post > 1.0:
            unc = torch.randn(b, n_text, 1024, generator=torch.Generator().manual_seed(seed + 7))
            e_, m_ = torch.cat([unc, e_]), torch.cat([torch.ones_like(m_), m_])
        eng.run(n_, e_, m_, guidance, post, stages=stg)
        var = eng.last[1]
        for _ in range(W):
            var["graph"].replay()
        return var["graph"], var["launches"]

    # ---- warm-up: first call packs nothing (host-packed weights), captures the CUDA graph, then W replays
    eng.run(noise, enc, mask, 4.0, stages=stages)
    ent, var = eng.last
    graph, io = var["graph"], ent["io"]
    for _ in range(W):
        graph.replay()
    sync_all()

    # ---- timed region 1 (value): K graph replays, inputs already resident in HBM
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else local)
    if rank == 0:
        sampler.start()
    ms = time_replays(graph, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    # the buffers this region produced (checked against the oracle below)
    timed_out = {"latent": io["latent"].clone()}
    if stages == "all":
        timed_out["mel"] = var["refs"]["mel"].view(B, 1, 1024, 64).clone()
        timed_out["wav"] = var["refs"]["wav"].clone()

    # ---- timed region 2 (e2e): public engine call with HOST buffers: pinned H2D, compute, results into host memory.
    # N > 1: every rank's int16 clips are gathered to rank 0 (NCCL, side stream, overlapped with the next step) and
    # rank 0 copies the whole job's clips to pinned host memory, where the files would be written from.
    gat = WaveformGatherer([B] * world, (CLIP_SAMPLES,), dev) if (world > 1 and stages == "all") else None

    def e2e_step():
        r = eng.run(noise_h, enc_h, mask, 4.0, stages=stages)
        if stages != "all":
            lat_h.copy_(r["latent"], non_blocking=True)
        elif gat is None:
            out_h.copy_(r["int16"][:, :CLIP_SAMPLES], non_blocking=True)
        else:
            gat.submit(r["int16"][:, :CLIP_SAMPLES])
            gat.to_host(out_h)
    for _ in range(2):
        e2e_step()
    sync_all()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    if gat is not None:
        gat.result()        # the compute stream waits for the last gather / host copy
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)

    # ---- instrumented eager pass: CUDA events around every ctta_gemm launch -> time share of the dominant kernel
    gemm_ms, n_gemm, eager_ms = None, 0, None
    pair_ms, pair_gflop, pair_evs = 0.0, 0.0, []
    if rank == 0:
        evs = []
        orig = ops.gemm

        def timed_gemm(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(*a, **k)
            e.record()
            evs.append((s, e))
            return r
        eager = SingleStepEngine(unet, vae, use_graphs=False)
        eager.run(noise, enc, mask, 4.0, stages=stages)
        torch.cuda.synchronize(dev)
        # the module-API path without CUDA graphs: wall clock of two eager steps incl. launch overhead
        t0 = time.perf_counter()
        for _ in range(2):
            eager.run(noise, enc, mask, 4.0, stages=stages)
        torch.cuda.synchronize(dev)
        eager_ms = (time.perf_counter() - t0) * 1e3 / 2
        # the fused HiFi-GAN ResBlock pairs are tensor-core work outside ctta_gemm: timed separately, and their FLOPs are
        # taken out of the numerator of the gemm_tc_kernel roofline below
        pair_evs, orig_pair = [], ops.resblock_pair

        def timed_pair(lx, pw1, pw2, *a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig_pair(lx, pw1, pw2, *a, **k)
            e.record()
            pair_evs.append((s, e, 2.0 * 2.0 * lx.shape[0] * lx.shape[1] * lx.shape[2] * lx.shape[2] * pw1.ntaps / 1e9))
            return r
        ops.gemm = timed_gemm
        ops.resblock_pair = timed_pair
        try:
            eager.run(noise, enc, mask, 4.0, stages=stages)
            torch.cuda.synchronize(dev)
            gemm_ms = sum(s.elapsed_time(e) for s, e in evs)
            n_gemm = len(evs)
            pair_ms = sum(s.elapsed_time(e) for s, e, _ in pair_evs)
            pair_gflop = sum(f for _, _, f in pair_evs)
        finally:
            ops.gemm = orig
            ops.resblock_pair = orig_pair
        del eager

    # ---- the other BASELINE configs, short graph-replay loops (same process, same box)
    extra = {}
    K2 = max(3, min(args.steps, 10))
    if not args.unet_only and not args.no_extras:
        # configs[3]: guidance sweep w in {1,3,5} as one batch with a per-sample guidance tensor, 16 clips per GPU
        # (global 128 at N = 8), plus the external-CFG variant guidance_scale_post = 2 (2x UNet rows)
        bs = 16
        w_mix = torch.tensor([1.0, 3.0, 5.0]).repeat(bs)[:bs]
        g1, _ = capture(bs, L, w_mix, 1.0, "all", seed=300)
        ms_sweep = time_replays(g1, K2) / K2
        g2, _ = capture(bs, L, w_mix, 2.0, "all", seed=300)
        ms_sweep_cf = time_replays(g2, K2) / K2
        t = torch.tensor([ms_sweep, ms_sweep_cf], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_sweep, ms_sweep_cf = t.tolist()
        extra["guidance_sweep"] = {"config": "BASELINE configs[3]: w in {1,3,5} mixed per sample, batched guided consistency model",
                                   "global_batch": bs * world, "batch_per_gpu": bs,
                                   "clips_per_s": bs * world / ms_sweep * 1e3, "ms_per_step": ms_sweep,
                                   "guidance_scale_post_2": {"clips_per_s": bs * world / ms_sweep_cf * 1e3,
                                                             "ms_per_step": ms_sweep_cf}}
        # strong-scaling point: a FIXED global request of 128 prompts split over the N GPUs (N = 1: two 64-clip micro-batches)
        gb = 128
        if gb % world == 0:
            bl = gb // world
            n_, e_, m_ = weights.synthetic_inputs(bl, L, seed=500 + rank)
            for _ in range(2):
                eng.run(n_, e_, m_, 4.0)
            sync_all()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(K2):
                eng.run(n_, e_, m_, 4.0)
            s1.record()
            sync_all()
            t = torch.tensor([s0.elapsed_time(s1) / K2], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            extra["strong_scaling"] = {"global_batch": gb, "batch_per_gpu": bl, "ms_per_request": t.item(),
                                       "clips_per_s": gb / t.item() * 1e3,
                                       "note": "engine.run() with device-resident inputs; divide by the N = 1 figure for efficiency"}
        if world == 1:
            # configs[1]: UNet forward only at batch 32 (the second half of BASELINE.json's metric)
            gu, _ = capture(32, L, 4.0, 1.0, "unet", seed=700)
            extra["unet_b32_ms"] = time_replays(gu, max(K2, 10)) / max(K2, 10)
            # configs[4]: small batches per GPU (latency end of the throughput sweep)
            sb = {}
            for b_ in (1, 8, 16):
                gs, _ = capture(b_, L, 4.0, 1.0, "all", seed=800 + b_)
                m_ = time_replays(gs, K2) / K2
                sb[str(b_)] = {"ms_per_step": m_, "clips_per_s": b_ / m_ * 1e3}
            extra["small_batch"] = sb

    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank == 0:
        ms_step = ms / args.steps
        tf_peak, hbm_peak, peak_kind = peaks()
        clips_per_step = B * world
        if args.unet_only:
            metric, unit, value = "UNet forward ms at batch %d" % B, "ms", ms_step
            hib = False
            e2e_val = ms_e2e / args.steps
            gflop_gemm = (267.9 + 163.9) * B
            h2d = noise_h.numel() * 4 + enc_h.numel() * 4
            d2h = lat_h.numel() * 4
        else:
            metric, unit, value = "10-s audio clips/sec end-to-end", "clips/s", clips_per_step / (ms_step / 1e3)
            hib = True
            e2e_val = clips_per_step / (ms_e2e / args.steps / 1e3)
            gflop_gemm = GEMM_GFLOP_PER_CLIP * B
            h2d = (noise_h.numel() * 4 + enc_h.numel() * 4) * world
            d2h = B * world * CLIP_SAMPLES * 2
        achieved = (gflop_gemm - pair_gflop) / gemm_ms if gemm_ms else None  # GFLOP / ms == TFLOP/s; ctta_gemm launches only
        # DRAM bytes per launch of the dominant kernel from the committed ncu pass (same batch only), else null
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")) as f:
                tj = json.load(f)
            if int(tj.get("batch", -1)) == B and not args.unet_only:
                traffic = tj["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            traffic = None
        line = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": hib, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
            "config": {"workload": ("UNet2DConditionGuidedModel forward only" if args.unet_only else
                                    "full pipeline UNet + AudioLDM VAE decode + HiFi-GAN, 160000-sample 16 kHz output"),
                       "batch_per_gpu": B, "global_batch": clips_per_step, "text_len": L, "guidance": 4.0,
                       "weights": "random-init (reference architecture, deterministic synthetic state_dict)",
                       "l2": "working set (GBs of activations per step) far exceeds the 126 MB L2; no flush needed",
                       "parallelism": "dp%d (prompts sharded by batch, no collective on the hot path)" % world,
                       "cuda_graph": True},
            "e2e": {"value": e2e_val, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "output": ("latents copied to pinned host memory" if args.unet_only else
                               "every rank's int16 clips gathered to rank 0 (NCCL gather, side stream) and copied to pinned "
                               "host memory there" if gat is not None else "int16 clips copied to pinned host memory")},
            "gpu_launches": int(var.get("launches") or 0) * args.steps,
            "clocks": clocks,
            "roofline": {"kernel": "gemm_tc_kernel (tcgen05 implicit GEMM)", "bound": "tensor", "achieved": achieved,
                         "peak": tf_peak, "unit": "TFLOP/s", "frac": (achieved / tf_peak) if achieved else None,
                         "traffic": traffic, "traffic_unit": "DRAM bytes per launch (ncu, profiles/gemm_dram_traffic.json)",
                         "algorithmic_bytes_note": "tensor-bound kernel: achieved/peak are FLOP rates; see DESIGN.md 3",
                         "peak_kind": peak_kind + " bf16_tflops_sustained",
                         "launches_per_step": n_gemm, "kernel_ms_per_step": gemm_ms,
                         "share_of_step": (gemm_ms / ms_step) if gemm_ms else None,
                         "other_tensor_kernels": {"resblock_pair_kernel": {
                             "launches_per_step": len(pair_evs), "kernel_ms_per_step": pair_ms, "gflop_per_step": pair_gflop,
                             "tflops": (pair_gflop / pair_ms) if pair_ms else None,
                             "note": "fused HiFi-GAN ResBlock pairs (C <= 64); FLOPs excluded from `achieved` above"}},
                         "whole_step_tflops": (TOTAL_GFLOP_PER_CLIP if not args.unet_only else GFLOP_PER_CLIP["unet"]) * B / ms_step},
        }
        if eager_ms:
            line["eager_module_api"] = {"ms_per_step": eager_ms, ("clips_per_s" if hib else "note"):
                                        (B / eager_ms * 1e3 if hib else "UNet only"),
                                        "note2": "same kernels launched one by one from Python, no CUDA graph, wall clock"}
        line.update(extra)
        if not args.no_cpu_baseline:
            # the oracle on the host cores, on ROW 0 OF THE TIMED BATCH: it is the CPU baseline (N = 1) and the parity checker
            rows = [0] if (world > 1 or B == 1) else sorted({0, B // 2, B - 1})
            par = {"rows": rows, "latent_rel_l2": [], "mel_rel_l2": [], "wav_snr_db": [],
                   "against": "fp32 CPU oracle (pinned to the reference at 2e-6), run at batch 1 on the same rows",
                   "buffers": "outputs of the timed CUDA-graph replays"}
            for i, r in enumerate(rows):
                inp = (noise[r:r + 1], enc[r:r + 1], mask[r:r + 1])
                v, sec, threads, (lat, mel, wav) = cpu_reference_clips_per_s(L, 1, 1 if i == 0 else 0, inputs=inp)
                if i == 0 and world == 1:
                    line["cpu_baseline"] = {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                                            "sample": "2 clips (1 warm-up + 1 timed), batch 1, L=%d, fp32 oracle port, %.1f s/clip"
                                                      % (L, sec)}
                par["latent_rel_l2"].append(rel_l2(timed_out["latent"][r:r + 1], lat))
                if stages == "all":
                    par["mel_rel_l2"].append(rel_l2(timed_out["mel"][r:r + 1], mel))
                    par["wav_snr_db"].append(snr_db(timed_out["wav"][r:r + 1], wav))
            par["ok"] = bool(max(par["latent_rel_l2"]) <= 1e-2 and (stages != "all" or (
                max(par["mel_rel_l2"]) <= 1e-2 and min(par["wav_snr_db"]) >= 35.0)))
            line["parity"] = par
            if world == 1 and not args.unet_only and not args.no_extras:
                v8, sec8, threads, _ = cpu_reference_clips_per_s(L, 1, 0, batch=8)
                line["cpu_baseline"]["batch8"] = {"value": v8, "unit": "clips/s", "cores": threads,
                                                  "sample": "1 step of 8 clips, L=%d, fp32 oracle port, %.1f s/step" % (L, sec8)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
