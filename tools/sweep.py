#!/usr/bin/env python
"""BASELINE configs[4]: throughput sweep over the request size (prompts per GPU) on this rank's GPU — clips/s of the public
engine call (device-resident inputs, CUDA graphs; requests above max_batch = 64 run as micro-batches), plus the B = 1
latency.  Under torchrun every rank runs the sweep on its own GPU and rank 0 prints the whole-job aggregate.

    python tools/sweep.py --out profiles/r2_sweep.json [--sizes 1,2,4,...]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from consistencytta_b200 import SingleStepEngine, build_random_init_models, weights  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="1,2,4,8,16,32,64,128,256,512,1024")
ap.add_argument("--text-len", type=int, default=32)
ap.add_argument("--out", default="gpurun_out/sweep.json")
a = ap.parse_args()
world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
unet, vae = build_random_init_models(dev)
eng = SingleStepEngine(unet, vae, use_graphs=True, max_buckets=8)
rows = []
for b in [int(x) for x in a.sizes.split(",")]:
    noise, enc, mask = weights.synthetic_inputs(min(b, 64), a.text_len, seed=77 + rank)
    if b > 64:   # large requests: repeat the 64 distinct prompts (content does not change the timing)
        rep = b // 64
        noise, enc, mask = noise.repeat(rep, 1, 1, 1), enc.repeat(rep, 1, 1), mask.repeat(rep, 1)
    noise, enc, mask = noise.to(dev), enc.to(dev), mask.to(dev)
    for _ in range(3):
        eng.run(noise, enc, mask, 4.0)
    iters = max(3, min(20, 2048 // b))
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        eng.run(noise, enc, mask, 4.0)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    rows.append({"prompts_per_gpu": b, "global_prompts": b * world, "ms_per_request": ms, "clips_per_s": b * world / ms * 1e3})
    if rank == 0:
        print("B/GPU %5d  x%d GPUs: %9.2f ms/request  %8.1f clips/s" % (b, world, ms, b * world / ms * 1e3), flush=True)
if rank == 0:
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"n_gpus": world, "text_len": a.text_len, "gpu": torch.cuda.get_device_name(dev),
               "note": "engine.run(), inputs resident in HBM, CUDA graphs, max_batch 64 (larger requests = micro-batches); "
                       "weak scaling: every GPU serves prompts_per_gpu", "rows": rows}, open(a.out, "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
