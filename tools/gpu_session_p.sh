#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "upsample or conv2d" 2>&1 | tail -12
timeout 900 python -m pytest tests -m gpu -x -q -rP 2>&1 | grep -E "passed|failed|rel-L2|SNR|Error|error" | head -20
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64_r1j.json 2> gpurun_out/bench_err.log; cut -c1-330 gpurun_out/bench_b64_r1j.json; tail -3 gpurun_out/bench_err.log
python tools/profile_layers.py --batch 64 --out gpurun_out/layers_b64_r1j.json > gpurun_out/layers_b64_r1j.txt 2>&1; head -24 gpurun_out/layers_b64_r1j.txt
