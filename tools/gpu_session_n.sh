#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -rP 2>&1 | grep -E "passed|failed|rel-L2|SNR|Error" | head -20
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64_r1i.json 2> gpurun_out/bench_err.log; cut -c1-330 gpurun_out/bench_b64_r1i.json
python tools/profile_layers.py --batch 64 --out gpurun_out/layers_b64_r1i.json > gpurun_out/layers_b64_r1i.txt 2>&1; head -24 gpurun_out/layers_b64_r1i.txt
