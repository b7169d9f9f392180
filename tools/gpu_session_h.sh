#!/bin/bash
# GPU session H: find the flaky launch failure
set -x
mkdir -p gpurun_out
timeout 600 python tools/stress.py --batch 64 --iters 8 > gpurun_out/stress_sync.txt 2>&1; echo "rc=$?" >> gpurun_out/stress_sync.txt; tail -5 gpurun_out/stress_sync.txt
timeout 600 python tools/stress.py --batch 64 --iters 8 --sync 0 > gpurun_out/stress_nosync.txt 2>&1; echo "rc=$?" >> gpurun_out/stress_nosync.txt; tail -5 gpurun_out/stress_nosync.txt
timeout 600 python tools/profile_layers.py --batch 64 --out gpurun_out/layers_b64_r1e.json > gpurun_out/layers_b64_r1e.txt 2>&1; echo "rc=$?" >> gpurun_out/layers_b64_r1e.txt; tail -5 gpurun_out/layers_b64_r1e.txt
CTTA_NO_STREAM=1 timeout 600 python tools/stress.py --batch 64 --iters 8 --sync 0 > gpurun_out/stress_nostream.txt 2>&1; echo "rc=$?" >> gpurun_out/stress_nostream.txt; tail -3 gpurun_out/stress_nostream.txt
