#!/bin/bash
# GPU session D: stream-mode GEMM bring-up
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "conv" > gpurun_out/pytest_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_conv.log
tail -30 gpurun_out/pytest_conv.log
{
for env in "" "CTTA_NO_STREAM=1"; do
echo "== env: $env"
env $env python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 64 --kind c1
env $env python tools/run_one_gemm.py conv1d --c 128 --taps 3 --dil 1 --rows 40968 --batch 64 --kind c2
env $env python tools/run_one_gemm.py conv1d --c 256 --taps 11 --dil 1 --rows 20484 --batch 64 --kind c1
env $env python tools/run_one_gemm.py conv1d --c 512 --taps 7 --dil 3 --rows 5121 --batch 64 --kind c2
env $env python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res
env $env python tools/run_one_gemm.py conv2d --c 256 --h 512 --w 32 --batch 64 --kind f32res
env $env python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f32res
env $env python tools/run_one_gemm.py conv2d --c 1024 --h 64 --w 4 --batch 64 --kind f32res
done
} > gpurun_out/ops_timing_d.txt 2>&1
grep -v "^+" gpurun_out/ops_timing_d.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
