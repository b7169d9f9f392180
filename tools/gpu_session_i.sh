#!/bin/bash
# GPU session I: HiFi-GAN on a single lrelu 16-bit stream, fast GELU, gn_apply MLP
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_kernels.log
tail -30 gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -rP > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
grep -E "rel|SNR|passed|failed|rc=|Error|error" gpurun_out/pytest_parity.log | head -40
{
for env in ""; do
echo "== env: $env"
env $env python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 64 --kind c1
env $env python tools/run_one_gemm.py conv1d --c 32 --taps 3 --dil 1 --rows 163872 --batch 64 --kind c2
env $env python tools/run_one_gemm.py conv1d --c 64 --taps 11 --dil 1 --rows 81936 --batch 64 --kind c1
env $env python tools/run_one_gemm.py conv1d --c 64 --taps 7 --dil 1 --rows 81936 --batch 64 --kind c2
env $env python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind f16
env $env python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 64 --kind c1
env $env python tools/run_one_gemm.py conv1d --c 128 --taps 3 --dil 1 --rows 40968 --batch 64 --kind c2
env $env python tools/run_one_gemm.py conv1d --c 128 --taps 7 --dil 1 --rows 40968 --batch 64 --kind c2
env $env python tools/run_one_gemm.py conv1d --c 256 --taps 11 --dil 1 --rows 20484 --batch 64 --kind c1
env $env python tools/run_one_gemm.py conv1d --c 256 --taps 3 --dil 1 --rows 20484 --batch 64 --kind c2
env $env python tools/run_one_gemm.py conv1d --c 512 --taps 7 --dil 3 --rows 5121 --batch 64 --kind c2
env $env python tools/run_one_gemm.py conv1d --c 512 --taps 3 --dil 1 --rows 5121 --batch 64 --kind c1
env $env python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res
env $env python tools/run_one_gemm.py conv2d --c 256 --h 512 --w 32 --batch 64 --kind f32res
env $env python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f32res
env $env python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f16
env $env python tools/run_one_gemm.py conv2d --c 256 --h 256 --w 16 --batch 64 --kind f32res
env $env python tools/run_one_gemm.py conv2d --c 1024 --h 64 --w 4 --batch 64 --kind f32res
done
} > gpurun_out/ops_timing_i.txt 2>&1
grep -v "^+" gpurun_out/ops_timing_i.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64_r1f.json 2> gpurun_out/bench_err.log
cat gpurun_out/bench_b64_r1f.json; tail -3 gpurun_out/bench_err.log
python tools/profile_layers.py --batch 64 --out gpurun_out/layers_b64_r1f.json > gpurun_out/layers_b64_r1f.txt 2>&1
head -30 gpurun_out/layers_b64_r1f.txt
timeout 600 python tools/stress.py --batch 64 --iters 20 --sync 0 > gpurun_out/stress_i.txt 2>&1; echo "rc=$?" >> gpurun_out/stress_i.txt; tail -3 gpurun_out/stress_i.txt
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128 --in16
python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind f16
