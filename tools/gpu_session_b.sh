#!/bin/bash
# GPU session B: re-validate after elect.sync + gn_apply v2; timings + bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
{
python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 32 --taps 3 --dil 1 --rows 163872 --batch 64 --kind c2
python tools/run_one_gemm.py conv1d --c 64 --taps 11 --dil 1 --rows 81936 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 256 --taps 11 --dil 1 --rows 20484 --batch 64 --kind c1
python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res
python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f32res
python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind f16
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128 --in16
python tools/run_one_op.py gn_apply --n 64 --h 256 --w 16 --c 512
} > gpurun_out/ops_timing_b.txt 2>&1
grep -v "^+" gpurun_out/ops_timing_b.txt
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64_r1c.json 2> gpurun_out/bench_err.log
cat gpurun_out/bench_b64_r1c.json; tail -3 gpurun_out/bench_err.log
