#!/bin/bash
# GPU session A: parity tests, UMMA microbenchmark, per-kernel timings and ncu captures of representative kernels.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 tools/microbench/umma_rate > gpurun_out/umma_rate.txt 2>&1
{
python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 32 --taps 3 --dil 1 --rows 163872 --batch 64 --kind c2
python tools/run_one_gemm.py conv1d --c 64 --taps 11 --dil 1 --rows 81936 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 256 --taps 11 --dil 1 --rows 20484 --batch 64 --kind c1
python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res
python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f32res
python tools/run_one_op.py attention --b 64 --heads 5 --lq 4096 --lk 4096
python tools/run_one_op.py attention --b 64 --heads 10 --lq 1024 --lk 1024
python tools/run_one_op.py attention --b 64 --heads 5 --lq 4096 --lk 32
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128 --in16
python tools/run_one_op.py gn_stats --n 64 --h 1024 --w 64 --c 128
python tools/run_one_op.py gn_stats --n 64 --h 1024 --w 64 --c 128 --in16
python tools/run_one_op.py layernorm --rows 262144 --d 255
} > gpurun_out/ops_timing.txt 2>&1
cat gpurun_out/ops_timing.txt
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1_ncu_conv1d_c32 python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 16 --kind c1 > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1_ncu_conv1d_c128 python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 16 --kind c1 > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1_ncu_conv2d_c512 python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res > /dev/null 2>&1
$NCU -k regex:flash_attn -s 3 -c 1 -f -o gpurun_out/r1_ncu_attn python tools/run_one_op.py attention --b 16 --heads 5 --lq 4096 --lk 4096 > /dev/null 2>&1
$NCU -k regex:gn_apply -s 3 -c 1 -f -o gpurun_out/r1_ncu_gn_apply python tools/run_one_op.py gn_apply --n 16 --h 1024 --w 64 --c 128 > /dev/null 2>&1
ls -la gpurun_out
