#!/bin/bash
# GPU session F: ncu of stream vs legacy on an HBM-bound c2 conv, and of the C=32 halo conv
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1_ncu_c128k3c2_stream python tools/run_one_gemm.py conv1d --c 128 --taps 3 --dil 1 --rows 40968 --batch 16 --kind c2 > /dev/null 2>&1
CTTA_NO_STREAM=1 $NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1_ncu_c128k3c2_legacy python tools/run_one_gemm.py conv1d --c 128 --taps 3 --dil 1 --rows 40968 --batch 16 --kind c2 > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1_ncu_conv1d_c32_v2 python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 16 --kind c1 > /dev/null 2>&1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k layernorm > gpurun_out/pytest_ln.log 2>&1; tail -3 gpurun_out/pytest_ln.log
python tools/run_one_op.py layernorm --rows 262144 --d 255
python tools/run_one_op.py layernorm --rows 65536 --d 510
python tools/run_one_op.py layernorm --rows 16384 --d 1020
