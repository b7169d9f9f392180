#!/bin/bash
# ncu --set full of the short-K UNet linears at level 0 (262144 rows): qkv (K = 256 -> N = 960, f16 out) and GEGLU (K = 256 -> 2048)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r2_lin_qkv python tools/run_one_gemm.py linear --c 256 --n 960 --rows 262144 --kind f16 --iters 4 > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r2_lin_geglu python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind geglu --iters 4 > /dev/null 2>&1
for n in qkv geglu; do { python tools/ncu_summary.py gpurun_out/r2_lin_$n.ncu-rep; python tools/ncu_hot.py gpurun_out/r2_lin_$n.ncu-rep 30; } > gpurun_out/r2_ncu_lin_$n.txt 2>&1; done
python tools/run_one_gemm.py linear --c 256 --n 960 --rows 262144 --kind f16 --seconds 0.5 2>&1 | tail -1
python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind geglu --seconds 0.5 2>&1 | tail -1
