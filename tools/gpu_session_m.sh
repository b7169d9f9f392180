#!/bin/bash
# GPU session M: final-state evidence: bench (with CPU baseline), UNet-only bench, reference arm, ncu launch list, ncu full on top kernels
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b64_r1h.json 2> gpurun_out/bench_err.log; cat gpurun_out/bench_b64_r1h.json | cut -c1-600
python bench.py --unet-only --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unet_b32_r1h.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_unet_b32_r1h.json | cut -c1-400
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_r1h.json 2>> gpurun_out/bench_err.log; cat gpurun_out/bench_reference_r1h.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r1_launches_b16.csv python bench.py --batch 16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b16.log 2>&1
tail -2 gpurun_out/ncu_b16.log | cut -c1-300
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1h_ncu_conv1d_c512 python tools/run_one_gemm.py conv1d --c 512 --taps 11 --dil 1 --rows 5121 --batch 64 --kind c1 > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1h_ncu_conv1d_c128 python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 16 --kind c1 > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1h_ncu_conv1d_c32 python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 16 --kind c1 > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1h_ncu_conv2d_c512 python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res > /dev/null 2>&1
$NCU -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/r1h_ncu_conv2d_c128 python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 16 --kind f32res > /dev/null 2>&1
$NCU -k regex:flash_attn_tc -s 3 -c 1 -f -o gpurun_out/r1h_ncu_attn_tc python tools/run_one_op.py attention --b 16 --heads 5 --lq 4096 --lk 4096 > /dev/null 2>&1
$NCU -k regex:gn_apply -s 3 -c 1 -f -o gpurun_out/r1h_ncu_gn_apply_f32 python tools/run_one_op.py gn_apply --n 16 --h 1024 --w 64 --c 128 > /dev/null 2>&1
$NCU -k regex:gn_apply -s 3 -c 1 -f -o gpurun_out/r1h_ncu_gn_apply_f16 python tools/run_one_op.py gn_apply --n 16 --h 1024 --w 64 --c 128 --in16 > /dev/null 2>&1
$NCU -k regex:layernorm -s 3 -c 1 -f -o gpurun_out/r1h_ncu_layernorm python tools/run_one_op.py layernorm --rows 262144 --d 255 > /dev/null 2>&1
ls -la gpurun_out | tail -15
