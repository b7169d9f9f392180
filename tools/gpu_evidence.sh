#!/bin/bash
# Round evidence on one B200: bench lines, ncu launch list of the bench command, ncu --set full of the top kernels
# (summarised ON the box: gpurun copies back at most 64 MiB, a full-set report with sources is ~18 MiB).
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b64.json 2> gpurun_out/bench_err.log; cut -c1-700 gpurun_out/bench_b64.json
python bench.py --unet-only --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_unet_b32.json 2>> gpurun_out/bench_err.log; cut -c1-300 gpurun_out/bench_unet_b32.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_err.log; cut -c1-300 gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_b8.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b8.log 2>&1
tail -1 gpurun_out/ncu_b8.log | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on"
run_ncu () {  # name kernel-regex command...
  name=$1; shift; regex=$1; shift
  $NCU -k regex:$regex -s 3 -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
  { python tools/ncu_summary.py /tmp/$name.ncu-rep; python tools/ncu_hot.py /tmp/$name.ncu-rep 14; } > gpurun_out/ncu_$name.txt 2>&1
}
run_ncu conv1d_c512 gemm_tc python tools/run_one_gemm.py conv1d --c 512 --taps 11 --dil 1 --rows 5121 --batch 64 --kind c1
run_ncu conv1d_c128 gemm_tc python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 16 --kind c1
run_ncu conv1d_c32 gemm_tc python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 16 --kind c1
run_ncu conv1d_c32_c2h gemm_tc python tools/run_one_gemm.py conv1d --c 32 --taps 3 --rows 163872 --batch 16 --kind c2h
run_ncu conv1d_c64_c2h gemm_tc python tools/run_one_gemm.py conv1d --c 64 --taps 7 --rows 81936 --batch 16 --kind c2h
run_ncu linear_geglu gemm_tc python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind geglu
run_ncu tap_sum tap_sum python tools/run_one_op.py tap_sum --n 16 --h 1024 --w 64
run_ncu conv2d_c512 gemm_tc python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res
run_ncu conv2d_c128 gemm_tc python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 16 --kind f32res
run_ncu attn_tc flash_attn_tc python tools/run_one_op.py attention --b 16 --heads 5 --lq 4096 --lk 4096
run_ncu gn_apply_f32 gn_apply python tools/run_one_op.py gn_apply --n 16 --h 1024 --w 64 --c 128
run_ncu gn_apply_f16 gn_apply python tools/run_one_op.py gn_apply --n 16 --h 1024 --w 64 --c 128 --in16
run_ncu layernorm layernorm python tools/run_one_op.py layernorm --rows 262144 --d 255
cp /tmp/conv1d_c512.ncu-rep gpurun_out/ncu_conv1d_c512.ncu-rep
{
python tools/run_one_gemm.py conv1d --c 512 --taps 11 --dil 1 --rows 5121 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 256 --taps 11 --dil 1 --rows 20484 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 128 --taps 11 --dil 1 --rows 40968 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 64 --taps 11 --dil 1 --rows 81936 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 32 --taps 11 --dil 5 --rows 163872 --batch 64 --kind c1
python tools/run_one_gemm.py conv1d --c 512 --taps 11 --dil 1 --rows 5121 --batch 64 --kind c2h --seconds 1
python tools/run_one_gemm.py conv1d --c 64 --taps 7 --dil 1 --rows 81936 --batch 64 --kind c2h --seconds 1
python tools/run_one_gemm.py conv1d --c 32 --taps 3 --dil 1 --rows 163872 --batch 64 --kind c2h --seconds 1
python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind geglu
python tools/run_one_gemm.py conv2d --c 512 --h 256 --w 16 --batch 64 --kind f32res
python tools/run_one_gemm.py conv2d --c 256 --h 512 --w 32 --batch 64 --kind f32res
python tools/run_one_gemm.py conv2d --c 128 --h 1024 --w 64 --batch 64 --kind f32res
python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind f16
python tools/run_one_op.py attention --b 64 --heads 5 --lq 4096 --lk 4096
python tools/run_one_op.py attention --b 64 --heads 10 --lq 1024 --lk 1024
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128 --in16
python tools/run_one_op.py gn_stats --n 64 --h 1024 --w 64 --c 128
python tools/run_one_op.py layernorm --rows 262144 --d 255
python tools/run_one_op.py tap_sum --n 64 --h 1024 --w 64
python tools/run_one_op.py mrf_combine --n 64 --rows 40968 --c 128
} 2>&1 | grep -v "^+" | sed -e "s/{[^}]*}//" > gpurun_out/ops_timing.txt
cat gpurun_out/ops_timing.txt
python tools/profile_layers.py --batch 64 --out gpurun_out/layers_b64.json > gpurun_out/layers_b64_summary.txt 2>&1
du -sh gpurun_out
