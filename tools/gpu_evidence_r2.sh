#!/bin/bash
# Round-2 evidence on one B200 (results under gpurun_out/r2f_*): GPU tests, smoke, bench lines, request-size sweep, ncu launch
# list of the bench command, ncu --set full of the kernels VERDICT r1 names, per-layer timing, DRAM traffic of the GEMM.
mkdir -p gpurun_out
P=gpurun_out/r2f
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > ${P}_pytest_gpu.log; cat ${P}_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -2 > ${P}_smoke.log; cat ${P}_smoke.log
python bench.py --steps 20 --warmup 3 > ${P}_bench_b64.json 2> ${P}_bench_err.log; cut -c1-300 ${P}_bench_b64.json
python bench.py --unet-only --batch 32 --steps 20 --warmup 5 --no-cpu-baseline --no-extras > ${P}_bench_unet_b32.json 2>> ${P}_bench_err.log; cut -c1-200 ${P}_bench_unet_b32.json
python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2>> ${P}_bench_err.log; cut -c1-200 ${P}_bench_reference.json
python tools/sweep.py --out ${P}_sweep.json 2>> ${P}_bench_err.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${P}_launches_b8.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > ${P}_ncu_b8.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
run_ncu () {  # name kernel-regex command...
  name=$1; shift; regex=$1; shift
  $NCU -k regex:$regex -s 3 -c 1 -f -o /tmp/$name "$@" > /dev/null 2>&1
  { python tools/ncu_summary.py /tmp/$name.ncu-rep; python tools/ncu_hot.py /tmp/$name.ncu-rep 14; } > ${P}_ncu_$name.txt 2>&1
}
run_ncu attn_tc flash_attn_tc python tools/run_one_op.py attention --b 16 --heads 5 --lq 4096 --lk 4096
run_ncu conv1d_c512 gemm_tc python tools/run_one_gemm.py conv1d --c 512 --taps 11 --dil 1 --rows 5121 --batch 64 --kind c1
run_ncu conv1d_c32_c2h gemm_tc python tools/run_one_gemm.py conv1d --c 32 --taps 3 --rows 163872 --batch 16 --kind c2h
run_ncu conv2d_c256_f32res gemm_tc python tools/run_one_gemm.py conv2d --c 256 --h 512 --w 32 --batch 16 --kind f32res
run_ncu resblock_pair_c32 resblock_pair python tools/run_one_pair.py --c 32 --taps 3 --dil 3 --t 163872 --batch 16 --iters 1
run_ncu gn_apply_f32 gn_apply python tools/run_one_op.py gn_apply --n 16 --h 1024 --w 64 --c 128
run_ncu gn_moments gn_moments python tools/run_one_op.py gn_stats --n 16 --h 1024 --w 64 --c 128
run_ncu mrf_combine mrf_combine python tools/run_one_op.py mrf_combine --n 16 --rows 40968 --c 128
run_ncu layernorm layernorm python tools/run_one_op.py layernorm --rows 262144 --d 255
python tools/profile_layers.py --batch 64 --out ${P}_layers_b64.json > ${P}_layers_b64_summary.txt 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_tc_kernel --csv --log-file /tmp/gemm_dram.csv python tools/profile_layers.py --batch 64 --out /tmp/l.json > /dev/null 2>&1
python tools/gemm_traffic.py /tmp/gemm_dram.csv 64 > ${P}_gemm_dram_traffic.json 2>> ${P}_bench_err.log
gzip -f ${P}_launches_b8.csv
du -sh gpurun_out
