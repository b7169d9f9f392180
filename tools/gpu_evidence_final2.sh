mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/final2_bench_b64.json 2> gpurun_out/final2_err.log; cut -c1-300 gpurun_out/final2_bench_b64.json
python bench.py --unet-only --batch 32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/final2_bench_unet_b32.json 2>> gpurun_out/final2_err.log; cut -c1-200 gpurun_out/final2_bench_unet_b32.json
NCU="ncu --set full --clock-control none --import-source on"
for mc in 0 1; do
  if [ $mc = 0 ]; then export CTTA_NO_MCAST=1; else unset CTTA_NO_MCAST; fi
  $NCU -k regex:gemm_tc -s 3 -c 1 -f -o /tmp/c512_mc$mc python tools/run_one_gemm.py conv1d --c 512 --taps 11 --dil 1 --rows 5121 --batch 64 --kind c1 > /dev/null 2>&1
  { python tools/ncu_summary.py /tmp/c512_mc$mc.ncu-rep; python tools/ncu_hot.py /tmp/c512_mc$mc.ncu-rep 10; } > gpurun_out/final2_ncu_conv1d_c512_mcast$mc.txt 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/final2_launches_b8.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/final2_ncu_b8.log 2>&1
python tools/profile_layers.py --batch 64 --out gpurun_out/final2_layers_b64.json > gpurun_out/final2_layers_b64_summary.txt 2>&1
head -20 gpurun_out/final2_layers_b64_summary.txt
grep -E "lts__t_bytes|gpu__time_duration|pipe_tensor_cycles_active" gpurun_out/final2_ncu_conv1d_c512_mcast*.txt
