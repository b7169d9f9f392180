#!/bin/bash
# Final round-2 evidence on one B200: GPU tests, smoke, both bench arms, launch list of the bench command (share of each
# kernel in a step), ncu --set full of the fused ResBlock-pair kernel.  Results under gpurun_out/r2z_*.
mkdir -p gpurun_out
P=gpurun_out/r2z
timeout 900 python -m pytest tests -m gpu -q > ${P}_pytest_gpu.log 2>&1; tail -2 ${P}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > ${P}_smoke.log 2>&1; tail -1 ${P}_smoke.log
python bench.py > ${P}_bench_b64.json 2> ${P}_bench_err.log; cut -c1-200 ${P}_bench_b64.json
python bench.py --impl reference --steps 2 --warmup 1 > ${P}_bench_reference.json 2>> ${P}_bench_err.log; cut -c1-200 ${P}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file ${P}_launches_b8.csv python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --no-extras > ${P}_ncu_b8.log 2>&1
python tools/launch_list_summary.py ${P}_launches_b8.csv > ${P}_launches_b8_summary.txt 2>&1; head -12 ${P}_launches_b8_summary.txt
NCU="ncu --set full --clock-control none --import-source on"
for cfg in "64 3 1 81936 c64k3" "32 3 1 163872 c32k3" "32 11 5 163872 c32k11"; do
  set -- $cfg
  $NCU -k regex:resblock_pair -s 3 -c 1 -f -o /tmp/pair_$5 python tools/run_one_pair.py --c $1 --taps $2 --dil $3 --t $4 --batch 16 --iters 1 > /dev/null 2>&1
  { python tools/ncu_summary.py /tmp/pair_$5.ncu-rep; python tools/ncu_lines.py /tmp/pair_$5.ncu-rep 12; } > ${P}_ncu_resblock_pair_$5.txt 2>&1
done
gzip -f ${P}_launches_b8.csv
