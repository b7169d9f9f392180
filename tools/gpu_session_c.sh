#!/bin/bash
# GPU session C: tcgen05 attention bring-up
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k attention > gpurun_out/pytest_attn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_attn.log
tail -30 gpurun_out/pytest_attn.log
{
timeout 120 python tools/run_one_op.py attention --b 64 --heads 5 --lq 4096 --lk 4096
timeout 120 python tools/run_one_op.py attention --b 64 --heads 10 --lq 1024 --lk 1024
timeout 120 python tools/run_one_op.py attention --b 64 --heads 20 --lq 256 --lk 256
CTTA_ATTN_LEGACY=1 timeout 120 python tools/run_one_op.py attention --b 64 --heads 20 --lq 256 --lk 256
} > gpurun_out/ops_timing_c.txt 2>&1
grep -v "^+" gpurun_out/ops_timing_c.txt | sed -e "s/{[^}]*}//"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flash_attn_tc -s 3 -c 1 -f -o gpurun_out/r1_ncu_attn_tc python tools/run_one_op.py attention --b 16 --heads 5 --lq 4096 --lk 4096 > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
