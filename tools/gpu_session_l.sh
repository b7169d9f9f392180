#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" 2>&1 | tail -3
python tools/attn_phases.py > gpurun_out/attn_phases_v3.txt 2>&1; cat gpurun_out/attn_phases_v3.txt
python tools/run_one_op.py attention --b 64 --heads 5 --lq 4096 --lk 4096
python tools/run_one_op.py attention --b 64 --heads 10 --lq 1024 --lk 1024
python tools/run_one_op.py attention --b 64 --heads 20 --lq 256 --lk 256
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
