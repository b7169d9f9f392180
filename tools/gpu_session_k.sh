#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/attn_phases.py > gpurun_out/attn_phases.txt 2>&1; cat gpurun_out/attn_phases.txt
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128 --in16
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "groupnorm or attention" 2>&1 | tail -3
