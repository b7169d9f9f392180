#!/bin/bash
# GPU session J: batched VAE attention GEMMs, tanh SiLU on 16-bit GN apply
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/pytest_kernels.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_kernels.log
tail -30 gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -rP > gpurun_out/pytest_parity.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_parity.log
grep -E "rel|SNR|passed|failed|rc=|Error|error" gpurun_out/pytest_parity.log | head -40
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b64_r1g.json 2> gpurun_out/bench_err.log
cat gpurun_out/bench_b64_r1g.json; tail -3 gpurun_out/bench_err.log
python tools/profile_layers.py --batch 64 --out gpurun_out/layers_b64_r1g.json > gpurun_out/layers_b64_r1g.txt 2>&1
head -30 gpurun_out/layers_b64_r1g.txt
timeout 600 python tools/stress.py --batch 64 --iters 20 --sync 0 > gpurun_out/stress_j.txt 2>&1; echo "rc=$?" >> gpurun_out/stress_j.txt; tail -3 gpurun_out/stress_j.txt
python tools/run_one_op.py gn_apply --n 64 --h 1024 --w 64 --c 128 --in16
python tools/run_one_gemm.py linear --c 256 --n 2048 --rows 262144 --kind f16
