#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --unet-only --batch 32 --steps 10 --warmup 3 --no-cpu-baseline | cut -c1-200
CTTA_NO_NARROW_TILES=1 python bench.py --unet-only --batch 32 --steps 10 --warmup 3 --no-cpu-baseline | cut -c1-200
python bench.py --steps 5 --warmup 3 --no-cpu-baseline | cut -c1-200
