#!/bin/bash
# cost of the fused GroupNorm-moments epilogue (stats) on the VAE / UNet conv shapes, sustained clocks
for args in "--c 256 --h 512 --w 32 --kind f32res" "--c 128 --h 1024 --w 64 --kind f32res" "--c 512 --h 256 --w 16 --kind f32res" "--c 512 --h 256 --w 16 --kind f16" "--c 128 --h 1024 --w 64 --kind f16" "--c 256 --h 256 --w 16 --kind f32res" "--c 256 --h 256 --w 16 --kind f16"; do
  python tools/run_one_gemm.py conv2d $args --batch 64 --taps 9 --iters 10 --seconds 0.5
  python tools/run_one_gemm.py conv2d $args --batch 64 --taps 9 --iters 10 --seconds 0.5 --stats
done
