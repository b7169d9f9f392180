#!/bin/bash
# VAE level-1 conv2 (256 ch at 512x32, fp32 residual + fp32 out) runs at half the rate of the same conv with a 16-bit output:
# isolate it, and compare with the 16-bit-output and other-width variants, at sustained clocks
for args in "--c 256 --h 512 --w 32 --kind f32res" "--c 256 --h 512 --w 32 --kind f16" "--c 256 --h 256 --w 16 --kind f32res" "--c 128 --h 1024 --w 64 --kind f32res" "--c 128 --h 1024 --w 64 --kind f16" "--c 512 --h 256 --w 16 --kind f32res"; do
  python tools/run_one_gemm.py conv2d $args --batch 64 --taps 9 --iters 10 --seconds 0.5
  CTTA_NO_MCAST=1 python tools/run_one_gemm.py conv2d $args --batch 64 --taps 9 --iters 10 --seconds 0.5 | sed "s/^/  NO_MCAST /"
done
