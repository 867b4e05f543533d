#!/bin/bash
# ncu --set full capture of the block-0 dilated conv and the 1x1 conv of the encoder (second forward of a B=32 bench run)
set -x
OUT=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 30 -c 2 -o $OUT/prof_gemm_encoder \
    python bench.py --sequences 32 --steps 1 --warmup 1 --no-cpu-baseline --no-fast > $OUT/prof_encoder.stdout 2> $OUT/prof_encoder.stderr
ls -la $OUT/*.ncu-rep
