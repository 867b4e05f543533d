#!/bin/bash
# Round profile capture (run under gpurun, 1 GPU):
#  1. launch list (per-kernel device time) of a short bench.py run -> share of the step per kernel
#  2. one `--set full` capture of the dominant kernel (pair-scorer GEMM) and of the encoder's dilated-conv GEMM
set -x
OUT=gpurun_out
if [ -z "$SKIP_LAUNCH_LIST" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --sequences 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/launches_bench.stdout 2> $OUT/launches_bench.stderr
fi
# one forward = 29 GEMM launches (11 encoder, 5 W_p, 5 W_l, 8 scorer); take them from the second forward:
# launches 50/51 = scorer layer 2 (split-store epilogue) / layer 3 (dot epilogue); 30/31 = block-0 dilated / pointwise conv
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 50 -c 2 -o $OUT/prof_gemm_scorer \
    python bench.py --sequences 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/prof_scorer.stdout 2> $OUT/prof_scorer.stderr
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 30 -c 2 -o $OUT/prof_gemm_encoder \
    python bench.py --sequences 32 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/prof_encoder.stdout 2> $OUT/prof_encoder.stderr
ls -la $OUT/*.ncu-rep
