#!/bin/bash
# per-kernel device time + tensor-pipe + DRAM bytes of one encoder pass (128 x 1024 aa) and one scorer chunk, current defaults
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,sm__cycles_elapsed.max
OUT=gpurun_out
ncu -k regex:"gemm|conv_input|pool_mean" --metrics $M --clock-control none -s 0 -c 14 --csv --log-file $OUT/r02_enc_launches.csv python tools/encoder_probe.py 128 1024 0 > $OUT/ll_enc.log 2>&1
ncu -k regex:"gemm|pair_features|finalize|split_rows" --metrics $M --clock-control none -s 0 -c 40 --csv --log-file $OUT/r02_scorer_launches.csv python tools/scorer_launch_probe.py > $OUT/ll_sc.log 2>&1
tail -2 $OUT/ll_enc.log $OUT/ll_sc.log
