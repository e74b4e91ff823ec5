#!/bin/bash
# ncu --set full (with source) of the mixed solver's kernels on one block, single lane; the reports travel back for
# `ncu -i ... --page source --csv` here.  Launch numbers: -s skips the warm-up block's launches of the same kernel.
export RG_B200_LANES=1
OUT=gpurun_out
CMD="python bench.py --blocks 1 --steps 1 --warmup 1 --no-cpu --no-step2"
ncu --set full --clock-control none --import-source on -k regex:potrf128 -s 28 -c 1 -o $OUT/prof_r2k_potrf -f $CMD > $OUT/ncu_r2k_potrf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tf32x3_gemm -s 52 -c 2 -o $OUT/prof_r2k_gemm -f $CMD > $OUT/ncu_r2k_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mx_trisolve|mx_residual_fused|l0_predict_i8' -s 8 -c 3 -o $OUT/prof_r2k_misc -f $CMD > $OUT/ncu_r2k_misc.log 2>&1
ls -la $OUT/*.ncu-rep
