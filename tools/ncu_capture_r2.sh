#!/bin/bash
# Round-2 profile set (single lane so launches are not interleaved):
#  1. the launch list of two level-0 blocks (gpu__time_duration.sum of every launch, cold cache, serialised: SHARES)
#  2. ncu --set full of the key kernels of one block -> text summary (tools/ncu_summarise.py); the .ncu-rep stays in gpurun_out/
export RG_B200_LANES=1
OUT=gpurun_out
TAG=${1:-r2o}
CMD="python bench.py --blocks 2 --steps 1 --warmup 1 --no-cpu --no-step2"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
python tools/launch_summary.py $OUT/launches_$TAG.csv > $OUT/launches_$TAG.txt
mkdir -p /tmp/ncu
ncu --set full --clock-control none --import-source on \
    -k regex:'gram_fp8_tcgen05|l0_predict_i8|potrf128|mx_trisolve|mx_residual_fused|tf32x3_gemm|l0_assemble_sym|bed_expand_fp8|bed_relayout|l0_std_apply' \
    -s 60 -c 40 -o /tmp/ncu/full_$TAG -f $CMD > $OUT/ncu_full_$TAG.log 2>&1
ncu -i /tmp/ncu/full_$TAG.ncu-rep --page raw --csv > /tmp/ncu/full_raw_$TAG.csv 2>/dev/null
python tools/ncu_summarise.py /tmp/ncu/full_raw_$TAG.csv > $OUT/ncu_${TAG}_key_kernels.txt
tail -2 $OUT/ncu_full_$TAG.log | cut -c1-200
wc -l $OUT/ncu_${TAG}_key_kernels.txt $OUT/launches_$TAG.txt
