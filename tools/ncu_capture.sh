#!/bin/bash
# One ncu --set full capture of the key kernels (single lane so launches are not interleaved), plus the launch list.
# Only the CSV / text summaries are kept: the .ncu-rep is too large to travel back.
export RG_B200_LANES=1
OUT=gpurun_out
mkdir -p /tmp/ncu
ncu --set full --clock-control none -k regex:'gram_fp8_tcgen05|l0_predict_tcgen05|chol_update_trsm|chol_diag|chol_backsolve|l0_stats_finish|l0_assemble|bed_relayout|bed_expand' \
    -c 60 -o /tmp/ncu/full -f python bench.py --blocks 1 --steps 1 --warmup 1 --no-cpu --no-step2 > $OUT/ncu_full.log 2>&1
ncu -i /tmp/ncu/full.ncu-rep --page raw --csv > /tmp/ncu/full_raw.csv 2>/dev/null
python tools/ncu_summarise.py /tmp/ncu/full_raw.csv > $OUT/ncu_key_kernels.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --blocks 2 --steps 1 --warmup 1 --no-cpu --no-step2 > $OUT/ncu_launch.log 2>&1
tail -2 $OUT/ncu_full.log | cut -c1-300
ls -la $OUT
