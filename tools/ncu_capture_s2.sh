#!/bin/bash
# Step-2 profile set: launch list of tools/step2_probe.py (3 QT blocks of 1000 .bed variants, then 400-variant blocks of
# 8-bit dosages with one binary trait, N = 100k, inputs resident in HBM) and ncu --set full of the kernels that carry it.
OUT=gpurun_out
TAG=${1:-r2t}
CMD="python tools/step2_probe.py"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}_step2.csv $CMD > $OUT/ncu_launch_${TAG}_step2.log 2>&1
python tools/launch_summary_s2.py $OUT/launches_${TAG}_step2.csv > $OUT/launches_${TAG}_step2.txt
mkdir -p /tmp/ncu
export RG_PROBE_BLOCKS=2 RG_PROBE_REPS=1     # full-set pass: 2 QT blocks + 2 BT blocks, every matching launch captured
ncu --set full --clock-control none --import-source on \
    -k regex:'dosage_stats|dosage_relayout|s2_bt_finalize|s2_finalize|bed_expand3_fp8|gram_fp8_tcgen05|s2_stats_finish|bed_relayout|s2_firth' \
    -c 40 -o /tmp/ncu/full_${TAG}_s2 -f $CMD > $OUT/ncu_full_${TAG}_s2.log 2>&1
ncu -i /tmp/ncu/full_${TAG}_s2.ncu-rep --page raw --csv > /tmp/ncu/full_raw_${TAG}_s2.csv 2>/dev/null
python tools/ncu_summarise.py /tmp/ncu/full_raw_${TAG}_s2.csv > $OUT/ncu_${TAG}_step2_kernels.txt
cp /tmp/ncu/full_${TAG}_s2.ncu-rep $OUT/ 2>/dev/null
tail -2 $OUT/ncu_full_${TAG}_s2.log | cut -c1-200
wc -l $OUT/ncu_${TAG}_step2_kernels.txt $OUT/launches_${TAG}_step2.txt
