"""Summarise an ncu --page raw --csv export: one block per kernel launch with the metrics DESIGN.md quotes."""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__registers_per_thread", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("---")
    print("%-72s %s" % ("Kernel Name", d.get("Kernel Name", "")[:110]))
    for k in KEEP:
        if k in d:
            print("%-72s %s %s" % (k, d[k], u.get(k, "")))
