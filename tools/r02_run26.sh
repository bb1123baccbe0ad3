#!/bin/sh
TAG=r02fin
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sympart_kernel|sym_finish_kernel" -c 2 -o $OUT/${TAG}_sym python tools/prof_targets.py l2x > /dev/null 2>&1
ncu -i $OUT/${TAG}_sym.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__inst_executed.sum > $OUT/${TAG}_sym_ncu.csv 2>&1
cut -c1-900 $OUT/${TAG}_sym_ncu.csv | tail -4
