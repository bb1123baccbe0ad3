#!/bin/sh
TAG=r02z
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:solve_panel_kernel -c 1 -o $OUT/${TAG}_solve python tools/solve_target.py > /dev/null 2>&1
ls -la $OUT/${TAG}_solve.ncu-rep
