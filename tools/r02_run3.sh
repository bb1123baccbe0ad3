#!/bin/sh
TAG=r02c
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/mg_debug.py 2 16384 > $OUT/${TAG}_mg_debug.txt 2>&1
grep -v "piece" $OUT/${TAG}_mg_debug.txt | tail -40
