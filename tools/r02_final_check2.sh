#!/bin/sh
TAG=r02final2
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
tail -3 $OUT/${TAG}_gpu_tests.log
timeout 120 python tools/l2x_perf.py 2>&1 | grep -E "spmv|symv|hemv|sbmv" | tee $OUT/${TAG}_l2x_sym.txt
