#!/bin/sh
# round 2, 1 GPU: panel solve with the coefficient loads off the dependency chain (TBSV / TPSV), tests + timing
TAG=r02x
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_zz_level2_struct_gpu.py -m gpu -x -q -p no:cacheprovider -k "struct or one_pass" 2>&1 | tail -3
timeout 300 python tools/l2x_perf.py 2>&1 | tee $OUT/${TAG}_l2x_perf.txt | grep -E "tpsv|tbsv|spmv|symv|hemv"
