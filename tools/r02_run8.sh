#!/bin/sh
TAG=r02h
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python tools/chol_parts.py > $OUT/${TAG}_chol_parts.txt 2>&1; cat $OUT/${TAG}_chol_parts.txt
B200BLAS_MG_TRACE=1 timeout 200 python tools/chol_perf.py 4 32768 512 > $OUT/${TAG}_chol4_trace.txt 2>&1
grep "cholesky n=" $OUT/${TAG}_chol4_trace.txt
timeout 200 python tools/mg_debug_host.py 4 > $OUT/${TAG}_mg_debug4_host.txt 2>&1
grep -v "piece" $OUT/${TAG}_mg_debug4_host.txt | tail -14
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider -k "behind_the_symbol and 4" 2>&1 | tail -3
