#!/bin/sh
TAG=r02k
OUT=gpurun_out
mkdir -p $OUT
B200BLAS_MG_TRACE=1 timeout 200 python tools/chol_perf.py 8 32768 512 > $OUT/${TAG}_chol8_trace.txt 2>&1
grep "cholesky n=" $OUT/${TAG}_chol8_trace.txt
