#!/bin/sh
TAG=r02l
OUT=gpurun_out
mkdir -p $OUT
timeout 200 python -m pytest tests/test_multigpu_gpu.py tests/test_cholesky_gpu.py -m gpu -x -q -p no:cacheprovider -k "single_call or (over_the_devices and 8)" 2>&1 | tail -4
timeout 200 python tools/chol_perf.py 8 32768 512,1024 2>&1 | tee $OUT/${TAG}_chol8.txt
B200BLAS_CHOL_HOLD=0 timeout 200 python tools/chol_perf.py 8 32768 512 2>&1 | sed 's/^/hold=0 /' | tee -a $OUT/${TAG}_chol8.txt
timeout 200 python tools/chol_perf.py 1 32768 2048 2>&1 | tee $OUT/${TAG}_chol1.txt
