#!/bin/sh
TAG=r02f
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_cholesky_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider -k "single_call or (over_the_devices and 2) or (behind_the_symbol and 2)" > $OUT/${TAG}_tests.log 2>&1
tail -30 $OUT/${TAG}_tests.log
timeout 300 python tools/chol_perf.py 2 32768 512,1024,2048 > $OUT/${TAG}_chol2.txt 2>&1; cat $OUT/${TAG}_chol2.txt
timeout 300 python tools/chol_perf.py 1 32768 1024,2048 > $OUT/${TAG}_chol1.txt 2>&1; cat $OUT/${TAG}_chol1.txt
for K in 0 128 256; do B200BLAS_DGEMM_SMALLK=$K timeout 100 python tools/trsm_target.py 2>&1 | sed "s/^/smallk=$K /"; done | tee $OUT/${TAG}_trsm_smallk.txt
