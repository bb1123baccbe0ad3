#!/bin/sh
# round 2, N=2: partitioned ?syrk_/?trsm_/?trmm_ behind the symbol -- parity test, then timing on 1 and 2 devices
TAG=r02n
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider -k "syrk_trsm_trmm and 2" > $OUT/${TAG}_tests.log 2>&1
tail -15 $OUT/${TAG}_tests.log
timeout 300 python tools/ml3_perf.py 2 16384 2>&1 | tee $OUT/${TAG}_ml3_perf_n2.txt
timeout 300 python -m pytest tests/test_level12_gpu.py -m gpu -x -q -p no:cacheprovider -k "bounce" 2>&1 | tail -5
