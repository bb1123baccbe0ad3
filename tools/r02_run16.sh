#!/bin/sh
# round 2, N=2: partitioned ?gemm_ with k-chunked bulk types + ?syrk_/?trsm_/?trmm_ parity; timing; one-pass symmetric Level-2 v2 on GPU 0
TAG=r02q
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider -k "(behind_the_symbol and 2)" > $OUT/${TAG}_tests.log 2>&1
tail -12 $OUT/${TAG}_tests.log
timeout 300 python tools/mg_bulk_perf.py 2 2>&1 | tee $OUT/${TAG}_bulk_perf_n2.txt
timeout 300 python tools/l2x_perf.py 2>&1 | grep -E "spmv|sbmv|hemv|symv" | tee $OUT/${TAG}_l2x_sym_v2.txt
timeout 300 python -m pytest tests/test_zz_level2_struct_gpu.py -m gpu -x -q -p no:cacheprovider -k "struct or one_pass" 2>&1 | tail -3
