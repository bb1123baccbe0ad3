#!/bin/sh
# round 2, 1 GPU: one-pass symmetric product v3 (v in shared memory, ping-pong loads, faster finish), 4 vs 5 CTAs per SM
TAG=r02w
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_zz_level2_struct_gpu.py tests/test_level12_gpu.py -m gpu -x -q -p no:cacheprovider -k "struct or one_pass or level2_more" 2>&1 | tail -3
for B in 5 4; do B200BLAS_SYM_MINB=$B timeout 300 python tools/l2x_perf.py 2>&1 | grep -E "spmv|sbmv|hemv|symv" | sed "s/^/minb=$B  /"; done | tee $OUT/${TAG}_l2x_sym_v3.txt
