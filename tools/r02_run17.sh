#!/bin/sh
TAG=r02r
OUT=gpurun_out
mkdir -p $OUT
for K in 1 2 4 8; do B200BLAS_MG_KCHUNKS=$K timeout 200 python tools/mg_bulk_perf.py 2 2>&1 | grep "devices=2" | sed "s/^/kchunks=$K /"; done | tee $OUT/${TAG}_kchunks_n2.txt
B200BLAS_MG_KCHUNKS=4 timeout 100 python tools/mg_bulk_trace.py 2 s 2>&1 | awk '/---- call 2/{p=1} p' | grep -E "kernel|call|enqueue|piece 0:|piece 3:" | head -60 | tee $OUT/${TAG}_trace_sgemm_n2.txt
B200BLAS_MG_KCHUNKS=4 timeout 100 python tools/mg_bulk_trace.py 2 c 2>&1 | awk '/---- call 2/{p=1} p' | grep -E "kernel|call|enqueue" | head -30 | tee $OUT/${TAG}_trace_cgemm_n2.txt
timeout 300 python tools/l2x_perf.py 2>&1 | grep -E "tpmv|gbmv|spr|her2|geru" | tee $OUT/${TAG}_l2x_npart_v5.txt
