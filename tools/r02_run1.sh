#!/bin/sh
# round 2, first GPU call: full GPU suite, bench line, TRSM launch list, Level-1/2 timings (python + C driver)
TAG=r02a
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
tail -15 $OUT/${TAG}_gpu_tests.log
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
tail -c 600 $OUT/${TAG}_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_trsm.csv python tools/trsm_target.py > $OUT/${TAG}_trsm_target.txt 2>&1
timeout 120 python tools/trsm_target.py >> $OUT/${TAG}_trsm_target.txt 2>&1
timeout 120 python tools/quick_perf_l12.py > $OUT/${TAG}_l12_perf.txt 2>&1
cat $OUT/${TAG}_l12_perf.txt
B200BLAS_L1_WAVES=4 timeout 120 python tools/quick_perf_l12.py > $OUT/${TAG}_l12_perf_waves4.txt 2>&1
ls -la $OUT | tail -8
