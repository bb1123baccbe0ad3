#!/bin/sh
# round 2, 1-GPU call: CGEMM on tcgen05 (tests + perf), TRSM after the leaf fix, Level-1 ncu + C driver
TAG=r02d
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
tail -15 $OUT/${TAG}_gpu_tests.log
timeout 200 python tools/cgemm_perf.py > $OUT/${TAG}_cgemm_perf.txt 2>&1; cat $OUT/${TAG}_cgemm_perf.txt
timeout 120 python tools/trsm_target.py > $OUT/${TAG}_trsm.txt 2>&1; cat $OUT/${TAG}_trsm.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_trsm.csv python tools/trsm_target.py > /dev/null 2>&1
OB=$(python -c "import sys; sys.path.insert(0,'tests'); from helpers import find_openblas; import os; print(os.path.dirname(find_openblas()))")
python -c "import sys; sys.path.insert(0,'tests'); from test_preload import build_driver; build_driver('l1_chain'); build_driver('cg_chain')"
LD_LIBRARY_PATH=$OB LD_PRELOAD=$PWD/libgpublas_b200/libb200blas.so timeout 120 tests/drivers/_build/l1_chain 67108864 40 > $OUT/${TAG}_l1_chain_c.txt 2>&1; cat $OUT/${TAG}_l1_chain_c.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"ddot_kernel|nrm2_kernel|iamax_kernel|daxpy_vec" -c 8 -o $OUT/${TAG}_l1 python tools/prof_targets.py l12 > /dev/null 2>&1
ncu -i $OUT/${TAG}_l1.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__cycles_active.avg > $OUT/${TAG}_l1_ncu.csv 2>&1
cat $OUT/${TAG}_l1_ncu.csv | cut -c1-400
ls -la $OUT | tail -8
