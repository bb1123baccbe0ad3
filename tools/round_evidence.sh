#!/bin/sh
# One gpurun call that regenerates the round's evidence (run from the repository root ON the GPU box):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'sh tools/round_evidence.sh r02'
# Writes into gpurun_out/ (merged back by gpurun); copy what is to be judged into profiles/ afterwards.
#   <tag>_gpu_tests.log        pytest -m gpu
#   <tag>_bench_n1.json        python bench.py (the line the driver also produces)
#   <tag>_launches_bench.csv   ncu launch list of a short bench run (share of the step per kernel; cold-cache, serialised)
#   <tag>_full.ncu-rep         ncu --set full of the headline kernels + the structured Level-2 bodies, 1 launch each
#   <tag>_l2x_perf.txt, <tag>_l12_perf.txt   CUDA-event timings of the Level-2 family and of Level-1/2 at the symbol boundary
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
tail -3 $OUT/${TAG}_gpu_tests.log
timeout 600 python bench.py > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"dgemm_dmma_kernel|zgemm_dmma_kernel|sgemm_tf32x3|dgemv|ddot_kernel|nrm2_kernel|iamax_kernel|npart_kernel|tpart_kernel|rank_kernel|solve_panel_kernel" \
    -c 24 -o $OUT/${TAG}_full python tools/prof_targets.py dgemm sgemm zgemm l12 l2x > /dev/null 2>&1
timeout 120 python tools/l2x_perf.py > $OUT/${TAG}_l2x_perf.txt 2>&1
timeout 120 python tools/quick_perf_l12.py > $OUT/${TAG}_l12_perf.txt 2>&1
ls -la $OUT | tail -12
