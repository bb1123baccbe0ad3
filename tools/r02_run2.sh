#!/bin/sh
# round 2, 2-GPU call: single-process partitioned GEMM tests + bench --gpus 2 as the driver launches it
TAG=r02b
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_multigpu_tests.log 2>&1
tail -25 $OUT/${TAG}_multigpu_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
tail -c 1500 $OUT/${TAG}_bench_n2.err
head -c 3000 $OUT/${TAG}_bench_n2.json
