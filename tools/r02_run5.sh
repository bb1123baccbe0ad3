#!/bin/sh
# round 2, 8-GPU call: partitioned GEMM behind the symbol at N = 4 and 8 (tests), bench --gpus 8 / 4 as the driver launches it, timeline
TAG=r02e
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider -k "behind_the_symbol and (4 or 8) or unmodified and 8" > $OUT/${TAG}_multigpu_tests.log 2>&1
tail -25 $OUT/${TAG}_multigpu_tests.log
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
  tail -c 800 $OUT/${TAG}_bench_n$N.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_n$N.json"))
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], "verified", d.get("verified"), "partitioned", d.get("partitioned"), "others", d.get("others"))
except Exception as e:
    print("N=$N parse error", e)
PY
done
timeout 200 python tools/mg_debug.py 8 16384 > $OUT/${TAG}_mg_debug8.txt 2>&1
grep -v "piece" $OUT/${TAG}_mg_debug8.txt | tail -30
