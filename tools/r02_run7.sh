#!/bin/sh
# round 2, second 8-GPU call: ordered push stream re-measured (bench N=8, 4), Cholesky over 8 devices (tests + perf), host-mode timeline
TAG=r02g
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider -k "over_the_devices and 8" > $OUT/${TAG}_tests.log 2>&1
tail -12 $OUT/${TAG}_tests.log
timeout 300 python tools/chol_perf.py 8 32768 512,1024 > $OUT/${TAG}_chol8.txt 2>&1; cat $OUT/${TAG}_chol8.txt
timeout 200 python tools/chol_perf.py 4 32768 512,1024 > $OUT/${TAG}_chol4.txt 2>&1; cat $OUT/${TAG}_chol4.txt
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 10 --warmup 3 --no-others 2> $OUT/${TAG}_bench_n$N.err | grep '^{' > $OUT/${TAG}_bench_n$N.json
  python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_n$N.json"))
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "verified", d.get("verified", {}).get("max_abs_diff_vs_1gpu"), d.get("partitioned"))
except Exception as e:
    print("N=$N parse error", e)
PY
done
timeout 200 python tools/mg_debug.py 8 16384 > $OUT/${TAG}_mg_debug8.txt 2>&1
grep -v "piece" $OUT/${TAG}_mg_debug8.txt | grep -A22 "call 2" | head -30
timeout 200 python tools/mg_debug_host.py 8 > $OUT/${TAG}_mg_debug8_host.txt 2>&1
grep -v "piece" $OUT/${TAG}_mg_debug8_host.txt | tail -24
