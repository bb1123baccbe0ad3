#!/bin/sh
# round 2: the bench line at N GPUs exactly as the driver launches it (+ the behind-the-symbol tests for that N)
N=${1:-2}
TAG=${2:-r02u}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_cholesky_gpu.py -m gpu -x -q -p no:cacheprovider -k "(behind_the_symbol and $N) or (over_the_devices and $N) or single_call" > $OUT/${TAG}_tests_n$N.log 2>&1
tail -6 $OUT/${TAG}_tests_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 10 --warmup 3 2> $OUT/${TAG}_bench_n$N.err | grep '^{' > $OUT/${TAG}_bench_n$N.json
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_n$N.json"))
    print("N=$N value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "verified", d.get("verified", {}).get("max_abs_diff_vs_1gpu"))
    for k, v in (d.get("others") or {}).items(): print(" ", k, v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "note"})
except Exception as e:
    print("parse error", e)
PY
tail -c 800 $OUT/${TAG}_bench_n$N.err
