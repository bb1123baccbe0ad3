#!/bin/sh
# round 2, third 8-GPU call: Cholesky over 8 devices with the chain fixes (+ tests), full bench line at N = 8
TAG=r02j
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_cholesky_gpu.py tests/test_multigpu_gpu.py -m gpu -x -q -p no:cacheprovider -k "single_call or (over_the_devices and 8) or (behind_the_symbol and 8)" > $OUT/${TAG}_tests.log 2>&1
tail -8 $OUT/${TAG}_tests.log
timeout 300 python tools/chol_perf.py 8 32768 256,512,1024 > $OUT/${TAG}_chol8.txt 2>&1; cat $OUT/${TAG}_chol8.txt
B200BLAS_CHOL_HOLD=0 timeout 200 python tools/chol_perf.py 8 32768 512 2>&1 | sed 's/^/hold=0 /' | tee -a $OUT/${TAG}_chol8.txt
timeout 200 python tools/chol_perf.py 4 32768 512,1024 > $OUT/${TAG}_chol4.txt 2>&1; cat $OUT/${TAG}_chol4.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29638 bench.py --gpus 8 --steps 10 --warmup 3 2> $OUT/${TAG}_bench_n8.err | grep '^{' > $OUT/${TAG}_bench_n8.json
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_n8.json"))
    print("N=8 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"] and (d["e2e"]["value"], d["e2e"]["ms_per_step"]), "verified", d.get("verified", {}).get("max_abs_diff_vs_1gpu"))
    for k, v in (d.get("others") or {}).items(): print(" ", k, v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "note"})
except Exception as e:
    print("N=8 parse error", e)
PY
tail -c 600 $OUT/${TAG}_bench_n8.err
