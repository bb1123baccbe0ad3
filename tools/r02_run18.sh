#!/bin/sh
# round 2, 1 GPU: first-touch migration under the multiply -- tests, the preload tests that use managed blocks, the bench e2e keys
TAG=r02t
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_preload.py -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
tail -8 $OUT/${TAG}_tests.log
timeout 600 python bench.py --no-others --no-cpu 2> $OUT/${TAG}_bench.err | grep '^{' > $OUT/${TAG}_bench.json
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "pageable", d["e2e"].get("pageable"), "managed", d["e2e"].get("managed_first_touch"))
PY
tail -c 300 $OUT/${TAG}_bench.err
