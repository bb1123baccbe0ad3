#!/bin/sh
# round 2: full single-GPU suite with the pageable bounce ring + version-scripted exports, then the N=1 bench line
TAG=r02o
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
tail -6 $OUT/${TAG}_tests.log
timeout 600 python bench.py 2> $OUT/${TAG}_bench_n1.err | grep '^{' > $OUT/${TAG}_bench_n1.json
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "pageable", d["e2e"].get("pageable"), "managed", d["e2e"].get("managed_first_touch"))
for k, v in (d.get("others") or {}).items(): print(" ", k, v if not isinstance(v, dict) else {a: b for a, b in v.items() if a != "note"})
PY
tail -c 400 $OUT/${TAG}_bench_n1.err
