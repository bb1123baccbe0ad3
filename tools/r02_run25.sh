#!/bin/sh
TAG=r02fin
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python bench.py 2> $OUT/${TAG}_bench_n1.err | grep '^{' > $OUT/${TAG}_bench_n1.json
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "pageable", d["e2e"].get("pageable", {}).get("value"), "managed", d["e2e"].get("managed_first_touch", {}).get("value"))
for k, v in (d.get("others") or {}).items(): print(" ", k, v if not isinstance(v, dict) else {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a not in ("note", "how")})
PY
tail -c 300 $OUT/${TAG}_bench_n1.err
