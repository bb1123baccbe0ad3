#!/bin/sh
# round 2 final evidence on ONE GPU: full suite, bench line, ncu launch list of the bench command, ncu --set full of the top kernels
TAG=${1:-r02s}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_gpu_tests.log 2>&1
tail -3 $OUT/${TAG}_gpu_tests.log
timeout 600 python bench.py 2> $OUT/${TAG}_bench_n1.err | grep '^{' > $OUT/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-others > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"dgemm_dmma_kernel|zgemm_dmma_kernel|sgemm_tf32x3|dgemv|ddot_kernel|nrm2_kernel|iamax_kernel|daxpy_vec|npart_kernel|tpart_kernel|rank_kernel|sympart_kernel|sym_finish_kernel" \
    -c 20 -o $OUT/${TAG}_full python tools/prof_targets.py dgemm sgemm zgemm l12 l2x > /dev/null 2>&1
ncu -i $OUT/${TAG}_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,lts__t_sector_hit_rate.pct > $OUT/${TAG}_full_raw.csv 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/${TAG}_full_raw.csv")))
hdr = None
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr and r and r[0].isdigit():
        d = dict(zip(hdr, r))
        print(d["Kernel Name"][:60], "|", " | ".join("%s=%s" % (k.split(".")[0].replace("__", ":"), d[k]) for k in hdr[11:]))
PY
timeout 200 python tools/l2x_perf.py > $OUT/${TAG}_l2x_perf.txt 2>&1
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "pageable", d["e2e"].get("pageable", {}).get("value"))
for k, v in (d.get("others") or {}).items(): print(" ", k, v if not isinstance(v, dict) else {a: b for a, b in v.items() if a not in ("note", "how")})
PY
