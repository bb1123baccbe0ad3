#!/bin/sh
# round 2, 1 GPU: host-staged ?syrk_/?trsm_/?trmm_ tests, one-pass symmetric Level-2 (tests, timing against the two-pass form, ncu)
TAG=r02p
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_level3_gpu.py tests/test_zz_level2_struct_gpu.py -m gpu -x -q -p no:cacheprovider -k "pipelined or struct or one_pass" > $OUT/${TAG}_tests.log 2>&1
tail -12 $OUT/${TAG}_tests.log
timeout 300 python tools/l2x_perf.py 2>&1 | sed 's/^/one-pass  /' > $OUT/${TAG}_l2x_perf.txt
B200BLAS_SYM_TWO_PASS=1 timeout 300 python tools/l2x_perf.py 2>&1 | sed 's/^/two-pass  /' >> $OUT/${TAG}_l2x_perf.txt
grep -E "spmv|sbmv|hemv|symv|hpmv" $OUT/${TAG}_l2x_perf.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sympart_kernel|sym_finish_kernel" -c 4 -o $OUT/${TAG}_sym python tools/prof_targets.py l2x > /dev/null 2>&1
ncu -i $OUT/${TAG}_sym.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_shfl.sum > $OUT/${TAG}_sym_ncu.csv 2>&1
cut -c1-600 $OUT/${TAG}_sym_ncu.csv | tail -6
python - <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, libgpublas_b200 as g
g.load(); g.set_sync(True)
hn = 8192
hA = (torch.rand((hn, hn), dtype=torch.float64) * 2 - 1).pin_memory(); hC = torch.zeros((hn, hn), dtype=torch.float64).pin_memory()
hT = torch.triu(hA).contiguous(); hT.mul_(1.0 / hn); hT.diagonal().fill_(1.0); hT = hT.pin_memory()
def wall(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = None
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best * 1e3
for name in ("dtrsm_", "dtrmm_"):
    for side in "LR":
        hC.uniform_(-1, 1)
        ms = wall(lambda: g.call(name, side, "L", "N", "N", hn, hn, 1.0, hT, hn, hC, hn))
        print("%s %sLNN 8192 host pinned: %.2f ms %.1f TFLOP/s" % (name, side, ms, float(hn) ** 3 / ms / 1e9), flush=True)
PY
