#!/bin/sh
# round 2, 1 GPU: solves after hoisting the pivot division (TBSV / TPSV / TRSV), tests + timing
TAG=r02y
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_zz_level2_struct_gpu.py tests/test_level12_gpu.py tests/test_blat3_gpu.py -m gpu -x -q -p no:cacheprovider -k "struct or trsv or level2" 2>&1 | tail -3
timeout 300 python tools/l2x_perf.py 2>&1 | tee $OUT/${TAG}_l2x_perf.txt | grep -E "tpsv|tbsv"
python - <<'PY' 2>&1 | tee -a gpurun_out/r02y_l2x_perf.txt
import sys, os
sys.path.insert(0, os.getcwd())
import torch, libgpublas_b200 as g
g.load(); g.use_torch_stream(); g.set_sync(False)
n = 32768
A = torch.rand((n, n), dtype=torch.float64, device="cuda") * 1e-5
A.diagonal().fill_(2.0)
x = torch.ones(n, dtype=torch.float64, device="cuda")
for ul, tr in (("L", "N"), ("U", "T")):
    ts = []
    for _ in range(4):
        x.fill_(1.0); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); g.call("dtrsv_", ul, tr, "N", n, A, n, x, 1); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[1]
    print("dtrsv %s%sN n=32768               %9.4f ms  %8.1f GB/s  %5.1f%% of 6535 GB/s" % (ul, tr, ms, 8.0 * n * (n + 1) / 2 / ms / 1e6, 8.0 * n * (n + 1) / 2 / ms / 1e6 / 6535.4 * 100), flush=True)
PY
