#!/bin/sh
# round 2, 1 GPU: dtrsv_ through the panel solver (B200BLAS_TRSV_STRUCT=1) against the diagonal-kernel + GEMV form: tests, timing
TAG=r02y7
OUT=gpurun_out
mkdir -p $OUT
B200BLAS_TRSV_STRUCT=1 timeout 400 python -m pytest tests/test_level12_gpu.py tests/test_blat3_gpu.py tests/test_zz_level2_struct_gpu.py -m gpu -x -q -p no:cacheprovider -k "trsv or level2 or struct_fortran or micro" 2>&1 | tail -3
for S in 0 1; do B200BLAS_TRSV_STRUCT=$S python - <<'PY' 2>&1 | sed "s/^/struct=$S  /"
import sys, os
sys.path.insert(0, os.getcwd())
import torch, libgpublas_b200 as g
g.load(); g.use_torch_stream(); g.set_sync(False)
for n in (2048, 8192, 32768):
    A = torch.rand((n, n), dtype=torch.float64, device="cuda") * (1.0 / n)
    A.diagonal().fill_(2.0)
    x = torch.ones(n, dtype=torch.float64, device="cuda")
    for ul, tr in (("L", "N"), ("U", "T"), ("U", "N")):
        ts = []
        for _ in range(4):
            x.fill_(1.0); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); g.call("dtrsv_", ul, tr, "N", n, A, n, x, 1); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        print("dtrsv %s%sN n=%-6d %9.4f ms  %8.1f GB/s" % (ul, tr, n, ms, 8.0 * n * (n + 1) / 2 / ms / 1e6), flush=True)
    del A
PY
done | tee $OUT/${TAG}_trsv_struct.txt
