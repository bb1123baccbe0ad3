#!/bin/sh
TAG=r02i
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_level3_gpu.py tests/test_blat3_gpu.py tests/test_cholesky_gpu.py tests/test_preload.py -m gpu -x -q -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1
tail -8 $OUT/${TAG}_tests.log
for CO in 0 1024; do B200BLAS_POTRF_COOP=$CO timeout 100 python tools/chol_parts.py 2>&1 | sed "s/^/potrf_coop<=$CO /"; done | tee $OUT/${TAG}_chol_parts.txt
for S in 0 1; do B200BLAS_TRSM_SPLIT=$S timeout 100 python tools/trsm_target.py 2>&1 | sed "s/^/split=$S /"; done | tee $OUT/${TAG}_trsm_split.txt
timeout 200 python tools/chol_perf.py 1 32768 1024,2048 2>&1 | tee $OUT/${TAG}_chol1.txt
python - <<'PY' 2>&1 | tee gpurun_out/r02i_trsm_shapes.txt
import sys, os
sys.path.insert(0, os.getcwd())
import torch, libgpublas_b200 as g
g.load(); g.use_torch_stream(); g.set_sync(False)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
m = 8192
A = torch.rand((m, m), dtype=torch.float64, device="cuda")
T = torch.triu(A).contiguous(); T.mul_(1.0 / m); T.diagonal().fill_(1.0)
B = torch.rand((m, m), dtype=torch.float64, device="cuda")
for side, uplo, tr in (("L","L","N"), ("L","L","T"), ("R","L","N"), ("R","L","T")):
    B.uniform_(-1, 1)
    ms = timed(lambda: g.call("dtrsm_", side, uplo, tr, "N", m, m, 1.0, T, m, B, m))
    print("dtrsm %s%s%sN 8192x8192: %.3f ms %.2f TFLOP/s" % (side, uplo, tr, ms, float(m)**3/ms/1e9))
    B.uniform_(-1, 1)
    ms = timed(lambda: g.call("dtrmm_", side, uplo, tr, "N", m, m, 1.0, T, m, B, m))
    print("dtrmm %s%s%sN 8192x8192: %.3f ms %.2f TFLOP/s" % (side, uplo, tr, ms, float(m)**3/ms/1e9))
PY
OB=$(python -c "import sys; sys.path.insert(0,'tests'); from helpers import find_openblas; import os; print(os.path.dirname(find_openblas()))")
python -c "import sys; sys.path.insert(0,'tests'); from test_preload import build_driver; build_driver('l1_chain')"
LD_LIBRARY_PATH=$OB LD_PRELOAD=$PWD/libgpublas_b200/libb200blas.so timeout 120 tests/drivers/_build/l1_chain 67108864 40 268435456 2>&1 | tee $OUT/${TAG}_l1_chain_c.txt
