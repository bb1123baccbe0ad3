"""CGEMM 8192^3 timing: tensor-core path (3xTF32 on tcgen05 over the doubled real product) vs the FFMA tile kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g
lib = g.load(); g.use_torch_stream(); g.set_sync(False)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for n in (4096, 8192):
    A = torch.rand((n, n), dtype=torch.complex64, device="cuda"); B = torch.rand((n, n), dtype=torch.complex64, device="cuda"); C = torch.zeros((n, n), dtype=torch.complex64, device="cuda")
    for ta, tb in (("N", "N"), ("C", "N"), ("N", "T")):
        ms = timed(lambda: g.call("cgemm_", ta, tb, n, n, n, 0.7 - 0.9j, A, n, B, n, 1.3 - 1.1j, C, n))
        print("cgemm %s%s n=%d %-16s %8.3f ms %7.1f TFLOP/s (8mnk)" % (ta, tb, n, g.last_variant(), ms, 8.0 * n ** 3 / ms / 1e9), flush=True)
    if n == 4096:
        g.force_variant("generic_tile")
        ms = timed(lambda: g.call("cgemm_", "N", "N", n, n, n, 0.7 - 0.9j, A, n, B, n, 1.3 - 1.1j, C, n), reps=1)
        print("cgemm NN n=%d %-16s %8.3f ms %7.1f TFLOP/s (8mnk)" % (n, g.last_variant(), ms, 8.0 * n ** 3 / ms / 1e9), flush=True)
        g.force_variant("auto")
    ms = timed(lambda: torch.matmul(A, B, out=C))
    print("torch.matmul complex64 n=%d (cuBLAS) %8.3f ms %7.1f TFLOP/s" % (n, ms, 8.0 * n ** 3 / ms / 1e9), flush=True)
    del A, B, C
