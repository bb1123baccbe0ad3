"""Stand-alone timings (1 GPU, idle) of the pieces of one Cholesky step at n = 32768: diagonal-block potrf, panel trsm, column update."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g
lib = g.load(); g.use_torch_stream(); g.set_sync(False)
lib.b200blas_dpotrf_lower.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
def timed(fn, prep=None, reps=5):
    ts = []
    for _ in range(reps + 1):
        if prep: prep()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts[1:])[len(ts[1:]) // 2]
n = 32768
for nb in (256, 512, 1024, 2048):
    S = torch.rand((nb, nb), dtype=torch.float64, device="cuda"); S = S @ S.T + nb * torch.eye(nb, dtype=torch.float64, device="cuda")
    W = torch.empty_like(S)
    g.set_sync(True)
    ms_potrf = timed(lambda: lib.b200blas_dpotrf_lower(nb, ctypes.c_void_p(W.data_ptr()), nb), prep=lambda: W.copy_(S))
    g.set_sync(False)
    rows = n - nb
    L = torch.triu(torch.rand((nb, nb), dtype=torch.float64, device="cuda")).contiguous(); L.mul_(1.0 / nb); L.diagonal().fill_(1.0)
    B = torch.rand((nb, rows), dtype=torch.float64, device="cuda")
    ms_trsm = timed(lambda: g.call("dtrsm_", "R", "L", "T", "N", rows, nb, 1.0, L, nb, B, rows))
    P = torch.rand((nb, rows), dtype=torch.float64, device="cuda"); C = torch.zeros((nb, rows), dtype=torch.float64, device="cuda")
    ms_upd = timed(lambda: g.call("dgemm_", "N", "T", rows, nb, nb, -1.0, P, rows, P, rows, 1.0, C, rows))
    print("nb=%4d: potrf %.3f ms | trsm %dx%d %.3f ms (%.1f TF) | column update %dx%dx%d %.3f ms (%.1f TF)" % (
        nb, ms_potrf, rows, nb, ms_trsm, rows * nb * nb / ms_trsm / 1e9, rows, nb, nb, ms_upd, 2.0 * rows * nb * nb / ms_upd / 1e9), flush=True)
    del S, W, L, B, P, C
