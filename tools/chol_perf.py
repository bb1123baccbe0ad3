"""Blocked Cholesky workload n=32768 through b200blas_cholesky_lower (one call, C++ driver) on `devices` GPUs, for several block sizes."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g
lib = g.load(); g.use_torch_stream(); g.set_sync(False)
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
nbs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [512, 1024, 2048]
lib.b200blas_cholesky_lower.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
lib.b200blas_cholesky_lower.restype = ctypes.c_int
torch.cuda.set_device(0)
lib.b200blas_set_options(("devices=%d" % ndev).encode())
M = torch.rand((n, n), dtype=torch.float64, device="cuda:0") * 2 - 1
M = torch.tril(M, -1); M = M + M.T; M.diagonal().fill_(float(n))
W = torch.empty_like(M)
x = torch.rand(n, dtype=torch.float64, device="cuda:0")
for nb in nbs:
    best = None
    for rep in range(3):
        W.copy_(M); torch.cuda.synchronize()
        t0 = time.perf_counter(); info = lib.b200blas_cholesky_lower(n, ctypes.c_void_p(W.data_ptr()), n, nb); torch.cuda.synchronize(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    L = torch.triu(W)          # row-major upper == column-major lower
    r = (M @ x - L.T @ (L @ x)).norm().item() / (M.norm().item() * x.norm().item())
    print("cholesky n=%d devices=%d nb=%d: %.2f ms  %.1f TFLOP/s  info=%d  resid=%.2e" % (n, ndev, nb, best * 1e3, n ** 3 / 3.0 / best / 1e12, info, r), flush=True)
