"""Quick device-resident timing of the GEMM family (CUDA events on the library's stream)."""
import sys, os, ctypes, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g

lib = g.load()
g.use_torch_stream()
g.set_sync(False)


def time_call(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for n in [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192]:
    A = torch.rand((n, n), dtype=torch.float64, device="cuda") * 2 - 1
    B = torch.rand((n, n), dtype=torch.float64, device="cuda") * 2 - 1
    C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    for (ta, tb) in [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")]:
        ms = time_call(lambda: g.call("dgemm_", ta, tb, n, n, n, 1.0, A, n, B, n, 0.0, C, n))
        print(f"dgemm {ta}{tb} n={n}: {ms:.3f} ms  {2.0*n**3/ms/1e9:.2f} TFLOP/s  variant={g.last_variant()}", flush=True)
    ms = time_call(lambda: torch.matmul(A, B, out=C))
    print(f"torch/cuBLAS fp64 n={n}: {ms:.3f} ms  {2.0*n**3/ms/1e9:.2f} TFLOP/s", flush=True)
