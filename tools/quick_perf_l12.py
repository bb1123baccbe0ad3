"""Device-resident timing of the Level-1/2 kernels at the BASELINE config sizes (CUDA events)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g

lib = g.load(); g.use_torch_stream(); g.set_sync(False)
import json
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6553.9     # the round's first measurement, used when the driver's file is absent


def time_call(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


n = 1 << 26
x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.rand(n, dtype=torch.float64, device="cuda")
big = torch.rand(1 << 28, dtype=torch.float64, device="cuda")
rows = []
rows.append(("ddot 2^26", 16.0 * n, time_call(lambda: g.call("ddot_", n, x, 1, y, 1, restype=ctypes.c_double))))
rows.append(("dnrm2 2^26", 8.0 * n, time_call(lambda: g.call("dnrm2_", n, x, 1, restype=ctypes.c_double))))
rows.append(("daxpy 2^26", 24.0 * n, time_call(lambda: g.call("daxpy_", n, 1e-9, x, 1, y, 1))))
rows.append(("idamax 2^28", 8.0 * (1 << 28), time_call(lambda: g.call("idamax_", 1 << 28, big, 1, restype=ctypes.c_int))))
rows.append(("torch.dot 2^26 (cuBLAS)", 16.0 * n, time_call(lambda: torch.dot(x, y))))
del big
m = 32768
A = torch.rand((m, m), dtype=torch.float64, device="cuda"); xv = torch.rand(m, dtype=torch.float64, device="cuda"); yv = torch.zeros(m, dtype=torch.float64, device="cuda")
rows.append(("dgemv N 32768", 8.0 * (m * m + 3 * m), time_call(lambda: g.call("dgemv_", "N", m, m, 1.0, A, m, xv, 1, 0.0, yv, 1))))
rows.append(("dgemv T 32768", 8.0 * (m * m + 3 * m), time_call(lambda: g.call("dgemv_", "T", m, m, 1.0, A, m, xv, 1, 0.0, yv, 1))))
rows.append(("torch.mv 32768 (cuBLAS)", 8.0 * (m * m + 3 * m), time_call(lambda: torch.mv(A, xv, out=yv))))
for name, byts, ms in rows:
    print(f"{name:28s} {ms:8.4f} ms  {byts/ms/1e6:8.1f} GB/s  {byts/ms/1e6/PEAK*100:5.1f}% of measured HBM peak", flush=True)
