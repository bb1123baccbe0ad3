"""Debug timeline of the single-process partitioned DGEMM (B200BLAS_MG_TRACE=1) + raw peer-copy bandwidth idle and under a GEMM."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200BLAS_MG_TRACE"] = "1"
import torch
import libgpublas_b200 as g
lib = g.load(); g.use_torch_stream(); g.set_sync(False)
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
torch.cuda.set_device(0)
A = torch.rand((n, n), dtype=torch.float64, device="cuda:0"); B = torch.rand((n, n), dtype=torch.float64, device="cuda:0"); C = torch.zeros((n, n), dtype=torch.float64, device="cuda:0")
lib.b200blas_set_options(("devices=%d" % ndev).encode())
for it in range(3):
    sys.stderr.write("---- call %d\n" % it); sys.stderr.flush()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g.call("dgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n); e1.record(); torch.cuda.synchronize()
    sys.stderr.write("call %d: %.3f ms\n" % (it, e0.elapsed_time(e1)))
# raw peer copy bandwidth
lib.b200blas_set_options(b"devices=1")
x0 = torch.empty(1 << 27, dtype=torch.float64, device="cuda:0"); x1 = torch.empty(1 << 27, dtype=torch.float64, device="cuda:1")
lib.b200blas_copy2d_async.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
side = torch.cuda.Stream(device="cuda:0")
def copy_flat():
    with torch.cuda.stream(side):
        x1.copy_(x0, non_blocking=True)
def copy_2d():
    lib.b200blas_copy2d_async(x1.data_ptr(), 8192, x0.data_ptr(), 131072, 8192, (1 << 30) // 131072, side.cuda_stream)
for name, fn, nbytes in (("flat 1 GiB peer copy", copy_flat, 1 << 30), ("2-D 8 KiB x 8192 rows (64 MiB) peer copy", copy_2d, 8192 * ((1 << 30) // 131072))):
    for busy in (False, True):
        fn(); torch.cuda.synchronize()
        if busy:
            g.call("dgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n)      # ~245 ms of DMMA on device 0
            time.sleep(0.02)
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record(side)
        for _ in range(4):
            fn()
        s1.record(side); s1.synchronize()
        ms = s0.elapsed_time(s1) / 4
        torch.cuda.synchronize()
        sys.stderr.write("%s, source GPU %s: %.3f ms  %.1f GB/s\n" % (name, "running DGEMM" if busy else "idle", ms, nbytes / ms / 1e6))
