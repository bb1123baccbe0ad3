"""Timeline of the partitioned DGEMM with HOST-resident (pinned) operands (B200BLAS_MG_TRACE=1)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["B200BLAS_MG_TRACE"] = "1"
import torch
import libgpublas_b200 as g
lib = g.load()
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
n = 16384
torch.cuda.set_device(0)
hA = torch.rand((n, n), dtype=torch.float64).pin_memory(); hB = torch.rand((n, n), dtype=torch.float64).pin_memory(); hC = torch.empty((n, n), dtype=torch.float64).pin_memory()
lib.b200blas_set_options(("devices=%d" % ndev).encode())
for it in range(3):
    sys.stderr.write("---- host call %d\n" % it); sys.stderr.flush()
    t0 = time.perf_counter(); g.call("dgemm_", "N", "N", n, n, n, 1.0, hA.numpy(), n, hB.numpy(), n, 0.0, hC.numpy(), n); dt = time.perf_counter() - t0
    sys.stderr.write("host call %d: %.3f ms\n" % (it, dt * 1e3))
