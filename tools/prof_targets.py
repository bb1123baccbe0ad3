"""One launch of each headline kernel at its BASELINE.json size, for `ncu --set full` (tools only)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g
g.load(); g.use_torch_stream(); g.set_sync(False)
which = sys.argv[1:] or ["dgemm", "sgemm", "zgemm", "l12"]
if "dgemm" in which:
    n = 16384
    A = torch.rand((n, n), dtype=torch.float64, device="cuda"); B = torch.rand((n, n), dtype=torch.float64, device="cuda"); C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    g.call("dgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n); torch.cuda.synchronize(); del A, B, C
if "sgemm" in which:
    n = 16384
    A = torch.rand((n, n), dtype=torch.float32, device="cuda"); B = torch.rand((n, n), dtype=torch.float32, device="cuda"); C = torch.zeros((n, n), dtype=torch.float32, device="cuda")
    g.call("sgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n); torch.cuda.synchronize(); del A, B, C
if "zgemm" in which:
    n = 8192
    A = torch.rand((n, n), dtype=torch.complex128, device="cuda"); B = torch.rand((n, n), dtype=torch.complex128, device="cuda"); C = torch.zeros((n, n), dtype=torch.complex128, device="cuda")
    g.call("zgemm_", "N", "N", n, n, n, 1.0 + 0j, A, n, B, n, 0j, C, n); torch.cuda.synchronize(); del A, B, C
if "l12" in which:
    m = 32768
    A = torch.rand((m, m), dtype=torch.float64, device="cuda"); x = torch.rand(m, dtype=torch.float64, device="cuda"); y = torch.zeros(m, dtype=torch.float64, device="cuda")
    g.call("dgemv_", "N", m, m, 1.0, A, m, x, 1, 0.0, y, 1); g.call("dgemv_", "T", m, m, 1.0, A, m, x, 1, 0.0, y, 1); torch.cuda.synchronize(); del A
    n = 1 << 26
    x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.rand(n, dtype=torch.float64, device="cuda")
    g.call("ddot_", n, x, 1, y, 1, restype=ctypes.c_double); g.call("daxpy_", n, 1e-9, x, 1, y, 1); g.call("dnrm2_", n, x, 1, restype=ctypes.c_double)
    z = torch.rand(1 << 28, dtype=torch.float64, device="cuda")
    g.call("idamax_", 1 << 28, z, 1, restype=ctypes.c_int); torch.cuda.synchronize()
if "l2x" in which:      # banded / packed Level-2 (csrc/level2_struct.cu): N part, T part, rank row, panel solve
    nb, kl, ku = 1 << 22, 63, 64
    ab = torch.rand((nb, kl + ku + 1), dtype=torch.float64, device="cuda"); xb = torch.rand(nb, dtype=torch.float64, device="cuda"); yb = torch.zeros(nb, dtype=torch.float64, device="cuda")
    g.call("dgbmv_", "N", nb, nb, kl, ku, 1.0, ab, kl + ku + 1, xb, 1, 0.0, yb, 1); g.call("dgbmv_", "T", nb, nb, kl, ku, 1.0, ab, kl + ku + 1, xb, 1, 0.0, yb, 1)
    torch.cuda.synchronize(); del ab, xb, yb
    n = 32768
    ap = torch.rand(n * (n + 1) // 2, dtype=torch.float64, device="cuda") * 1e-5; x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.rand(n, dtype=torch.float64, device="cuda")
    g.call("dtpmv_", "U", "N", "N", n, ap, x, 1); g.call("dtpmv_", "U", "T", "N", n, ap, x, 1); g.call("dspmv_", "U", n, 1.0, ap, x, 1, 0.0, y, 1)
    g.call("dspr2_", "L", n, 1e-9, y, 1, x, 1, ap); torch.cuda.synchronize()
