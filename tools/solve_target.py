"""One dtbsv_ (n = 2^15, k = 127) and one dtpsv_ (n = 8192) for an ncu capture of solve_panel_kernel (tools only)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch, libgpublas_b200 as g
g.load(); g.use_torch_stream(); g.set_sync(False)
n, k = 1 << 15, 127
ab = torch.rand((n, k + 1), dtype=torch.float64, device="cuda"); ab[:, k] = 4.0 * k
x = torch.ones(n, dtype=torch.float64, device="cuda")
g.call("dtbsv_", "U", "N", "N", n, k, ab, k + 1, x, 1); torch.cuda.synchronize()
