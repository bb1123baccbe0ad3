"""One dtrsm_ of the Cholesky panel shape (X * L^T = B, L 2048^2, B 30720 x 2048) + one left-side 8192^2, for an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g
g.load(); g.use_torch_stream(); g.set_sync(False)
mm, nn = 30720, 2048
A = torch.rand((nn, nn), dtype=torch.float64, device="cuda")
Lp = torch.triu(A).contiguous(); Lp.mul_(1.0 / nn); Lp.diagonal().fill_(1.0)
Bp = torch.rand((nn, mm), dtype=torch.float64, device="cuda")
for _ in range(2):
    g.call("dtrsm_", "R", "L", "T", "N", mm, nn, 1.0, Lp, nn, Bp, mm)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); g.call("dtrsm_", "R", "L", "T", "N", mm, nn, 1.0, Lp, nn, Bp, mm); e1.record(); e1.synchronize()
ms = e0.elapsed_time(e1)
print("dtrsm RLTN 30720x2048: %.3f ms  %.2f TFLOP/s" % (ms, mm * nn * nn / ms / 1e9))
