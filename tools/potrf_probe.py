import torch, sys, time
sys.path.insert(0, ".")
import libgpublas_b200 as g
from libgpublas_b200.cholesky import potrf_lower
g.load(); g.use_torch_stream()
n = 2048
A = torch.rand((n, n), dtype=torch.float64, device="cuda") * 2 - 1
A = torch.tril(A, -1); A = A + A.T; A.diagonal().fill_(float(n))
B = A.clone(); torch.cuda.synchronize()
potrf_lower(n, B, n); B.copy_(A); torch.cuda.synchronize()
t0 = time.perf_counter(); potrf_lower(n, B, n); torch.cuda.synchronize(); print("potrf 2048 wall ms", (time.perf_counter() - t0) * 1e3)
