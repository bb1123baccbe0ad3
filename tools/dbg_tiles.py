import torch, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libgpublas_b200 as g
from helpers import f77
lib = g.load()
torch.manual_seed(1)
nbad = 0
for it in range(60):
    for (m, n, k, tb, beta) in ((2048, 64, 64, "T", 0.0), (2048, 64, 64, "T", 1.0), (2048, 256, 256, "T", 1.0), (2048, 128, 128, "N", 1.0), (2048, 64, 448, "T", 1.0)):
        A = torch.rand((k, m), dtype=torch.float64, device="cuda")              # col-major m x k
        Bm = torch.rand((k, n) if tb == "T" else (n, k), dtype=torch.float64, device="cuda")   # 'T': stored n x k col-major = (k,n) row-major
        C0 = torch.rand((n, m), dtype=torch.float64, device="cuda"); C = C0.clone(); torch.cuda.synchronize()
        ldb = n if tb == "T" else k
        f77(lib, "dgemm_", "N", tb, m, n, k, -1.0, A, m, Bm, ldb, beta, C, m); torch.cuda.synchronize()
        Bop = Bm if tb == "T" else Bm.T                                           # k x n as row-major (k,n)
        ref = beta * C0.T - A.T @ Bop
        err = (C.T - ref).abs()
        if err.max().item() > 1e-9:
            bad = (err > 1e-9).nonzero()
            nbad += 1
            print("BAD it", it, (m, n, k, tb, beta), "n", bad.shape[0], "rows", sorted(set(bad[:, 0].tolist()))[:9], "cols", sorted(set(bad[:, 1].tolist()))[:9], flush=True)
print("done, bad cases:", nbad)
