import torch, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libgpublas_b200 as g
from helpers import f77
lib = g.load()
torch.manual_seed(1)
m, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tb = sys.argv[4] if len(sys.argv) > 4 else "T"
nbad = 0
for it in range(100):
    A = torch.rand((k, m), dtype=torch.float64, device="cuda")
    Bm = torch.rand((k, n) if tb == "T" else (n, k), dtype=torch.float64, device="cuda")
    C0 = torch.rand((n, m), dtype=torch.float64, device="cuda"); C = C0.clone(); torch.cuda.synchronize()
    f77(lib, "dgemm_", "N", tb, m, n, k, -1.0, A, m, Bm, n if tb == "T" else k, 1.0, C, m); torch.cuda.synchronize()
    ref = C0.T - A.T @ (Bm if tb == "T" else Bm.T)
    if (C.T - ref).abs().max().item() > 1e-9: nbad += 1
print(sys.argv[1:], g.last_variant(), "bad iterations:", nbad, "of 100")
