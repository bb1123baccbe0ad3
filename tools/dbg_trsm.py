import torch, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libgpublas_b200 as g
from helpers import f77
lib = g.load()
n, k = 2048, 512
gen = torch.Generator(device="cuda").manual_seed(9)
A = torch.rand((k, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
if "nosyrk" not in sys.argv:
    f77(lib, "dsyrk_", "L", "N", n, k, 1.0, A, n, 0.0, C, n)
L = torch.tril(torch.rand((k, k), dtype=torch.float64, device="cuda", generator=gen)) + k * torch.eye(k, dtype=torch.float64, device="cuda")
Lcm = L.T.contiguous()
B0 = torch.rand((k, n), dtype=torch.float64, device="cuda", generator=gen)
for it in range(3):
    B = B0.clone(); torch.cuda.synchronize()
    f77(lib, "dtrsm_", "R", "L", "T", "N", n, k, 1.0, Lcm, k, B, n)
    torch.cuda.synchronize()
    X = B.T; back = X @ L.T
    err = (back - B0.T).abs()
    bad = (err > 1e-9).nonzero()
    cols = sorted(set(bad[:, 1].tolist())); rows = sorted(set(bad[:, 0].tolist()))
    if bad.shape[0]:
        import collections
        byc = collections.defaultdict(list)
        for r_, c_ in bad.tolist(): byc[c_].append(r_)
        for c_ in sorted(byc)[:6] + sorted(byc)[-2:]: print("   col", c_, "rows", byc[c_][:12], "err", err[byc[c_][0], c_].item())
    print("iter", it, "max err", err.max().item(), "n bad", bad.shape[0], "cols", cols[:4], "..", cols[-4:], "rows", rows[:4], "..", rows[-4:], "stats", g.stats()["calls"])
