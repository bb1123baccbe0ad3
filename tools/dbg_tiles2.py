import torch, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import libgpublas_b200 as g
from helpers import f77
lib = g.load()
torch.manual_seed(1)
m, n, k = 2048, 256, 256
found = 0
for it in range(200):
    A = torch.rand((k, m), dtype=torch.float64, device="cuda")
    Bm = torch.rand((k, n), dtype=torch.float64, device="cuda")      # 'T': stored n x k col-major
    C0 = torch.rand((n, m), dtype=torch.float64, device="cuda"); C = C0.clone(); torch.cuda.synchronize()
    f77(lib, "dgemm_", "N", "T", m, n, k, -1.0, A, m, Bm, n, 1.0, C, m); torch.cuda.synchronize()
    ref = C0.T - A.T @ Bm
    err = (C.T - ref).abs()
    if err.max().item() > 1e-9:
        bad = (err > 1e-9).nonzero()
        r, c = bad[0].tolist()
        got = C.T[r, c].item(); c0 = C0.T[r, c].item(); rf = ref[r, c].item()
        # partial sums over 16-wide k stages
        a = A.T[r]; b = Bm[:, c]
        parts = (a * b).view(16, 16).sum(1)
        cum = c0 - torch.cumsum(parts, 0)
        diffs = (cum - got).abs()
        # which single stage missing?
        miss = [(c0 - (parts.sum() - parts[s])).item() for s in range(16)]
        best = min(range(16), key=lambda s: abs(miss[s] - got))
        print("it", it, "bad", bad.shape[0], "at", (r, c), "got", got, "ref", rf, "c0", c0, "| closest prefix stage", int(diffs.argmin()), float(diffs.min()),
              "| missing one stage", best, abs(miss[best] - got), flush=True)
        found += 1
        if found >= 6: break
print("found", found)
