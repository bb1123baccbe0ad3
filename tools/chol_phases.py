"""Per-phase timing of the blocked Cholesky workload (potrf / trsm / syrk), CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g
from libgpublas_b200 import DevPtr, call
from libgpublas_b200.cholesky import potrf_lower
g.load(); g.use_torch_stream()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
A = torch.rand((n, n), dtype=torch.float64, device="cuda") * 2 - 1
A = torch.tril(A, -1); A = A + A.T; A.diagonal().fill_(float(n))
base = A.data_ptr(); at = lambda i, j: DevPtr(base + 8 * (i + j * n))
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
tot = {"potrf": 0.0, "trsm": 0.0, "syrk": 0.0}; fl = {"potrf": 0.0, "trsm": 0.0, "syrk": 0.0}
rows = []
for j in range(0, n, nb):
    jb = min(nb, n - j); rest = n - j - jb
    e0 = ev(); potrf_lower(jb, at(j, j), n); e1 = ev()
    if rest > 0:
        call("dtrsm_", "R", "L", "T", "N", rest, jb, 1.0, at(j, j), n, at(j + jb, j), n)
    e2 = ev()
    if rest > 0:
        call("dsyrk_", "L", "N", rest, jb, -1.0, at(j + jb, j), n, 1.0, at(j + jb, j + jb), n)
    e3 = ev(); torch.cuda.synchronize()
    t = (e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3))
    f = (jb ** 3 / 3, rest * jb * jb, rest * (rest + 1) * jb)
    for k, tt, ff in zip(("potrf", "trsm", "syrk"), t, f):
        tot[k] += tt; fl[k] += ff
    rows.append((j, t, tuple(ff / max(tt, 1e-9) / 1e9 for tt, ff in zip(t, f))))
for j, t, r in rows:
    print("j=%6d potrf %.2f ms (%.1f TF)  trsm %.2f ms (%.1f TF)  syrk %.2f ms (%.1f TF)" % (j, t[0], r[0], t[1], r[1], t[2], r[2]))
print({k: (round(tot[k], 1), round(fl[k] / max(tot[k], 1e-9) / 1e9, 2)) for k in tot}, "total ms", round(sum(tot.values()), 1))
