"""SGEMM accuracy/perf study: 3xTF32 tcgen05 kernel vs the FFMA tile kernel vs cuBLAS fp32 (torch, TF32 off)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import libgpublas_b200 as g
g.load(); g.use_torch_stream(); g.set_sync(False)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
EPS = 2.0 ** -24

def run(variant, A, B, C, m, n, k):
    g.force_variant(variant)
    g.call("sgemm_", "T", "N", m, n, k, 1.0, A, k, B, k, 0.0, C, m)   # A: (m,k) row-major == k x m col-major, op = T
    g.force_variant("auto")

def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

if "acc" in sys.argv:
    m = n = 256
    for k in (64, 512, 2048, 8192, 16384):
        gen = torch.Generator(device="cuda").manual_seed(k)
        A = torch.rand((m, k), device="cuda", generator=gen) * 2 - 1      # row-major (m,k)
        B = torch.rand((n, k), device="cuda", generator=gen) * 2 - 1      # row-major (n,k) == k x n col-major
        ref = A.double() @ B.double().T                                   # (m,n)
        gb = A.double().abs() @ B.double().abs().T
        out = {}
        for v in ("tf32x3_tcgen05", "generic_tile"):
            C = torch.zeros((n, m), device="cuda")                        # col-major m x n == row-major (n,m)
            run(v, A, B, C, m, n, k); torch.cuda.synchronize()
            err = (C.T.double() - ref).abs()
            out[v] = ((err / (EPS * gb)).max().item(), (err.norm() / ref.norm()).item())
        Cc = A @ B.T
        err = (Cc.double() - ref).abs()
        out["cublas_fp32"] = ((err / (EPS * gb)).max().item(), (err.norm() / ref.norm()).item())
        print("k=%6d " % k + "  ".join("%s: ratio %.2f relF %.2e" % (v, a, b) for v, (a, b) in out.items()), flush=True)

if "perf" in sys.argv:
    for n in (2048, 4096, 8192, 16384):
        A = torch.rand((n, n), device="cuda") * 2 - 1; B = torch.rand((n, n), device="cuda") * 2 - 1; C = torch.zeros((n, n), device="cuda")
        for ta, tb in (("N", "N"), ("T", "N"), ("N", "T")):
            ms = t(lambda: g.call("sgemm_", ta, tb, n, n, n, 1.0, A, n, B, n, 0.0, C, n))
            print(f"sgemm {ta}{tb} n={n}: {ms:.3f} ms {2.0*n**3/ms/1e9:.1f} TFLOP/s {g.last_variant()}", flush=True)
        ms = t(lambda: torch.matmul(A, B, out=C))
        print(f"cuBLAS fp32 n={n}: {ms:.3f} ms {2.0*n**3/ms/1e9:.1f} TFLOP/s", flush=True)
