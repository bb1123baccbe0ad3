"""Partitioned dsyrk_/dtrsm_/dtrmm_ behind the symbol: time on 1 and N devices (tools only).  usage: ml3_perf.py <ndev> [n]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import libgpublas_b200 as g

ndev = int(sys.argv[1]); n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
lib = g.load(); g.use_torch_stream(); g.set_sync(False)
torch.cuda.set_device(0)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


A = torch.rand((n, n), dtype=torch.float64, device="cuda") * 2 - 1
C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
T = torch.triu(A).contiguous(); T.mul_(1.0 / n); T.diagonal().fill_(1.0)
base = {}
for nd in (1, ndev):
    lib.b200blas_set_options(("devices=%d" % nd).encode())
    for name, fn, flops in (("dsyrk LN", lambda: g.call("dsyrk_", "L", "N", n, n, 1.0, A, n, 0.0, C, n), float(n) ** 3),
                            ("dsyrk UT", lambda: g.call("dsyrk_", "U", "T", n, n, 1.0, A, n, 0.5, C, n), float(n) ** 3),
                            ("dtrsm LLNN", lambda: g.call("dtrsm_", "L", "L", "N", "N", n, n, 1.0, T, n, C, n), float(n) ** 3),
                            ("dtrsm RLTN", lambda: g.call("dtrsm_", "R", "L", "T", "N", n, n, 1.0, T, n, C, n), float(n) ** 3),
                            ("dtrmm LLNN", lambda: g.call("dtrmm_", "L", "L", "N", "N", n, n, 1.0, T, n, C, n), float(n) ** 3)):
        C.uniform_(-1, 1)
        ms = timed(fn)
        base.setdefault(name, ms)
        print("%s n=%d devices=%d: %.2f ms  %.1f TFLOP/s  speed-up %.2fx" % (name, n, nd, ms, flops / ms / 1e9, base[name] / ms), flush=True)
