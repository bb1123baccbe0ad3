"""sgemm_/zgemm_/cgemm_ behind the symbol on 1 and N devices (bulk types, k consumed in chunks).  usage: mg_bulk_perf.py <ndev>"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import libgpublas_b200 as g

ndev = int(sys.argv[1])
lib = g.load(); g.use_torch_stream(); g.set_sync(False)
torch.cuda.set_device(0)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


base = {}
for name, dt, n, alpha, beta, fl in (("sgemm_", torch.float32, 16384, 1.0, 0.0, 2.0), ("zgemm_", torch.complex128, 8192, 0.7 - 0.9j, 1.3 - 1.1j, 8.0),
                                     ("cgemm_", torch.complex64, 8192, 0.7 - 0.9j, 0.0j, 8.0)):
    A = torch.rand((n, n), dtype=dt, device="cuda"); B = torch.rand((n, n), dtype=dt, device="cuda"); C = torch.zeros((n, n), dtype=dt, device="cuda")
    for nd in (1, ndev):
        lib.b200blas_set_options(("devices=%d" % nd).encode())
        ms = timed(lambda: g.call(name, "N", "N", n, n, n, alpha, A, n, B, n, beta, C, n))
        base.setdefault(name, ms)
        print("%s n=%d devices=%d: %.2f ms  %.1f TFLOP/s  speed-up %.2fx" % (name, n, nd, ms, fl * n ** 3 / ms / 1e9, base[name] / ms), flush=True)
    del A, B, C
