"""Timeline (B200BLAS_MG_TRACE) of one partitioned sgemm_ / cgemm_ call.  usage: mg_bulk_trace.py <ndev> <s|c|z>"""
import os, sys
sys.path.insert(0, os.getcwd())
os.environ["B200BLAS_MG_TRACE"] = "1"
import torch
import libgpublas_b200 as g
ndev = int(sys.argv[1]); which = sys.argv[2]
lib = g.load(); g.use_torch_stream(); g.set_sync(False)
torch.cuda.set_device(0)
name, dt, n = {"s": ("sgemm_", torch.float32, 16384), "c": ("cgemm_", torch.complex64, 8192), "z": ("zgemm_", torch.complex128, 8192)}[which]
A = torch.rand((n, n), dtype=dt, device="cuda"); B = torch.rand((n, n), dtype=dt, device="cuda"); C = torch.zeros((n, n), dtype=dt, device="cuda")
lib.b200blas_set_options(("devices=%d" % ndev).encode())
for it in range(3):
    sys.stderr.write("---- call %d\n" % it); sys.stderr.flush()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); g.call(name, "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n); e1.record(); torch.cuda.synchronize()
    sys.stderr.write("call %d: %.3f ms\n" % (it, e0.elapsed_time(e1)))
