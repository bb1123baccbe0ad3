"""Partitioned SGEMM/ZGEMM/DSYRK/DTRSM over N GPUs (torchrun): correctness against the 1-GPU routine on rank 0 and timing
(max over ranks, CUDA events, operands and result on rank 0)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import libgpublas_b200 as g
from libgpublas_b200 import DevPtr, call
from libgpublas_b200.partitioned import PartitionedGemm, PartitionedSyrk, PartitionedTrsm

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g.load(); g.use_torch_stream(); g.set_sync(False)
which = sys.argv[1:] or ["sgemm", "zgemm", "dsyrk", "dtrsm"]
steps = 3


def timed(fn):
    fn(); torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(steps):
        torch.cuda.synchronize(); dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(t.item())
    return sorted(ts)[len(ts) // 2]


def raw_view(ptr, numel, dtype):
    """copy of `numel` elements at device address ptr into a torch tensor (plumbing for the checks)"""
    out = torch.empty(numel, dtype=dtype, device=dev)
    es = out.element_size()
    import ctypes
    g.load().b200blas_copy2d_async(ctypes.c_void_p(out.data_ptr()), numel * es, ctypes.c_void_p(ptr), numel * es, numel * es, 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out


for name in which:
    if name in ("sgemm", "zgemm"):
        p = name[0]; n = 16384 if p == "s" else 8192
        dt = torch.float32 if p == "s" else torch.complex128
        pg = PartitionedGemm(p, n, n, n, dev, rank, world)
        A = B = None
        if rank == 0:
            A = torch.rand(n * n, dtype=dt, device=dev); B = torch.rand(n * n, dtype=dt, device=dev)
        ms = timed(lambda: pg.run(A, B))
        if rank == 0:
            ref = torch.empty(n * n, dtype=dt, device=dev)
            one, zero = ((1.0 + 0j), 0j) if p == "z" else (1.0, 0.0)
            call(p + "gemm_", "N", "N", n, n, n, one, A, n, B, n, zero, ref, n); torch.cuda.synchronize()
            got = raw_view(pg.result_ptr(), n * n, dt)
            flops = (2.0 if p == "s" else 8.0) * n ** 3
            print(json.dumps({"routine": name, "n": n, "n_gpus": world, "ms": ms, "tflops": flops / ms / 1e9,
                              "max_abs_diff_vs_1gpu": float((got - ref).abs().max())}), flush=True)
            del ref, got
        del pg, A, B
    elif name == "dsyrk":
        n, k = 32768, 2048
        ps = PartitionedSyrk(n, k, dev, rank, world)
        A = None
        if rank == 0:
            A = torch.rand(n * k, dtype=torch.float64, device=dev) * 2 - 1
            C0 = torch.rand(n * n, dtype=torch.float64, device=dev)
            import ctypes
            lib = g.load()
            def reset():
                lib.b200blas_copy2d_async(ctypes.c_void_p(ps.c_ptr()), n * n * 8, ctypes.c_void_p(C0.data_ptr()), n * n * 8, n * n * 8, 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            reset()
        ms = timed(lambda: ps.run(A, alpha=-1.0, beta=1.0))
        if rank == 0:
            reset(); torch.cuda.synchronize()
        dist.barrier(); ps.run(A, alpha=-1.0, beta=1.0); torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            ref = C0.clone(); call("dsyrk_", "L", "N", n, k, -1.0, A, n, 1.0, ref, n); torch.cuda.synchronize()
            got = raw_view(ps.c_ptr(), n * n, torch.float64)
            print(json.dumps({"routine": "dsyrk L,N", "n": n, "k": k, "n_gpus": world, "ms": ms, "tflops": k * n * (n + 1.0) / ms / 1e9,
                              "max_abs_diff_vs_1gpu": float((got - ref).abs().max()), "strips": ps.b}), flush=True)
            del ref, got, C0
        del ps, A
    elif name == "dtrsm":
        m, n = 30720, 2048
        pt = PartitionedTrsm(m, n, dev, rank, world)
        L = None
        if rank == 0:
            L = (torch.tril(torch.rand((n, n), dtype=torch.float64, device=dev)) + n * torch.eye(n, dtype=torch.float64, device=dev)).T.contiguous().view(-1)
            B0 = torch.rand(m * n, dtype=torch.float64, device=dev)
            import ctypes
            lib = g.load()
            def resetb():
                lib.b200blas_copy2d_async(ctypes.c_void_p(pt.b_ptr()), m * n * 8, ctypes.c_void_p(B0.data_ptr()), m * n * 8, m * n * 8, 1, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            resetb()
        ms = timed(lambda: pt.run(L))
        if rank == 0:
            resetb(); torch.cuda.synchronize()
        dist.barrier(); pt.run(L); torch.cuda.synchronize(); dist.barrier()
        if rank == 0:
            ref = B0.clone(); call("dtrsm_", "R", "L", "T", "N", m, n, 1.0, L, n, ref, m); torch.cuda.synchronize()
            got = raw_view(pt.b_ptr(), m * n, torch.float64)
            print(json.dumps({"routine": "dtrsm R,L,T,N", "m": m, "n": n, "n_gpus": world, "ms": ms, "tflops": 1.0 * m * n * n / ms / 1e9,
                              "max_abs_diff_vs_1gpu": float((got - ref).abs().max())}), flush=True)
        del pt, L
dist.destroy_process_group()
