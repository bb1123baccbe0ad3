"""Blocked Cholesky n x n over the GPUs of one box (torchrun): time = max over ranks, CUDA events, matrix starts on rank 0.
  python -m torch.distributed.run --nproc-per-node N tools/chol_multigpu.py [n] [nb] [steps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import libgpublas_b200 as g
from libgpublas_b200.cholesky import TiledCholesky, blocked_cholesky

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
g.load(); g.use_torch_stream(); g.set_sync(False)
src = torch.empty(n * n, dtype=torch.float64, device=dev)
if rank == 0:
    gen = torch.Generator(device=dev).manual_seed(9)
    M = torch.rand((n, n), dtype=torch.float64, device=dev, generator=gen) * 2 - 1
    M = torch.tril(M, -1); M = M + M.T; M.diagonal().fill_(float(n))
    src.copy_(M.view(-1)); del M
work = torch.empty(n * n, dtype=torch.float64, device=dev)
tc = TiledCholesky(n, nb, dev, rank, world)
times = []
for it in range(steps + 1):
    if rank == 0:
        work.copy_(src)
    torch.cuda.synchronize(); dist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    tc.set_matrix(work)          # broadcast of the input is part of the timed span (operands start on rank 0)
    info = tc.run()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if it > 0:
        times.append(t.item())
if rank == 0:
    # probe ||A x - L (L^T x)|| / (n eps ||A|| ||x||)
    A0 = src.view(n, n); L = torch.tril(work.view(n, n).T)     # buffer is column-major: view(n,n) is the transpose
    x = torch.rand(n, dtype=torch.float64, device=dev) * 2 - 1
    r = (A0 @ x - L @ (L.T @ x)).norm().item() / (n * 2.0 ** -53 * A0.norm().item() * x.norm().item())
    ms = sorted(times)[len(times) // 2]
    print(json.dumps({"workload": "blocked cholesky (lower) n=%d nb=%d f64, 1-D block-cyclic columns, panel broadcast, look-ahead 1" % (n, nb),
                      "n_gpus": world, "ms": ms, "tflops": n ** 3 / 3 / ms / 1e9, "info": info, "probe_residual_over_n_eps": r}))
dist.destroy_process_group()
