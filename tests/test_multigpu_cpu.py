"""Host logic of the 2-D tile-partitioned GEMM (libgpublas_b200/multigpu.py) on CPU: world_size 2 and 4 over
gloo, the per-tile GEMM replaced by the oracle's ref_dgemm on raw pointers.  Checks grid choice, row/column
block ranges, the k-chunk broadcast schedule (ragged last chunk, double buffering) and the tile gather."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import load_oracle  # noqa: E402
from libgpublas_b200.multigpu import TiledGemm, block_range, grid_for  # noqa: E402


def test_grid_and_blocks():
    assert grid_for(1) == (1, 1) and grid_for(2) == (1, 2) and grid_for(4) == (2, 2) and grid_for(8) == (2, 4)
    for total, parts in [(16384, 2), (16384, 4), (300, 2), (5, 4), (1000, 3), (129, 2)]:
        rs = [block_range(total, parts, i) for i in range(parts)]
        assert rs[0][0] == 0 and rs[-1][1] == total
        assert all(rs[i][1] == rs[i + 1][0] for i in range(parts - 1)) and all(lo <= hi for lo, hi in rs)
    assert block_range(16384, 4, 1) == (4096, 8192)


def _oracle_gemm(m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc):
    lib = load_oracle()
    lib.ref_dgemm.restype = ctypes.c_int
    rc = lib.ref_dgemm(ctypes.c_char(b"N"), ctypes.c_char(b"N"), m, n, k, ctypes.c_double(alpha), ctypes.c_void_p(a_ptr), lda,
                       ctypes.c_void_p(b_ptr), ldb, ctypes.c_double(beta), ctypes.c_void_p(c_ptr), ldc)
    assert rc == 0


def _worker(rank, world, port, m, n, k, kchunk, q):
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        tg = TiledGemm(m, n, k, torch.device("cpu"), rank, world, kchunk=kchunk, gemm=_oracle_gemm, c_return="sendrecv")
        A = B = None
        if rank == 0:
            gen = torch.Generator().manual_seed(3)
            A = torch.rand(m * k, dtype=torch.float64, generator=gen) * 2 - 1
            B = torch.rand(k * n, dtype=torch.float64, generator=gen) * 2 - 1
        tg.set_inputs(A, B)
        for _ in range(2):      # twice: buffers and events must be reusable
            tg.run()
        if rank == 0:
            a = A.numpy().reshape((m, k), order="F"); b = B.numpy().reshape((k, n), order="F")
            c = tg.home_c().numpy().reshape((m, n), order="F")
            err = np.abs(c - a @ b).max()
            q.put(("ok", float(err), tg.describe()))
    except Exception as e:   # surface the failure to the parent
        q.put(("fail", repr(e), ""))
        raise
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world,shape", [(2, (300, 260, 500)), (4, (257, 300, 130)), (2, (64, 5, 33))])
def test_tiled_gemm_gloo(world, shape):
    m, n, k = shape
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, k, 128, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, err, desc = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", err
    assert err < 1e-10, err
    assert "2d-tile" in desc


# ---------------------------------------------------------------------------------------------
# Blocked Cholesky (BASELINE.json configs[3]) over ranks: ownership, broadcast order and look-ahead ordering of
# libgpublas_b200/cholesky.py::TiledCholesky with the four block operations replaced by numpy on the shared buffer.
class _NumpyBlas:
    def __init__(self, A, n):
        self.M = A.numpy().reshape((n, n), order="F")        # shares memory with the torch buffer

    def potrf(self, j, jb):
        blk = self.M[j:j + jb, j:j + jb]
        try:
            L = np.linalg.cholesky(np.tril(blk) + np.tril(blk, -1).T)
        except np.linalg.LinAlgError:
            return 1
        blk[np.tril_indices(jb)] = L[np.tril_indices(jb)]
        return 0

    def trsm(self, j, jb, rest):
        L = np.tril(self.M[j:j + jb, j:j + jb])
        P = self.M[j + jb:j + jb + rest, j:j + jb]
        P[:] = np.linalg.solve(L, P.T).T

    def syrk(self, kcol, kb, j, jb):
        P = self.M[kcol:kcol + kb, j:j + jb]
        blk = self.M[kcol:kcol + kb, kcol:kcol + kb]
        upd = P @ P.T
        blk[np.tril_indices(kb)] -= upd[np.tril_indices(kb)]

    def gemm(self, kcol, kb, below, j, jb):
        self.M[kcol + kb:kcol + kb + below, kcol:kcol + kb] -= self.M[kcol + kb:kcol + kb + below, j:j + jb] @ self.M[kcol:kcol + kb, j:j + jb].T


def _chol_worker(rank, world, port, n, nb, q):
    from libgpublas_b200.cholesky import TiledCholesky
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        tc = TiledCholesky(n, nb, torch.device("cpu"), rank, world, blas=_NumpyBlas)
        gen = torch.Generator().manual_seed(9)
        G = torch.rand((n, n), dtype=torch.float64, generator=gen) * 2 - 1
        S = torch.tril(G, -1); S = S + S.T + n * torch.eye(n, dtype=torch.float64)
        A0 = S.numpy().copy()
        buf = S.T.contiguous().view(-1) if rank == 0 else torch.zeros(n * n, dtype=torch.float64)   # only rank 0 holds the input
        tc.set_matrix(buf)
        info = tc.run()
        L = np.tril(buf.numpy().reshape((n, n), order="F"))
        err = np.linalg.norm(L @ L.T - A0) / np.linalg.norm(A0)
        q.put(("ok", rank, info, float(err)))
    except Exception as e:
        q.put(("fail", rank, repr(e), 0.0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,nb", [(2, 300, 64), (3, 257, 50), (2, 64, 100)])
def test_tiled_cholesky_gloo(world, n, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chol_worker, args=(r, world, port, n, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for status, rank, info, err in res:      # every rank ends with the complete factor
        assert status == "ok", info
        assert info == 0 and err < 1e-13, (rank, info, err)


def test_syrk_strip_bounds_equal_area():
    """partitioned.PartitionedSyrk: column strips of the lower triangle with (nearly) equal areas, 128-aligned, covering n."""
    from libgpublas_b200.partitioned import strip_bounds
    for n, parts in [(32768, 8), (32768, 2), (16384, 4), (1000, 3), (128, 8)]:
        b = strip_bounds(n, parts)
        assert b[0] == 0 and b[-1] == n and len(b) == parts + 1 and all(b[i] <= b[i + 1] for i in range(parts))
        assert all(x % 128 == 0 for x in b[1:-1])
        if n >= 128 * parts * 4:
            areas = [(n - b[i]) ** 2 - (n - b[i + 1]) ** 2 for i in range(parts)]
            assert max(areas) <= 1.25 * (n * n / parts), (n, parts, b, areas)
