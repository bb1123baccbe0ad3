"""Partitioned Level-3 over the GPUs of one box for every precision and for SYRK / TRSM, in
"bulk" mode: the home GPU's copy engines push each rank's operand panels into that rank's memory over NVLink (CUDA
IPC), a flag written behind the data releases the rank's stream (a one-thread polling kernel), the rank runs the
ordinary single-GPU routine of this library on its panels, pushes its result tile back into the home allocation with
its own copy engine and raises a "done" flag the home stream waits on.  No NCCL and no host synchronisation on the
data path; torch.distributed is used once, to exchange the IPC handles.

(The DGEMM path of multigpu.TiledGemm goes further -- flag-polling inside the kernel and epilogue stores to the
home allocation -- because there the transfer can hide behind 30 ms of FP64 work.  A 16384^3 SGEMM takes 38 ms on
ONE GPU, less than pushing its operands to seven peers, so for the fast precisions the push is the critical path
whatever the kernel does; bulk mode keeps the code routine-agnostic.)

    PartitionedGemm(p, m, n, k, ...)     C := alpha*A*B + beta*C   p in 's','d','c','z', 2-D tiles like TiledGemm
    PartitionedSyrk(n, k, ...)           C := alpha*A*A^T + beta*C (lower, f64): column strips of equal area
    PartitionedTrsm(m, n, ...)           B := alpha*B*L^-T        (right, lower, transposed, f64): row slices of B
"""
import ctypes
import math

import torch
import torch.distributed as dist

from . import DevPtr, call, load
from .multigpu import block_range, grid_for

_ES = {"s": 4, "d": 8, "c": 8, "z": 16}
_DT = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}


def _lib():
    lib = load()
    if not getattr(lib, "_b200_partitioned_ready", False):
        lib.b200blas_device_malloc.restype = ctypes.c_void_p
        lib.b200blas_device_malloc.argtypes = [ctypes.c_size_t]
        lib.b200blas_ipc_open.restype = ctypes.c_void_p
        lib.b200blas_ipc_open.argtypes = [ctypes.c_void_p]
        lib.b200blas_ipc_get_handle.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.b200blas_copy2d_async.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
        lib.b200blas_copy2d_async.restype = None
        lib.b200blas_write_flag_async.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
        lib.b200blas_write_flag_async.restype = None
        lib.b200blas_wait_flag_async.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
        lib.b200blas_wait_flag_async.restype = None
        lib.b200blas_memset_async.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p]
        lib.b200blas_memset_async.restype = None
        lib._b200_partitioned_ready = True
    return lib


class Exchange:
    """IPC plumbing shared by the drivers below.  Rank 0 is the home rank.

    peers' memory:  `inbuf` (operand panels, written by home) and flags[0] = "inputs of epoch e have landed"
    home's memory:  `out` (the result matrix, written by every rank) and flags[r] = "rank r's tile of epoch e has landed"
    """

    def __init__(self, dev, rank, world, in_bytes, out_bytes):
        self.dev, self.rank, self.world = dev, rank, world
        self.lib = lib = _lib()
        self.epoch = 0

        def export(ptr):
            buf = ctypes.create_string_buffer(64)
            assert lib.b200blas_ipc_get_handle(ctypes.c_void_p(ptr), buf) == 64
            return bytes(buf.raw)

        def alloc(nbytes):
            p = lib.b200blas_device_malloc(max(256, nbytes))
            assert p, "device allocation failed"
            return p

        self.flags = alloc(4 * 64)
        lib.b200blas_memset_async(ctypes.c_void_p(self.flags), 0, 4 * 64, None)
        self.inbuf = alloc(in_bytes) if rank != 0 else None
        self.out = alloc(out_bytes) if rank == 0 else None
        torch.cuda.synchronize()
        mine = (export(self.flags), export(self.inbuf) if rank != 0 else None, export(self.out) if rank == 0 else None)
        allh = [None] * world
        dist.all_gather_object(allh, mine)

        def opn(h):
            p = lib.b200blas_ipc_open(ctypes.create_string_buffer(h, 64))
            if not p:
                raise RuntimeError("cudaIpcOpenMemHandle failed: peer access between the GPUs is required")
            return p

        if rank == 0:
            self.peer_flags = {r: opn(allh[r][0]) for r in range(1, world)}
            self.peer_in = {r: opn(allh[r][1]) for r in range(1, world)}
            self.push_streams = {r: torch.cuda.Stream(device=dev) for r in range(1, world)}
        else:
            self.home_flags = opn(allh[0][0])
            self.out = opn(allh[0][2])
        dist.barrier()

    # ---- home side ----
    def begin(self):
        self.epoch += 1
        assert self.epoch < 65535
        if self.rank == 0:
            ready = torch.cuda.Event(); ready.record(torch.cuda.current_stream(self.dev))
            for st in self.push_streams.values():
                st.wait_event(ready)

    def push(self, r, dst_off, dpitch, src_ptr, spitch, width, height):
        if width > 0 and height > 0:
            self.lib.b200blas_copy2d_async(ctypes.c_void_p(self.peer_in[r] + dst_off), dpitch, ctypes.c_void_p(src_ptr), spitch, width, height,
                                           ctypes.c_void_p(self.push_streams[r].cuda_stream))

    def release(self, r):
        self.lib.b200blas_write_flag_async(ctypes.c_void_p(self.peer_flags[r]), self.epoch, ctypes.c_void_p(self.push_streams[r].cuda_stream))

    def home_wait_results(self):
        comp = torch.cuda.current_stream(self.dev)
        for r in range(1, self.world):
            self.lib.b200blas_wait_flag_async(ctypes.c_void_p(self.flags + 4 * r), self.epoch, ctypes.c_void_p(comp.cuda_stream))
        for st in self.push_streams.values():
            comp.wait_stream(st)

    # ---- peer side ----
    def wait_inputs(self):
        comp = torch.cuda.current_stream(self.dev)
        self.lib.b200blas_wait_flag_async(ctypes.c_void_p(self.flags), self.epoch, ctypes.c_void_p(comp.cuda_stream))

    def give_back(self, dst_off, dpitch, src_ptr, spitch, width, height):
        comp = ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        if width > 0 and height > 0:
            self.lib.b200blas_copy2d_async(ctypes.c_void_p(self.out + dst_off), dpitch, ctypes.c_void_p(src_ptr), spitch, width, height, comp)
        self.lib.b200blas_write_flag_async(ctypes.c_void_p(self.home_flags + 4 * self.rank), self.epoch, comp)


class PartitionedGemm:
    """C(m x n) := alpha*A*B + beta*C for any precision, operands and result on rank 0 (column-major, ld = rows)."""

    def __init__(self, p, m, n, k, device, rank, world):
        self.p, self.m, self.n, self.k, self.dev, self.rank, self.world = p, m, n, k, device, rank, world
        self.es = _ES[p]
        self.P, self.Q = grid_for(world)
        pg, qg = rank // self.Q, rank % self.Q
        self.r0, self.r1 = block_range(m, self.P, pg)
        self.c0, self.c1 = block_range(n, self.Q, qg)
        tm, tn = self.r1 - self.r0, self.c1 - self.c0
        self.ex = Exchange(device, rank, world, (tm * k + k * tn) * self.es, m * n * self.es)
        if rank != 0:
            self.ctile = torch.empty(max(1, tm * tn), dtype=_DT[p], device=device)

    def run(self, A=None, B=None, alpha=1.0, beta=0.0):
        """A, B: torch tensors on rank 0 (1-D column-major buffers).  beta must be 0 (the result is produced, not updated)."""
        assert beta == 0.0
        ex, es, m, n, k = self.ex, self.es, self.m, self.n, self.k
        one = (1.0 + 0j) if self.p in "cz" else 1.0
        zero = 0 * one
        ex.begin()
        if self.rank == 0:
            a0, b0 = A.data_ptr(), B.data_ptr()
            for r in range(1, self.world):                     # A panels first, round-robin over the peers
                pg, qg = r // self.Q, r % self.Q
                r0, r1 = block_range(m, self.P, pg)
                ex.push(r, 0, (r1 - r0) * es, a0 + es * r0, m * es, (r1 - r0) * es, k)
            for r in range(1, self.world):
                pg, qg = r // self.Q, r % self.Q
                r0, r1 = block_range(m, self.P, pg); c0, c1 = block_range(n, self.Q, qg)
                ex.push(r, (r1 - r0) * k * es, k * es, b0 + es * c0 * k, k * es, k * es, c1 - c0)
                ex.release(r)
            tm, tn = self.r1 - self.r0, self.c1 - self.c0
            if tm > 0 and tn > 0:
                call(self.p + "gemm_", "N", "N", tm, tn, k, alpha * one, DevPtr(a0 + es * self.r0), m, DevPtr(b0 + es * self.c0 * k), k, zero,
                     DevPtr(ex.out + es * (self.r0 + self.c0 * m)), m)
            ex.home_wait_results()
        else:
            tm, tn = self.r1 - self.r0, self.c1 - self.c0
            ex.wait_inputs()
            if tm > 0 and tn > 0:
                call(self.p + "gemm_", "N", "N", tm, tn, k, alpha * one, DevPtr(ex.inbuf), tm, DevPtr(ex.inbuf + tm * k * es), k, zero, self.ctile, tm)
            ex.give_back(es * (self.r0 + self.c0 * m), m * es, self.ctile.data_ptr(), max(1, tm) * es, tm * es, tn)

    def result_ptr(self):
        return self.ex.out


def strip_bounds(n, parts, align=128):
    """Column strips of a lower-triangular n x n update with (nearly) equal areas: strip i covers columns [b[i], b[i+1])
    and rows b[i]..n, area ~ (n - c0)^2 - (n - c1)^2.  Boundaries rounded to `align`."""
    b = [0]
    for i in range(1, parts):
        c = n * (1.0 - math.sqrt(1.0 - i / parts))
        c = int(round(c / align)) * align
        b.append(min(n, max(b[-1], c)))
    b.append(n)
    return b


class PartitionedSyrk:
    """C := alpha*A*A^T + beta*C on the lower triangle (f64, trans 'N'); A (n x k, ld n) on rank 0, C (n x n, ld n) in the
    home allocation `c_ptr()`.  Rank r owns the column strip [b[r], b[r+1]) of the triangle (equal areas): it receives rows
    b[r]..n of A and its strip of C, runs dsyrk_ on the diagonal block and dgemm_('N','T') on the block below it, and pushes
    the strip back (the strictly upper part of the diagonal block travels both ways unchanged, so nothing outside the
    referenced triangle is modified at home)."""

    def __init__(self, n, k, device, rank, world):
        self.n, self.k, self.dev, self.rank, self.world = n, k, device, rank, world
        self.b = strip_bounds(n, world)
        c0, c1 = self.b[rank], self.b[rank + 1]
        self.rows, self.w = n - c0, c1 - c0
        self.ex = Exchange(device, rank, world, (self.rows * k + self.rows * self.w) * 8, n * n * 8)

    def c_ptr(self):
        return self.ex.out

    def run(self, A=None, alpha=1.0, beta=0.0):
        ex, n, k = self.ex, self.n, self.k
        ex.begin()
        if self.rank == 0:
            a0, cptr = A.data_ptr(), ex.out
            for r in range(1, self.world):
                rc0, rc1 = self.b[r], self.b[r + 1]
                rows, w = n - rc0, rc1 - rc0
                ex.push(r, 0, rows * 8, a0 + 8 * rc0, n * 8, rows * 8, k)                             # A[rc0:n, :]
                ex.push(r, rows * k * 8, rows * 8, cptr + 8 * (rc0 + rc0 * n), n * 8, rows * 8, w)     # C[rc0:n, rc0:rc1]
                ex.release(r)
            w, rows = self.w, self.rows
            if w > 0:
                call("dsyrk_", "L", "N", w, k, alpha, DevPtr(a0), n, beta, DevPtr(cptr), n)
                if rows - w > 0:
                    call("dgemm_", "N", "T", rows - w, w, k, alpha, DevPtr(a0 + 8 * w), n, DevPtr(a0), n, beta, DevPtr(cptr + 8 * w), n)
            ex.home_wait_results()
        else:
            rows, w, c0 = self.rows, self.w, self.b[self.rank]
            ex.wait_inputs()
            a, cs = ex.inbuf, ex.inbuf + rows * k * 8
            if w > 0:
                call("dsyrk_", "L", "N", w, k, alpha, DevPtr(a), rows, beta, DevPtr(cs), rows)
                if rows - w > 0:
                    call("dgemm_", "N", "T", rows - w, w, k, alpha, DevPtr(a + 8 * w), rows, DevPtr(a), rows, beta, DevPtr(cs + 8 * w), rows)
            ex.give_back(8 * (c0 + c0 * n), n * 8, cs, rows * 8, rows * 8, w)


class PartitionedTrsm:
    """B := alpha * B * L^-T (side 'R', lower, transposed, non-unit: the panel solve of a blocked Cholesky), f64.
    L (n x n, ld n) on rank 0, B (m x n, ld m) in the home allocation `b_ptr()`.  The rows of B are independent: rank r
    receives L and its row slice, solves with dtrsm_, pushes the slice back; no exchange between the ranks."""

    def __init__(self, m, n, device, rank, world):
        self.m, self.n, self.dev, self.rank, self.world = m, n, device, rank, world
        self.r0, self.r1 = block_range(m, world, rank)
        self.ex = Exchange(device, rank, world, (n * n + (self.r1 - self.r0) * n) * 8, m * n * 8)

    def b_ptr(self):
        return self.ex.out

    def run(self, L=None, alpha=1.0):
        ex, m, n = self.ex, self.m, self.n
        ex.begin()
        if self.rank == 0:
            l0, bptr = L.data_ptr(), ex.out
            for r in range(1, self.world):
                r0, r1 = block_range(m, self.world, r)
                ex.push(r, 0, n * 8, l0, n * 8, n * 8, n)
                ex.push(r, n * n * 8, (r1 - r0) * 8, bptr + 8 * r0, m * 8, (r1 - r0) * 8, n)
                ex.release(r)
            mr = self.r1 - self.r0
            if mr > 0:
                call("dtrsm_", "R", "L", "T", "N", mr, n, alpha, DevPtr(l0), n, DevPtr(bptr + 8 * self.r0), m)
            ex.home_wait_results()
        else:
            mr = self.r1 - self.r0
            ex.wait_inputs()
            if mr > 0:
                call("dtrsm_", "R", "L", "T", "N", mr, n, alpha, DevPtr(ex.inbuf), n, DevPtr(ex.inbuf + n * n * 8), mr)
            ex.give_back(8 * self.r0, m * 8, ex.inbuf + n * n * 8, max(1, mr) * 8, mr * 8, n)
