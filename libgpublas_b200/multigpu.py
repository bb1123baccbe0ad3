"""2-D tile-partitioned Level-3 across the GPUs of one box (north_star (4)).

One process per GPU (torchrun), `torch.distributed` for the plumbing.  For C := alpha*A*B + beta*C with
the operands resident on the home rank (rank 0), column-major:

  * device grid P x Q (2 -> 1x2, 4 -> 2x2, 8 -> 2x4); rank (p, q) owns C[p-th row block, q-th column block];
  * k is cut into chunks.  For each chunk the home rank broadcasts A[:, chunk] (a contiguous column block)
    and B[chunk, :] (packed to a contiguous row block) on a communication stream -- with NVSwitch/NVLS a
    broadcast costs about one send, and every rank's inbound bytes equal the operand size -- while every
    rank runs the DMMA GEMM kernel on the previous chunk (C_tile += A_chunk[p rows] * B_chunk[:, q cols]);
  * k is never split across ranks, so no reduction is needed and each C element is produced by exactly one
    rank with the same kernel and tile shape as the 1-GPU path;
  * C return, CUDA path: the home rank exports C through CUDA IPC; the last chunk's GEMM on every rank is
    launched with its C pointer inside the peer mapping, so the kernel's epilogue stores the finished tile
    straight into the home allocation over NVLink ("peer_store").  Fallback / CPU-gloo path: tiles are sent
    back with send/recv and unpacked ("sendrecv").

The compute and memory back ends are injected, so the host logic (grid, slicing, chunk schedule, ordering)
is exercised on CPU with gloo (tests/test_multigpu_cpu.py) and on B200s with NCCL (bench.py --gpus N).
"""
import ctypes
import math

import torch
import torch.distributed as dist


def grid_for(world):
    """P x Q with P <= Q, as square as possible (8 -> 2 x 4)."""
    p = int(math.sqrt(world))
    while world % p:
        p -= 1
    return p, world // p


def block_range(total, parts, idx):
    """[lo, hi) of block idx when `total` is split into `parts` nearly equal blocks (multiples of 128 when possible)."""
    base = (total // parts + 127) // 128 * 128 if total >= 128 * parts else (total + parts - 1) // parts
    lo = min(total, base * idx)
    hi = min(total, base * (idx + 1)) if idx < parts - 1 else total
    return lo, hi


class TiledGemm:
    """C(m x n) := alpha * A(m x k) * B(k x n) + beta * C, column-major, operands on rank 0."""

    def __init__(self, m, n, k, device, rank, world, kchunk=2048, dtype=torch.float64, gemm=None, c_return=None):
        self.m, self.n, self.k = m, n, k
        self.dev, self.rank, self.world, self.dtype = device, rank, world, dtype
        self.P, self.Q = grid_for(world)
        self.p, self.q = rank // self.Q, rank % self.Q
        self.r0, self.r1 = block_range(m, self.P, self.p)
        self.c0, self.c1 = block_range(n, self.Q, self.q)
        self.kchunk = min(kchunk, k)
        self.nchunks = (k + self.kchunk - 1) // self.kchunk
        self.cuda = device.type == "cuda"
        self.gemm = gemm or self._lib_gemm
        self.c_return = c_return or ("peer_store" if self.cuda else "sendrecv")
        self.kernels_per_step = self.nchunks
        self._kernel_ms = None
        # column-major matrices are held as 1-D buffers; element (i, j) of an ld-strided matrix is buf[i + j*ld]
        self.A = self.B = self.C = None
        self.peerC = None
        self._home_c_ptr = None
        # double-buffered chunk landing zones on every rank
        self.abuf = [torch.empty(m * self.kchunk, dtype=dtype, device=device) for _ in range(2)]
        self.bbuf = [torch.empty(self.kchunk * n, dtype=dtype, device=device) for _ in range(2)]
        self.ctile = torch.empty((self.r1 - self.r0) * (self.c1 - self.c0), dtype=dtype, device=device)
        if self.cuda:
            self.comm_stream = torch.cuda.Stream(device=device)
            self.ready = [torch.cuda.Event() for _ in range(2)]
            self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.alpha, self.beta = 1.0, 0.0

    # ------------------------------------------------------------------ set-up
    def make_inputs(self, seed=2):
        """Synthetic U(-1,1) operands on the home rank; C lives in memory the library allocated so it can be
        exported through CUDA IPC."""
        if self.rank == 0:
            gen = torch.Generator(device=self.dev).manual_seed(seed)
            self.A = torch.rand(self.m * self.k, dtype=self.dtype, device=self.dev, generator=gen) * 2 - 1
            self.B = torch.rand(self.k * self.n, dtype=self.dtype, device=self.dev, generator=gen) * 2 - 1
        self._alloc_c()

    def set_inputs(self, A, B, C=None):
        if self.rank == 0:
            self.A, self.B = A, B
        self._alloc_c(C)

    def _alloc_c(self, C=None):
        import libgpublas_b200 as g
        nbytes = self.m * self.n * torch.empty((), dtype=self.dtype).element_size()
        if self.c_return == "peer_store":
            lib = g.load()
            lib.b200blas_device_malloc.restype = ctypes.c_void_p
            lib.b200blas_device_malloc.argtypes = [ctypes.c_size_t]
            lib.b200blas_ipc_open.restype = ctypes.c_void_p
            lib.b200blas_ipc_open.argtypes = [ctypes.c_void_p]
            lib.b200blas_ipc_get_handle.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
            handle = [None]
            if self.rank == 0:
                self._home_c_ptr = lib.b200blas_device_malloc(nbytes)
                assert self._home_c_ptr, "device allocation of C failed"
                buf = ctypes.create_string_buffer(64)
                assert lib.b200blas_ipc_get_handle(ctypes.c_void_p(self._home_c_ptr), buf) == 64
                handle = [bytes(buf.raw)]
            dist.broadcast_object_list(handle, src=0)
            if self.rank == 0:
                self.peerC = self._home_c_ptr
            else:
                self.peerC = lib.b200blas_ipc_open(ctypes.create_string_buffer(handle[0], 64))
                if not self.peerC:
                    raise RuntimeError("cudaIpcOpenMemHandle failed: peer access to the home GPU is required")
        elif self.rank == 0:
            self.C = C if C is not None else torch.zeros(self.m * self.n, dtype=self.dtype, device=self.dev)

    def home_c(self):
        """The result on rank 0 as a (n, m) row-major == (m, n) column-major torch view (CUDA peer_store path: a
        copy out of the library-owned allocation)."""
        assert self.rank == 0
        if self.c_return == "peer_store":
            out = torch.empty(self.m * self.n, dtype=self.dtype, device=self.dev)
            import libgpublas_b200 as g
            g.call("dcopy_", self.m * self.n, g.DevPtr(self._home_c_ptr), 1, out, 1) if self.m * self.n < 2 ** 31 else \
                [g.call("dcopy_", min(2 ** 30, self.m * self.n - o), g.DevPtr(self._home_c_ptr + 8 * o), 1, g.DevPtr(out.data_ptr() + 8 * o), 1)
                 for o in range(0, self.m * self.n, 2 ** 30)]
            torch.cuda.synchronize()
            return out
        return self.C

    def describe(self):
        return "2d-tile %dx%d, k-chunk %d, nccl broadcast of A/B chunks overlapped with compute, C via %s" % (
            self.P, self.Q, self.kchunk, self.c_return)

    # ------------------------------------------------------------------ back ends
    def _lib_gemm(self, m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc):
        import libgpublas_b200 as g
        g.call("dgemm_", "N", "N", m, n, k, float(alpha), g.DevPtr(a_ptr), lda, g.DevPtr(b_ptr), ldb, float(beta), g.DevPtr(c_ptr), ldc)

    def _esize(self):
        return self.abuf[0].element_size()

    # ------------------------------------------------------------------ one full product
    def run(self):
        m, n, k, kc = self.m, self.n, self.k, self.kchunk
        es = self._esize()
        tm, tn = self.r1 - self.r0, self.c1 - self.c0
        comp = torch.cuda.current_stream(self.dev) if self.cuda else None
        for c in range(self.nchunks):
            slot = c % 2
            k0 = c * kc
            kk = min(kc, k - k0)
            a_dst = self.abuf[slot][: m * kk]
            b_dst = self.bbuf[slot][: kk * n]
            # ---- stage + broadcast chunk c (communication stream) ----
            if self.cuda:
                self.comm_stream.wait_event(self.consumed[slot]) if c >= 2 else None
                if c == 0:
                    self.comm_stream.wait_stream(comp)
                ctx = torch.cuda.stream(self.comm_stream)
            else:
                ctx = _Null()
            with ctx:
                if self.rank == 0:
                    a_dst.copy_(self.A[k0 * m:(k0 + kk) * m])                       # A[:, chunk]: contiguous column block
                    b_dst.view(n, kk).copy_(self.B.view(n, k)[:, k0:k0 + kk])       # B[chunk, :] packed: column j -> kk contiguous values
                dist.broadcast(a_dst, src=0)
                dist.broadcast(b_dst, src=0)
                if self.cuda:
                    self.ready[slot].record(self.comm_stream)
            # ---- compute on chunk c (compute stream) ----
            if self.cuda:
                comp.wait_event(self.ready[slot])
            last = c == self.nchunks - 1
            beta = self.beta if c == 0 else 1.0
            a_ptr = a_dst.data_ptr() + es * self.r0                  # rows r0.. of the m x kk chunk (ld = m)
            b_ptr = b_dst.data_ptr() + es * self.c0 * kk             # columns c0.. of the kk x n chunk (ld = kk)
            if tm > 0 and tn > 0:
                if self.c_return == "peer_store" and last:
                    # final chunk: D_home = alpha*A*B + beta*C_local, stored by the kernel's epilogue straight into the
                    # home allocation over NVLink (fused compute + C return)
                    self._gemm_out(tm, tn, kk, self.alpha, a_ptr, m, b_ptr, kk, beta, self.ctile.data_ptr(), tm,
                                   self.peerC + es * (self.r0 + self.c0 * m), m)
                else:
                    self.gemm(tm, tn, kk, self.alpha, a_ptr, m, b_ptr, kk, beta, self.ctile.data_ptr(), tm)
            if self.cuda:
                self.consumed[slot].record(comp)
        if self.c_return == "sendrecv":
            self._gather_sendrecv(tm, tn)
        elif self.cuda:
            # the home rank must not report completion before every peer's stores have landed
            dist.barrier()

    def _gemm_out(self, m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc, d_ptr, ldd):
        import libgpublas_b200 as g
        lib = g.load()
        lib.b200blas_dgemm_out.argtypes = [ctypes.c_char, ctypes.c_char, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                           ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_double,
                                           ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong]
        lib.b200blas_dgemm_out.restype = None
        lib.b200blas_dgemm_out(b"N", b"N", m, n, k, float(alpha), a_ptr, lda, b_ptr, ldb, float(beta), c_ptr, ldc, d_ptr, ldd)

    def _gather_sendrecv(self, tm, tn):
        m = self.m
        if self.rank == 0:
            for r in range(self.world):
                p, q = r // self.Q, r % self.Q
                r0, r1 = block_range(self.m, self.P, p)
                c0, c1 = block_range(self.n, self.Q, q)
                if r1 <= r0 or c1 <= c0:
                    continue
                if r == 0:
                    tile = self.ctile
                else:
                    tile = torch.empty((r1 - r0) * (c1 - c0), dtype=self.dtype, device=self.dev)
                    dist.recv(tile, src=r)
                self.C.view(self.n, m)[c0:c1, r0:r1].copy_(tile.view(c1 - c0, r1 - r0))
        elif tm > 0 and tn > 0:
            dist.send(self.ctile, dst=0)

    def last_kernel_ms(self):
        return self._kernel_ms


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
