"""Blocked right-looking Cholesky (lower) -- the workload of BASELINE.json configs[3]: per panel a diagonal-block
factorisation, a DTRSM on the panel below it and a DSYRK trailing update, all issued through the interposed Fortran
symbols exactly as a LAPACK-style caller would (SURVEY.md section 8d "C4").

    blocked_cholesky(n, A, lda, nb)            one GPU, A device/managed/host
    TiledCholesky(n, nb, device, rank, world)  one process per GPU: 1-D block-cyclic column ownership, the owner
                                               factors its panel and broadcasts it, every rank updates the block
                                               columns it owns (look-ahead: the next panel's column first)
"""
import ctypes

import torch
import torch.distributed as dist

from . import DevPtr, call, load


def _ptr(A):
    if isinstance(A, DevPtr):
        return A.addr
    if hasattr(A, "data_ptr"):
        return A.data_ptr()
    return A.ctypes.data


def potrf_lower(n, A, lda):
    """In-place lower Cholesky of an n x n column-major block (device, managed or host memory); returns LAPACK info."""
    lib = load()
    lib.b200blas_dpotrf_lower.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong]
    lib.b200blas_dpotrf_lower.restype = ctypes.c_int
    return lib.b200blas_dpotrf_lower(n, ctypes.c_void_p(_ptr(A)), lda)


def blocked_cholesky(n, A, lda, nb=2048, lookahead=True):
    """A := L (lower triangle) with A = L L^T, column-major, leading dimension lda.  Returns LAPACK info.

    lookahead (device-resident A, library in asynchronous mode `set_sync(False)`): the trailing update of step J is split
    into the next block column and the rest; as soon as the next block column is up to date its diagonal-block
    factorisation and panel solve (latency-bound, a few SMs) run on a second stream underneath the rest of step J's
    rank-nb update (compute-bound), so the tensor pipe does not idle during the panel work.  Only the update kernels
    (DSYRK/DGEMM, no workspace) run concurrently with the panel kernels (which use the per-thread workspace)."""
    base = _ptr(A)
    at = lambda i, j: DevPtr(base + 8 * (i + j * lda))
    dev = getattr(A, "device", None)
    use_la = bool(lookahead) and dev is not None and getattr(dev, "type", "") == "cuda" and n > 2 * nb
    if not use_la:
        for j in range(0, n, nb):
            jb = min(nb, n - j)
            info = potrf_lower(jb, at(j, j), lda)
            if info:
                return info + j
            rest = n - j - jb
            if rest > 0:
                call("dtrsm_", "R", "L", "T", "N", rest, jb, 1.0, at(j, j), lda, at(j + jb, j), lda)
                call("dsyrk_", "L", "N", rest, jb, -1.0, at(j + jb, j), lda, 1.0, at(j + jb, j + jb), lda)
        return 0

    from . import use_torch_stream
    use_torch_stream()              # the library's calls follow torch's current stream from here on (main, then panel)
    main = torch.cuda.current_stream(dev)
    # high priority: the update kernels keep every SM occupied, and the block scheduler hands freed SMs to the running
    # kernel's own pending CTAs first -- only a higher-priority stream gets its (small) panel kernels in between
    panel = torch.cuda.Stream(device=dev, priority=-1)

    class on:                       # issue the enclosed BLAS calls on `stream`
        def __init__(s, stream): s.stream = stream
        def __enter__(s): s.ctx = torch.cuda.stream(s.stream); s.ctx.__enter__(); use_torch_stream()
        def __exit__(s, *a): s.ctx.__exit__(*a); use_torch_stream()

    def factor(j, jb, rest):        # diagonal block + panel solve of block column j
        r = potrf_lower(jb, at(j, j), lda)
        if r == 0 and rest > 0:
            call("dtrsm_", "R", "L", "T", "N", rest, jb, 1.0, at(j, j), lda, at(j + jb, j), lda)
        return r + j if r else 0

    info = factor(0, min(nb, n), n - min(nb, n))
    panel_done = None
    for j in range(0, n, nb):
        if info:
            break
        jb = min(nb, n - j); rest = n - j - jb
        if rest <= 0:
            break
        if panel_done is not None:
            main.wait_event(panel_done)                     # panel j was factored on the panel stream
        t0 = j + jb; nb2 = min(nb, rest); below = rest - nb2
        # next block column first
        call("dsyrk_", "L", "N", nb2, jb, -1.0, at(t0, j), lda, 1.0, at(t0, t0), lda)
        if below > 0:
            call("dgemm_", "N", "T", below, nb2, jb, -1.0, at(t0 + nb2, j), lda, at(t0, j), lda, 1.0, at(t0 + nb2, t0), lda)
        col_ready = torch.cuda.Event(); col_ready.record(main)
        # the rest of the trailing matrix (queued BEFORE the panel work, whose info read-back blocks the host)
        if below > 0:
            call("dsyrk_", "L", "N", below, jb, -1.0, at(t0 + nb2, j), lda, 1.0, at(t0 + nb2, t0 + nb2), lda)
        panel.wait_event(col_ready)
        with on(panel):
            info = factor(t0, nb2, below)
            panel_done = torch.cuda.Event(); panel_done.record(panel)
    main.wait_stream(panel)
    return info


class LibBlas:
    """The four block operations of the workload on a column-major n x n buffer (ld = n), through the library's
    Fortran symbols.  TiledCholesky takes any object with these methods, so its host logic (ownership, ordering,
    broadcasts) can be exercised on CPU with gloo by injecting a reference back end (tests/test_multigpu_cpu.py)."""

    def __init__(self, A, n):
        self.base, self.n = A.data_ptr(), n

    def at(self, i, j):
        return DevPtr(self.base + 8 * (i + j * self.n))

    def potrf(self, j, jb):
        return potrf_lower(jb, self.at(j, j), self.n)

    def trsm(self, j, jb, rest):                  # A[j+jb:, j:j+jb] := A[j+jb:, j:j+jb] * L(j,j)^-T
        call("dtrsm_", "R", "L", "T", "N", rest, jb, 1.0, self.at(j, j), self.n, self.at(j + jb, j), self.n)

    def syrk(self, kcol, kb, j, jb):              # A[kcol:kcol+kb, kcol:kcol+kb] -= P P^T (lower), P = A[kcol:kcol+kb, j:j+jb]
        call("dsyrk_", "L", "N", kb, jb, -1.0, self.at(kcol, j), self.n, 1.0, self.at(kcol, kcol), self.n)

    def gemm(self, kcol, kb, below, j, jb):       # A[kcol+kb:, kcol:kcol+kb] -= A[kcol+kb:, j:j+jb] * A[kcol:kcol+kb, j:j+jb]^T
        call("dgemm_", "N", "T", below, kb, jb, -1.0, self.at(kcol + kb, j), self.n, self.at(kcol, j), self.n, 1.0,
             self.at(kcol + kb, kcol), self.n)


class TiledCholesky:
    """Strong-scaled lower Cholesky of one n x n matrix over `world` GPUs (one process each).

    Block column J (nb wide) is owned by rank J % world.  Every rank holds a full-size copy of the matrix but keeps
    only its own block columns up to date.  Step J: the owner factors the diagonal block and solves the panel below
    it, then the panel is broadcast; every rank applies the rank-nb update to the block columns it owns.  The owner
    of column J+1 updates that column FIRST and factors/broadcasts it on a second stream while its remaining
    columns are still being updated (look-ahead of one panel).  Every rank ends up with the complete factor: the
    panels arrive in the broadcasts anyway, so there is no separate gather."""

    def __init__(self, n, nb, device, rank, world, blas=None):
        self.n, self.nb, self.dev, self.rank, self.world = n, nb, device, rank, world
        self.blas_factory = blas or LibBlas
        self.nblk = (n + nb - 1) // nb
        self.A = None

    def owner(self, J):
        return J % self.world

    def set_matrix(self, A):
        """A: torch float64 1-D buffer of n*n (column-major, ld = n) on every rank; rank 0's content is the input."""
        self.A = A
        dist.broadcast(self.A, src=0)

    def run(self):
        import libgpublas_b200 as g
        n, nb, A = self.n, self.nb, self.A
        blas = self.blas_factory(A, n)
        cuda = self.dev.type == "cuda"
        main = torch.cuda.current_stream(self.dev) if cuda else None
        if cuda and not hasattr(self, "panel_stream"):
            self.panel_stream = torch.cuda.Stream(device=self.dev, priority=-1)    # see blocked_cholesky: panel kernels must get SMs in between
        info = 0

        class on:                       # run the enclosed BLAS calls / collectives on the given stream
            def __init__(s, stream): s.stream = stream
            def __enter__(s):
                if cuda:
                    s.ctx = torch.cuda.stream(s.stream); s.ctx.__enter__(); g.use_torch_stream()
            def __exit__(s, *a):
                if cuda:
                    s.ctx.__exit__(*a); g.use_torch_stream()

        def factor_panel(J):
            j = J * nb; jb = min(nb, n - j); rest = n - j - jb
            r = blas.potrf(j, jb)
            if rest > 0 and r == 0:
                blas.trsm(j, jb, rest)
            return r + j if r else 0

        if cuda:
            self.panel_stream.wait_stream(main)
        if self.owner(0) == self.rank:
            with on(self.panel_stream if cuda else None):
                info = factor_panel(0)
        for J in range(self.nblk):
            j = J * nb; jb = min(nb, n - j); rest = n - j - jb
            # panel J = rows j.. of columns j..j+jb; the column block [j*n, (j+jb)*n) is one contiguous range
            with on(self.panel_stream if cuda else None):
                dist.broadcast(A[j * n:(j + jb) * n], src=self.owner(J))
                if cuda:
                    arrived = torch.cuda.Event(); arrived.record(self.panel_stream)
            if rest <= 0:
                continue
            if cuda:
                main.wait_event(arrived)
            # trailing update of the block columns this rank owns; the next panel's column goes first so that its
            # owner can factor and broadcast it on the panel stream while the remaining columns are still updating
            mine = [K for K in range(J + 1, self.nblk) if self.owner(K) == self.rank]
            for K in mine:
                kcol = K * nb; kb = min(nb, n - kcol); below = n - kcol - kb
                blas.syrk(kcol, kb, j, jb)
                if below > 0:
                    blas.gemm(kcol, kb, below, j, jb)
                if K == J + 1 and cuda:
                    ready = torch.cuda.Event(); ready.record(main)
                    self.panel_stream.wait_event(ready)
            if mine and mine[0] == J + 1 and info == 0:
                with on(self.panel_stream if cuda else None):
                    info = factor_panel(J + 1)
        if cuda:
            main.wait_stream(self.panel_stream)
        t = torch.tensor([info], device=self.dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return int(t.item())
