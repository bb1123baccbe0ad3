"""GPU parity for the blocked Level-3 routines (SYRK / TRMM / TRSM) at sizes that exercise the
recursion (several diagonal blocks, ragged edges, odd leading dimensions), against the oracle."""
import numpy as np
import pytest

import libgpublas_b200 as g
from helpers import f77, fro, oracle_call, splitmix_uniform

pytestmark = pytest.mark.gpu
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": 2.0 ** -24, "d": 2.0 ** -53, "c": 2.0 ** -24, "z": 2.0 ** -53}


def F(x):
    return np.array(x, order="F")


@pytest.mark.parametrize("p", ["d", "z", "s"])
def test_syrk_vs_oracle(p):
    lib = g.load(); dt = DT[p]
    al, be = ((0.7 - 0.9j), (1.3 - 1.1j)) if p in "cz" else (0.7, 1.3)
    for (n, k) in [(70, 33), (129, 200), (300, 64), (257, 1), (390, 300)]:   # the last one reaches the tensor-core SGEMM tiles
        for uplo in "UL":
            for tr in "NT":
                ra, ca = (n, k) if tr == "N" else (k, n)
                A = splitmix_uniform(1, (ra + 1, ca), dt); C0 = splitmix_uniform(2, (n + 3, n), dt)
                C, R = F(C0), F(C0)
                f77(lib, p + "syrk_", uplo, tr, n, k, al, A, ra + 1, be, C, n + 3)
                assert oracle_call(p + "syrk", uplo, tr, n, k, al, A, ra + 1, be, R, n + 3) == 0
                tri = np.triu(np.ones((n, n), bool)) if uplo == "U" else np.tril(np.ones((n, n), bool))
                full = np.zeros((n + 3, n), bool); full[:n] = tri
                assert np.array_equal(C[~full], C0[~full]), "outside the referenced triangle must be untouched"
                err = fro((C - R)[full]); bound = 4 * (k + 2) * EPS[p] * (abs(al) * fro(A) ** 2 + abs(be) * fro(C0))
                assert err <= bound, (p, n, k, uplo, tr, err, bound)


@pytest.mark.parametrize("which", ["trmm", "trsm"])
@pytest.mark.parametrize("p", ["d", "z", "s"])
def test_trxm_vs_oracle(p, which):
    lib = g.load(); dt = DT[p]
    al = (0.7 - 0.9j) if p in "cz" else 0.7
    hi = np.complex128 if p in "cz" else np.float64
    for (m, n) in [(65, 40), (200, 130), (300, 70), (520, 33), (33, 520)]:
        for side in "LR":
            na = m if side == "L" else n
            T = splitmix_uniform(3, (na + 1, na), dt); T[np.arange(na), np.arange(na)] += dt(na)   # well conditioned
            for uplo in "UL":
                for ta in "NTC":
                    for diag in "NU":
                        if which == "trsm" and diag == "U":
                            T2 = F(T); T2[:na] = T2[:na] / dt(na)      # small off-diagonals keep the unit system tame
                        else:
                            T2 = T
                        B0 = splitmix_uniform(4, (m + 2, n), dt); B, R = F(B0), F(B0)
                        f77(lib, p + which + "_", side, uplo, ta, diag, m, n, al, T2, na + 1, B, m + 2)
                        assert oracle_call(p + which, side, uplo, ta, diag, m, n, al, T2, na + 1, R, m + 2) == 0
                        assert np.array_equal(B[m:], B0[m:])
                        if which == "trmm":
                            err = fro(B[:m] - R[:m]); bound = 4 * (na + 2) * EPS[p] * abs(al) * fro(np.triu(T2[:na]) if uplo == "U" else np.tril(T2[:na])) * fro(B0[:m])
                            assert err <= bound, (p, which, side, uplo, ta, diag, m, n, err, bound)
                        else:
                            # backward form: ||op(A) X - alpha B||_F <= c m eps ||A||_F ||X||_F, c = 4
                            Tm = (np.triu(T2[:na]) if uplo == "U" else np.tril(T2[:na])).astype(hi)
                            if diag == "U":
                                Tm[np.arange(na), np.arange(na)] = 1
                            opT = Tm if ta == "N" else (Tm.T if ta == "T" else Tm.conj().T)
                            X = B[:m].astype(hi)
                            resid = (opT @ X if side == "L" else X @ opT) - complex(dt(al)) * B0[:m].astype(hi) if p in "cz" else \
                                (opT @ X if side == "L" else X @ opT) - float(dt(al)) * B0[:m].astype(hi)
                            assert fro(resid) <= 4 * na * EPS[p] * fro(Tm) * fro(X) + 1e-300, (p, side, uplo, ta, diag, m, n, fro(resid))
                            assert np.allclose(B[:m], R[:m], rtol=1e4 * EPS[p], atol=1e4 * EPS[p])


def graded_triangle(na, cond, dt, seed):
    """Lower-triangular L (leading dimension na+1, rogue last row) with cond_2(L) = cond and O(1) entries: the Cholesky factor of
    Q diag(lambda) Q^H with the spectrum graded geometrically from 1 down to cond^-2.  Unlike a random triangle with a graded
    diagonal (whose inverse grows like 2^n) the solution stays within cond * |B|, and the diagonal BLOCKS of L are ill-conditioned
    to varying degrees -- the case explicit block inverses degrade on."""
    cplx = np.issubdtype(dt, np.complexfloating)
    G = splitmix_uniform(seed, (na, na), np.complex128 if cplx else np.float64)
    Q, _ = np.linalg.qr(G)
    lam = np.power(cond, -2.0 * np.arange(na) / max(na - 1, 1))
    A = (Q * lam) @ Q.conj().T
    A = (A + A.conj().T) / 2
    L = np.linalg.cholesky(A)
    out = np.full((na + 1, na), -77.0, dtype=dt, order="F")
    out[:na] = L.astype(dt)
    return out


@pytest.mark.parametrize("p", ["d", "s", "z"])
def test_trsm_graded_condition(p):
    """TRSM on ILL-CONDITIONED triangles (ADVICE r1 / VERDICT r1 2d): cond_2(T) = 1e2 ... 1e6 in double precision (Cholesky factors of
    SPD matrices with graded spectra, `graded_triangle`), 1e1 ... 1e3 in single -- nothing like the diagonally dominant blocks of
    the other tests.  The leaves solve by substitution (netlib's algorithm), so the result must satisfy the BACKWARD bound of SURVEY 8c
    with the same constant as the well-conditioned case,
        ||op(A) X - alpha B||_F <= c * na * eps * ||A||_F * ||X||_F,   c = 4,
    and the forward error against the float64 solution must stay within c * na * eps * cond.  Sizes cross several leaves and
    recursion levels (na = 200, 520) on both sides, all transposes, both triangles (the upper one is the transposed factor)."""
    lib = g.load(); dt = DT[p]
    hi = np.complex128 if p in "cz" else np.float64
    al = (0.7 - 0.9j) if p in "cz" else 0.7
    conds = [1e1, 1e3] if p == "s" else [1e2, 1e4, 1e6]
    for cond in conds:
        for (m, n) in [(200, 70), (70, 200), (520, 40)]:
            for side in "LR":
                na = m if side == "L" else n
                Llow = graded_triangle(na, cond, dt, 7)
                for uplo in "UL":
                    T = Llow if uplo == "L" else F(np.vstack([Llow[:na].conj().T, Llow[na:]]))
                    Tm = (np.triu(T[:na]) if uplo == "U" else np.tril(T[:na])).astype(hi)
                    for ta in ("NTC" if p == "z" else "NT"):
                        B0 = splitmix_uniform(8, (m + 2, n), dt); B = F(B0)
                        f77(lib, p + "trsm_", side, uplo, ta, "N", m, n, al, T, na + 1, B, m + 2)
                        assert np.array_equal(B[m:], B0[m:])
                        opT = Tm if ta == "N" else (Tm.T if ta == "T" else Tm.conj().T)
                        X = B[:m].astype(hi)
                        assert np.all(np.isfinite(X)), (p, cond, side, uplo, ta)
                        rhs = hi(dt(al)) * B0[:m].astype(hi)
                        resid = (opT @ X if side == "L" else X @ opT) - rhs
                        bound = 4 * na * EPS[p] * fro(Tm) * fro(X)
                        assert fro(resid) <= bound, (p, cond, side, uplo, ta, m, n, fro(resid), bound)
                        Xref = np.linalg.solve(opT, rhs) if side == "L" else np.linalg.solve(opT.T, rhs.T).T
                        assert fro(X - Xref) <= 4 * na * EPS[p] * cond * fro(Xref), (p, cond, side, uplo, ta, fro(X - Xref) / fro(Xref))


def test_dsyrk_dtrsm_large_device_property():
    """Config-4-shaped panels on the device: SYRK equals the masked GEMM; TRSM round trip X*L^T -> B."""
    import torch
    lib = g.load()
    n, k = 2048, 512
    gen = torch.Generator(device="cuda").manual_seed(9)
    A = torch.rand((k, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1   # column-major n x k
    C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    f77(lib, "dsyrk_", "L", "N", n, k, 1.0, A, n, 0.0, C, n)
    ref = (A.T @ A)                               # column-major A A^T == row-major (A^T A) transposed; symmetric
    low = torch.triu(torch.ones((n, n), dtype=torch.bool, device="cuda"))    # row-major upper == column-major lower
    assert ((C - ref)[low].abs().max().item()) <= 16 * 2.0 ** -53 * k
    assert C[~low].abs().max().item() == 0.0     # strictly upper (column-major) never written
    # TRSM: B := B * L^{-T}, then multiply back
    L = torch.tril(torch.rand((k, k), dtype=torch.float64, device="cuda", generator=gen)) + k * torch.eye(k, dtype=torch.float64, device="cuda")
    Lcm = L.T.contiguous()                        # column-major storage of L
    B0 = torch.rand((k, n), dtype=torch.float64, device="cuda", generator=gen)           # column-major n x k
    B = B0.clone()
    torch.cuda.synchronize()
    f77(lib, "dtrsm_", "R", "L", "T", "N", n, k, 1.0, Lcm, k, B, n)
    X = B.T                                        # n x k
    back = X @ L.T
    assert (back - B0.T).abs().max().item() <= 64 * 2.0 ** -53 * k * X.abs().max().item() * k


def test_cblas_level3_row_major_wrappers():
    """cblas_dsyrk / cblas_dsymm / cblas_dtrsm / cblas_dtrmm / cblas_dsyr2k in row-major order map onto the column-major
    core by flipping uplo/side/trans (reference cblas.h:693-824); checked against numpy on row-major arrays."""
    import ctypes
    lib = g.load()
    RowMajor, NoTrans, Trans, Upper, Lower, NonUnit, Left, Right = 101, 111, 112, 121, 122, 131, 141, 142
    d = ctypes.c_double; P = ctypes.c_void_p
    n, k, m = 70, 33, 45
    rng = np.random.default_rng(3)
    A = rng.uniform(-1, 1, (n, k)); C0 = rng.uniform(-1, 1, (n, n)); C = C0.copy()      # C-order (row-major) numpy arrays
    lib.cblas_dsyrk(RowMajor, Upper, NoTrans, n, k, d(0.7), P(A.ctypes.data), k, d(1.3), P(C.ctypes.data), n)
    ref = 0.7 * A @ A.T + 1.3 * C0
    iu = np.triu_indices(n); il = np.tril_indices(n, -1)
    assert np.allclose(C[iu], ref[iu], rtol=1e-13, atol=1e-13) and np.array_equal(C[il], C0[il])
    B = rng.uniform(-1, 1, (n, k)); C = C0.copy()
    lib.cblas_dsyr2k(RowMajor, Lower, NoTrans, n, k, d(0.7), P(A.ctypes.data), k, P(B.ctypes.data), k, d(1.3), P(C.ctypes.data), n)
    ref = 0.7 * (A @ B.T + B @ A.T) + 1.3 * C0
    il0 = np.tril_indices(n); iu1 = np.triu_indices(n, 1)
    assert np.allclose(C[il0], ref[il0], rtol=1e-13, atol=1e-13) and np.array_equal(C[iu1], C0[iu1])
    S = rng.uniform(-1, 1, (m, m)); S = S + S.T; Bm = rng.uniform(-1, 1, (m, n)); Cm0 = rng.uniform(-1, 1, (m, n)); Cm = Cm0.copy()
    Sl = np.tril(S) + np.triu(np.full((m, m), 1e10), 1)                                   # upper part poisoned: must not be read
    lib.cblas_dsymm(RowMajor, Left, Lower, m, n, d(0.7), P(Sl.ctypes.data), m, P(Bm.ctypes.data), n, d(1.3), P(Cm.ctypes.data), n)
    assert np.allclose(Cm, 0.7 * S @ Bm + 1.3 * Cm0, rtol=1e-12, atol=1e-12)
    T = np.triu(rng.uniform(-1, 1, (m, m))) + m * np.eye(m); X = Bm.copy()
    lib.cblas_dtrsm(RowMajor, Left, Upper, NoTrans, NonUnit, m, n, d(2.0), P(T.ctypes.data), m, P(X.ctypes.data), n)
    assert np.allclose(T @ X, 2.0 * Bm, rtol=1e-11, atol=1e-11)
    X = Bm.copy()
    lib.cblas_dtrmm(RowMajor, Right, Upper, Trans, NonUnit, m, n, d(2.0), P(np.ascontiguousarray(np.triu(rng.uniform(-1, 1, (n, n)))).ctypes.data), n, P(X.ctypes.data), n)
    # recompute with the same triangular factor
    rng2 = np.random.default_rng(3); _ = rng2.uniform(-1, 1, (n, k)); _ = rng2.uniform(-1, 1, (n, n)); _ = rng2.uniform(-1, 1, (n, k))
    _ = rng2.uniform(-1, 1, (m, m)); _ = rng2.uniform(-1, 1, (m, n)); _ = rng2.uniform(-1, 1, (m, n)); _ = rng2.uniform(-1, 1, (m, m))
    U = np.triu(rng2.uniform(-1, 1, (n, n)))
    assert np.allclose(X, 2.0 * Bm @ U.T, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("p", ["d", "z", "s"])
def test_next_family_larger_shapes_vs_oracle(p):
    """SYMM/SYR2K (+ HEMM/HERK/HER2K for z) at sizes that run on the tensor-pipe tiles (masked DMMA / tcgen05 launches,
    ragged edges, odd leading dimensions), against the oracle; untouched triangle / padding verified."""
    lib = g.load(); dt = DT[p]
    al, be = ((0.7 - 0.9j), (1.3 - 1.1j)) if p == "z" else (0.7, 1.3)
    n, k, m = 330, 270, 290
    tol = lambda kk, *mats: 8 * (kk + 2) * EPS[p] * np.prod([fro(x) for x in mats])
    for uplo in "UL":
        tri = np.triu(np.ones((n, n), bool)) if uplo == "U" else np.tril(np.ones((n, n), bool))
        full = np.zeros((n + 3, n), bool); full[:n] = tri
        for tr in "NT":
            ra, ca = (n, k) if tr == "N" else (k, n)
            A = splitmix_uniform(1, (ra + 1, ca), dt); B = splitmix_uniform(2, (ra + 1, ca), dt); C0 = splitmix_uniform(3, (n + 3, n), dt)
            C, R = F(C0), F(C0)
            f77(lib, p + "syr2k_", uplo, tr, n, k, al, A, ra + 1, B, ra + 1, be, C, n + 3)
            assert oracle_call(p + "syr2k", uplo, tr, n, k, al, A, ra + 1, B, ra + 1, be, R, n + 3) == 0
            assert np.array_equal(C[~full], C0[~full])
            assert fro((C - R)[full]) <= 2 * abs(al) * tol(k, A, B) + 8 * EPS[p] * abs(be) * fro(C0)
        if p == "z":
            for tr in "NC":
                ra, ca = (n, k) if tr == "N" else (k, n)
                A = splitmix_uniform(4, (ra + 1, ca), dt); B = splitmix_uniform(5, (ra + 1, ca), dt); C0 = splitmix_uniform(6, (n + 3, n), dt)
                C, R = F(C0), F(C0)
                f77(lib, "zherk_", uplo, tr, n, k, np.float64(0.7), A, ra + 1, np.float64(1.3), C, n + 3)
                assert oracle_call("zherk", uplo, tr, n, k, np.float64(0.7), A, ra + 1, np.float64(1.3), R, n + 3) == 0
                assert np.array_equal(C[~full], C0[~full]) and np.all(np.diag(C[:n]).imag == 0)
                assert fro((C - R)[full]) <= tol(k, A, A) + 8 * EPS[p] * 1.3 * fro(C0)
                C, R = F(C0), F(C0)
                f77(lib, "zher2k_", uplo, tr, n, k, al, A, ra + 1, B, ra + 1, np.float64(1.3), C, n + 3)
                assert oracle_call("zher2k", uplo, tr, n, k, al, A, ra + 1, B, ra + 1, np.float64(1.3), R, n + 3) == 0
                assert np.array_equal(C[~full], C0[~full]) and np.all(np.diag(C[:n]).imag == 0)
                assert fro((C - R)[full]) <= 2 * abs(al) * tol(k, A, B) + 8 * EPS[p] * 1.3 * fro(C0)
        for side in "LR":
            na = m if side == "L" else n
            S = splitmix_uniform(7, (na + 1, na), dt); Bm = splitmix_uniform(8, (m + 2, n), dt); C0 = splitmix_uniform(9, (m + 1, n), dt)
            for r in (("symm", "hemm") if p == "z" else ("symm",)):
                C, R = F(C0), F(C0)
                f77(lib, p + r + "_", side, uplo, m, n, al, S, na + 1, Bm, m + 2, be, C, m + 1)
                assert oracle_call(p + r, side, uplo, m, n, al, S, na + 1, Bm, m + 2, be, R, m + 1) == 0
                assert np.array_equal(C[m:], C0[m:])
                assert fro(C[:m] - R[:m]) <= abs(al) * tol(na, S, Bm) + 8 * EPS[p] * abs(be) * fro(C0)


@pytest.mark.parametrize("p", ["d", "s", "z"])
def test_syrk_pipelined_staging_host_operands(p):
    """Host-resident ?syrk_ operands above pipeline_min bytes: A in k-chunks, C out trapezoid by trapezoid under the multiply
    (csrc/staged_level3.cuh; the reference's miss path is whole-array blocking copies, runtime-mem.hpp:84-165).  Both triangles,
    both transposes, beta = 0 and != 0, ragged sizes, odd leading dimensions; same bar as the resident path, the other triangle
    and the padding rows untouched, and the strictly unreferenced part of C never crosses PCIe (byte counters)."""
    lib = g.load(); dt = DT[p]
    lib.b200blas_set_options(b"pipeline_min=1000")
    try:
        al, be = ((0.7 - 0.9j), (1.3 - 1.1j)) if p in "cz" else (0.7, 1.3)
        hi = np.complex128 if p in "cz" else np.float64
        n, k = 1500, 2300
        for uplo in "UL":
            for tr in "NT":
                for beta in (be, 0.0 * be):
                    ra, ca = (n, k) if tr == "N" else (k, n)
                    A = splitmix_uniform(11, (ra + 1, ca), dt); C0 = splitmix_uniform(12, (n + 3, n), dt)
                    C = F(C0)
                    s0 = g.stats()
                    f77(lib, p + "syrk_", uplo, tr, n, k, al, A, ra + 1, beta, C, n + 3)
                    s1 = g.stats()
                    opA = A[:ra].astype(hi); opA = opA if tr == "N" else opA.T
                    ref = complex(al) * (opA @ opA.T) + complex(beta) * C0[:n].astype(hi) if p in "cz" else al * (opA @ opA.T) + beta * C0[:n].astype(hi)
                    tri = np.triu(np.ones((n, n), bool)) if uplo == "U" else np.tril(np.ones((n, n), bool))
                    full = np.zeros((n + 3, n), bool); full[:n] = tri
                    assert np.array_equal(C[~full], C0[~full]), "outside the referenced triangle must be untouched"
                    err = fro((C[:n] - ref)[tri]); bound = 4 * (k + 2) * EPS[p] * (abs(al) * fro(A[:ra]) ** 2 + abs(beta) * fro(C0[:n]))
                    assert err <= bound, (p, uplo, tr, beta, err, bound)
                    es = np.dtype(dt).itemsize
                    assert s1["h2d_bytes"] - s0["h2d_bytes"] < (n * k + (0.66 if beta != 0 else 0.25) * n * n) * es, "A once, at most the trapezoids of C"
                    assert s1["d2h_bytes"] - s0["d2h_bytes"] < 0.66 * n * n * es, "only the trapezoids return"
    finally:
        lib.b200blas_set_options(b"pipeline_min=67108864")


@pytest.mark.parametrize("which", ["trsm", "trmm"])
@pytest.mark.parametrize("p", ["d", "s", "z"])
def test_trxm_pipelined_staging_host_operands(p, which):
    """Host-resident B above pipeline_min bytes goes through ?trsm_/?trmm_ in blocks of right-hand sides (columns for side L,
    rows for side R), block p+1 travelling in and block p-1 out under block p's work (csrc/staged_level3.cuh).  A in host memory
    or already on the device."""
    import torch
    lib = g.load(); dt = DT[p]
    lib.b200blas_set_options(b"pipeline_min=1000")
    try:
        al = (0.7 - 0.9j) if p in "cz" else 0.7
        hi = np.complex128 if p in "cz" else np.float64
        for (side, m, n) in [("L", 300, 2300), ("R", 2500, 260)]:
            na = m if side == "L" else n
            T = F(splitmix_uniform(13, (na + 1, na), dt)); T[:na] = T[:na] / dt(na); T[np.arange(na), np.arange(na)] = dt(2.0)
            for uplo in "UL":
                for ta in "NC":
                    for diag, a_on_device in (("N", False), ("U", True)):
                        B0 = splitmix_uniform(14, (m + 2, n), dt); B = F(B0)
                        a_arg = T
                        if a_on_device:
                            a_arg = torch.from_numpy(np.ascontiguousarray(T.ravel(order="F").view(np.float64 if p in "dz" else np.float32))).cuda()
                        s0 = g.stats()
                        f77(lib, p + which + "_", side, uplo, ta, diag, m, n, al, a_arg, na + 1, B, m + 2)
                        s1 = g.stats()
                        es = np.dtype(dt).itemsize
                        assert s1["d2h_bytes"] - s0["d2h_bytes"] == m * n * es
                        cg = max(256, -(-((na + 7) // 8) // 128) * 128)
                        trap = sum((min(c0 + cg, na) if uplo == "U" else na - c0) * min(cg, na - c0) for c0 in range(0, na, cg))
                        assert s1["h2d_bytes"] - s0["h2d_bytes"] == (m * n + (0 if a_on_device else trap)) * es, "B once, the referenced trapezoids of A once"
                        assert np.array_equal(B[m:], B0[m:])
                        Tm = (np.triu(T[:na]) if uplo == "U" else np.tril(T[:na])).astype(hi)
                        if diag == "U":
                            Tm[np.arange(na), np.arange(na)] = 1
                        opT = Tm if ta == "N" else Tm.conj().T
                        X = B[:m].astype(hi); Bh = B0[:m].astype(hi)
                        a_s = complex(dt(al)) if p in "cz" else float(dt(al))
                        if which == "trmm":
                            ref = a_s * (opT @ Bh if side == "L" else Bh @ opT)
                            assert fro(X - ref) <= 4 * (na + 2) * EPS[p] * abs(al) * fro(Tm) * fro(Bh), (p, side, uplo, ta, diag)
                        else:
                            resid = (opT @ X if side == "L" else X @ opT) - a_s * Bh
                            assert fro(resid) <= 4 * na * EPS[p] * (fro(Tm) * fro(X) + abs(al) * fro(Bh)), (p, side, uplo, ta, diag, fro(resid))
    finally:
        lib.b200blas_set_options(b"pipeline_min=67108864")
