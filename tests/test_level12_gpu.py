"""GPU parity tests for Level 1 / Level 2 through the C ABI, against the oracle (netlib restatement)
and the CPU BLAS (OpenBLAS).  Tolerances per SURVEY.md section 8c:
  DDOT  |r - r_ref| <= 2 n eps sum|x_i y_i|;  DNRM2 relative 4 eps log2 n;  DAXPY element-wise
  <= 2 eps (|alpha x| + |y|);  I?AMAX exact;  GEMV like GEMM (DMMCH ratio < 16); TRSV backward error."""
import ctypes

import numpy as np
import pytest

import libgpublas_b200 as g
from helpers import f77, load_openblas, oracle_call, splitmix_uniform

pytestmark = pytest.mark.gpu

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": 2.0 ** -24, "d": 2.0 ** -53, "c": 2.0 ** -24, "z": 2.0 ** -53}
REAL = {"s": ctypes.c_float, "d": ctypes.c_double}


class C32(ctypes.Structure):
    _fields_ = [("re", ctypes.c_float), ("im", ctypes.c_float)]


class C64(ctypes.Structure):
    _fields_ = [("re", ctypes.c_double), ("im", ctypes.c_double)]


SIZES = [1, 2, 3, 31, 256, 1000, 4097, 100003]
INCS = [(1, 1), (2, 3), (-1, 1), (-2, -3), (3, -1)]


def strided(seed, n, inc, dt):
    span = 1 + (n - 1) * abs(inc)
    return splitmix_uniform(seed, (span,), dt)


def logical(v, n, inc):
    """BLAS view of a strided vector as a length-n array (negative inc walks backwards)."""
    a = abs(inc)
    e = v[:: a][:n]
    return e if inc > 0 else e[::-1]


@pytest.mark.parametrize("p", ["s", "d"])
def test_dot_axpy_real(p):
    lib = g.load(); dt = DT[p]
    for n in SIZES:
        for (ix, iy) in INCS:
            x = strided(1, n, ix, dt); y = strided(2, n, iy, dt)
            r = f77(lib, p + "dot_", n, x, ix, y, iy, restype=REAL[p])
            lx, ly = logical(x, n, ix).astype(np.float64), logical(y, n, iy).astype(np.float64)
            ref = float(np.dot(lx, ly))
            assert abs(r - ref) <= 2 * n * EPS[p] * float(np.abs(lx * ly).sum()) + 1e-300, (p, n, ix, iy, r, ref)
            y0 = y.copy(); alpha = 0.7
            f77(lib, p + "axpy_", n, alpha, x, ix, y, iy)
            want = logical(y0, n, iy).astype(np.float64) + np.float64(dt(alpha)) * lx
            got = logical(y, n, iy).astype(np.float64)
            tol = 2 * EPS[p] * (np.abs(np.float64(dt(alpha)) * lx) + np.abs(logical(y0, n, iy).astype(np.float64)))
            assert np.all(np.abs(got - want) <= tol + 1e-300), (p, n, ix, iy)
            # gaps between strided elements untouched
            mask = np.ones(y.shape, bool); mask[:: abs(iy)][:n] = False
            assert np.array_equal(y[mask], y0[mask])


@pytest.mark.parametrize("p", ["c", "z"])
def test_dot_axpy_complex(p):
    lib = g.load(); dt = DT[p]; CT = C32 if p == "c" else C64
    for n in [1, 3, 257, 5001]:
        for (ix, iy) in [(1, 1), (2, -3)]:
            x = strided(3, n, ix, dt); y = strided(4, n, iy, dt)
            lx, ly = logical(x, n, ix).astype(np.complex128), logical(y, n, iy).astype(np.complex128)
            for name, ref in [("dotu_", np.dot(lx, ly)), ("dotc_", np.vdot(lx, ly))]:
                r = f77(lib, p + name, n, x, ix, y, iy, restype=CT)
                got = complex(r.re, r.im)
                assert abs(got - ref) <= 4 * n * EPS[p] * float(np.abs(lx * ly).sum()), (p, name, n, got, ref)
            y0 = y.copy(); alpha = 0.7 - 0.9j
            f77(lib, p + "axpy_", n, alpha, x, ix, y, iy)
            want = logical(y0, n, iy).astype(np.complex128) + complex(dt(alpha)) * lx
            assert np.allclose(logical(y, n, iy), want, rtol=8 * EPS[p], atol=8 * EPS[p])


def test_nrm2_asum_scal_copy_swap():
    lib = g.load()
    for p, name, rt in [("d", "dnrm2_", ctypes.c_double), ("s", "snrm2_", ctypes.c_float), ("z", "dznrm2_", ctypes.c_double), ("c", "scnrm2_", ctypes.c_float)]:
        dt = DT[p]
        for n in SIZES:
            for inc in (1, 3):
                x = strided(5, n, inc, dt)
                r = f77(lib, name, n, x, inc, restype=rt)
                ref = float(np.linalg.norm(logical(x, n, inc).astype(np.complex128 if p in "cz" else np.float64)))
                assert abs(r - ref) <= 4 * EPS[p] * max(np.log2(max(n, 2)), 1) * ref, (name, n, inc, r, ref)
    # no overflow / underflow (Blue's scaling): plain sum of squares would give inf / 0
    for scale in (1e200, 1e-200):
        x = splitmix_uniform(6, (1000,)) * scale
        r = f77(lib, "dnrm2_", 1000, x, 1, restype=ctypes.c_double)
        ref = float(np.linalg.norm(x / scale)) * scale
        assert abs(r - ref) <= 1e-13 * ref
    assert f77(lib, "dnrm2_", 0, np.zeros(1), 1, restype=ctypes.c_double) == 0.0
    assert f77(lib, "dnrm2_", 5, np.ones(5), 0, restype=ctypes.c_double) == 0.0
    x = splitmix_uniform(7, (4097,))
    r = f77(lib, "dasum_", 4097, x, 1, restype=ctypes.c_double)
    assert abs(r - np.abs(x).sum()) <= 4097 * 2.0 ** -52 * np.abs(x).sum()
    y = x.copy(); f77(lib, "dscal_", 4097, 1.3, y, 1); assert np.array_equal(y, 1.3 * x)
    y = np.zeros(4097); f77(lib, "dcopy_", 4097, x, 1, y, 1); assert np.array_equal(y, x)
    y = np.zeros(2 * 4097); f77(lib, "dcopy_", 4097, x, -1, y, 2); assert np.array_equal(y[::2], x[::-1]) and np.all(y[1::2] == 0)
    a = splitmix_uniform(8, (1000,)); b = splitmix_uniform(9, (1000,)); a0, b0 = a.copy(), b.copy()
    f77(lib, "dswap_", 1000, a, 1, b, 1); assert np.array_equal(a, b0) and np.array_equal(b, a0)


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_row_operations_inside_one_host_matrix(p):
    """LAPACK-style row operations on ONE untracked host matrix: dswap_(n, &A[i], lda, &A[j], lda) (dgetf2 / dlaswp row swaps),
    drot_ on two rows (dbdsqr), dscal_/daxpy_ on a row.  x and y interleave in the same host array, so a staged copy of the whole
    extent of one operand contains the other's elements: each operand must write back ONLY its own elements (ADVICE r1, high:
    the second whole-extent write-back used to overwrite the first with stale data, leaving row i unswapped except column 0),
    and every other element of A must be bit-unchanged."""
    lib = g.load(); dt = DT[p]
    for (m, n, i, j) in [(7, 5, 1, 4), (40, 33, 0, 39), (64, 64, 17, 3), (5, 300, 4, 0)]:
        A0 = splitmix_uniform(21, (m, n), dt); lda = m
        # swap
        A = A0.copy(order="F")
        f77(lib, p + "swap_", n, A[i:, 0], lda, A[j:, 0], lda)
        want = A0.copy(order="F"); want[[i, j], :] = want[[j, i], :]
        assert np.array_equal(A, want), (p, "swap", m, n, i, j)
        # scal on a row: the other rows stay bit-identical
        A = A0.copy(order="F"); alpha = 0.7 if p in "sd" else 0.7 - 0.9j
        f77(lib, p + "scal_", n, alpha, A[i:, 0], lda)
        want = A0.copy(order="F"); want[i, :] = (dt(alpha) * want[i, :].astype(dt)).astype(dt)
        rest = np.ones(m, bool); rest[i] = False
        assert np.array_equal(A[rest], A0[rest]) and np.allclose(A[i], want[i], rtol=4 * EPS[p], atol=0), (p, "scal", m, n)
        # axpy: row j += alpha * row i
        A = A0.copy(order="F")
        f77(lib, p + "axpy_", n, alpha, A[i:, 0], lda, A[j:, 0], lda)
        rest = np.ones(m, bool); rest[j] = False
        assert np.array_equal(A[rest], A0[rest]), (p, "axpy rest", m, n)
        assert np.allclose(A[j], A0[j] + dt(alpha) * A0[i], rtol=8 * EPS[p], atol=8 * EPS[p]), (p, "axpy", m, n)
        if p in "sd":
            A = A0.copy(order="F"); c, sn = 0.6, 0.8
            f77(lib, p + "rot_", n, A[i:, 0], lda, A[j:, 0], lda, c, sn)
            rest = np.ones(m, bool); rest[[i, j]] = False
            assert np.array_equal(A[rest], A0[rest]), (p, "rot rest", m, n)
            assert np.allclose(A[i], dt(c) * A0[i] + dt(sn) * A0[j], rtol=8 * EPS[p], atol=8 * EPS[p])
            assert np.allclose(A[j], dt(c) * A0[j] - dt(sn) * A0[i], rtol=8 * EPS[p], atol=8 * EPS[p])
            # columns of the same matrix (contiguous, disjoint): dswap of two columns
            A = A0.copy(order="F")
            f77(lib, p + "swap_", m, A[:, 0], 1, A[:, n - 1], 1)
            want = A0.copy(order="F"); want[:, [0, n - 1]] = want[:, [n - 1, 0]]
            assert np.array_equal(A, want)


def test_iamax_exact_and_tiebreak():
    """I?AMAX must be bit-exact: the FIRST index of maximum |x| (1-based), like the CPU BLAS."""
    lib = g.load(); ob = load_openblas()
    for n in SIZES + [2 ** 20 + 7]:
        for inc in (1, 2):
            x = strided(10, n, inc, np.float64)
            r = f77(lib, "idamax_", n, x, inc, restype=ctypes.c_int)
            assert r == oracle_call("idamax", n, x, inc)
            if ob is not None:
                assert r == f77(ob, "idamax_", n, x, inc, restype=ctypes.c_int)
    n = 2 ** 20 + 7
    x = splitmix_uniform(11, (n,))
    for (i1, i2, i3) in [(5, 70001, n - 1), (0, 1, 2), (n - 3, n - 2, n - 1), (123457, 123458, 999999)]:
        y = x.copy(); y[i1] = 1.5; y[i2] = 1.5; y[i3] = -1.5
        assert f77(lib, "idamax_", n, y, 1, restype=ctypes.c_int) == i1 + 1
        y[i1] = -1.5   # sign must not matter
        assert f77(lib, "idamax_", n, y, 1, restype=ctypes.c_int) == i1 + 1
        assert lib.cblas_idamax(n, y.ctypes.data_as(ctypes.c_void_p), 1) == i1   # CBLAS: 0-based
    assert f77(lib, "idamax_", 0, x, 1, restype=ctypes.c_int) == 0
    assert f77(lib, "idamax_", 5, x, 0, restype=ctypes.c_int) == 0
    xs = splitmix_uniform(12, (5000,), np.float32); xs[77] = 2.0; xs[4000] = -2.0
    assert f77(lib, "isamax_", 5000, xs, 1, restype=ctypes.c_int) == 78
    xz = splitmix_uniform(13, (3000,), np.complex128)
    assert f77(lib, "izamax_", 3000, xz, 1, restype=ctypes.c_int) == oracle_call("izamax", 3000, xz, 1)


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_gemv(p):
    lib = g.load(); dt = DT[p]
    hi = np.complex128 if p in "cz" else np.float64
    alpha, beta = ((0.7 - 0.9j), (1.3 - 1.1j)) if p in "cz" else (0.7, 1.3)
    for (m, n) in [(1, 1), (3, 5), (64, 64), (257, 129), (1000, 37), (37, 1000), (700, 900)]:
        for trans in ("N", "T", "C"):
            for (ix, iy, ldp) in [(1, 1, 0), (2, -1, 1), (-3, 2, 3)]:
                lda = m + ldp
                A = splitmix_uniform(20, (lda, n), dt)
                lenx, leny = (n, m) if trans == "N" else (m, n)
                x = strided(21, lenx, ix, dt); y = strided(22, leny, iy, dt); y0 = y.copy()
                for (al, be) in [(alpha, beta), (alpha, 0.0 * alpha)]:
                    y[:] = y0
                    f77(lib, p + "gemv_", trans, m, n, al, A, lda, x, ix, be, y, iy)
                    a = A[:m].astype(hi); a = a if trans == "N" else (a.T if trans == "T" else a.conj().T)
                    lx = logical(x, lenx, ix).astype(hi); ly0 = logical(y0, leny, iy).astype(hi)
                    ref = complex(dt(al)) * (a @ lx) + (complex(dt(be)) * ly0 if be != 0 else 0) if p in "cz" else \
                        float(dt(al)) * (a @ lx) + (float(dt(be)) * ly0 if be != 0 else 0)
                    gb = abs(al) * (np.abs(a) @ np.abs(lx)) + abs(be) * np.abs(ly0)
                    ratio = (np.abs(logical(y, leny, iy).astype(hi) - ref) / (EPS[p] * np.maximum(gb, 1e-300))).max()
                    assert ratio < 16, (p, m, n, trans, ix, iy, ratio)
                    mask = np.ones(y.shape, bool); mask[:: abs(iy)][:leny] = False
                    assert np.array_equal(y[mask], y0[mask])


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_trsv(p):
    lib = g.load(); dt = DT[p]
    hi = np.complex128 if p in "cz" else np.float64
    for n in [1, 2, 5, 64, 65, 200, 513]:
        # off-diagonals O(1/n) so that the unit-diagonal systems are well conditioned too (a random
        # unit-triangular matrix with O(1) entries has a solution growing like 2^n: overflow in c/s)
        A = splitmix_uniform(30, (n + 1, n), dt) * dt(min(1.0, 4.0 / n))
        A[np.arange(n), np.arange(n)] += dt(2)
        for uplo in "UL":
            for trans in "NTC":
                for diag in "NU":
                    for inc in (1, -2):
                        b = strided(31, n, inc, dt); x = b.copy()
                        f77(lib, p + "trsv_", uplo, trans, diag, n, A, n + 1, x, inc)
                        T = np.triu(A[:n]) if uplo == "U" else np.tril(A[:n])
                        T = T.astype(hi)
                        if diag == "U":
                            T[np.arange(n), np.arange(n)] = 1
                        opT = T if trans == "N" else (T.T if trans == "T" else T.conj().T)
                        lx = logical(x, n, inc).astype(hi); lb = logical(b, n, inc).astype(hi)
                        resid = np.abs(opT @ lx - lb)
                        bound = 4 * n * EPS[p] * (np.abs(opT) @ np.abs(lx) + np.abs(lb))
                        assert np.all(resid <= bound + 1e-300), (p, n, uplo, trans, diag, inc, (resid / np.maximum(bound, 1e-300)).max())
                        # same answer as the netlib restatement
                        xo = b.copy(); assert oracle_call(p + "trsv", uplo, trans, diag, n, A, n + 1, xo, inc) == 0
                        assert np.allclose(x, xo, rtol=1e3 * EPS[p], atol=1e3 * EPS[p])


def test_level2_error_exits():
    lib = g.load(); seen = []
    CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t)
    cb = CB(lambda name, info, ln: seen.append((name[:6].decode(), info[0])))
    lib.b200blas_set_xerbla(cb)
    try:
        A = np.zeros((4, 4), order="F"); v = np.ones(4)
        for args, want in [(("X", 2, 2, 1.0, A, 2, v, 1, 1.0, v, 1), 1), (("N", -1, 2, 1.0, A, 2, v, 1, 1.0, v, 1), 2), (("N", 2, -1, 1.0, A, 2, v, 1, 1.0, v, 1), 3),
                           (("N", 2, 2, 1.0, A, 1, v, 1, 1.0, v, 1), 6), (("N", 2, 2, 1.0, A, 2, v, 0, 1.0, v, 1), 8), (("N", 2, 2, 1.0, A, 2, v, 1, 1.0, v, 0), 11)]:
            seen.clear(); f77(lib, "dgemv_", *args); assert seen == [("DGEMV ", want)], seen
        for args, want in [(("X", "N", "N", 2, A, 2, v, 1), 1), (("U", "X", "N", 2, A, 2, v, 1), 2), (("U", "N", "X", 2, A, 2, v, 1), 3),
                           (("U", "N", "N", -1, A, 2, v, 1), 4), (("U", "N", "N", 2, A, 1, v, 1), 6), (("U", "N", "N", 2, A, 2, v, 0), 8)]:
            seen.clear(); f77(lib, "dtrsv_", *args); assert seen == [("DTRSV ", want)], seen
        assert np.all(v == 1)
    finally:
        lib.b200blas_set_xerbla(CB(0))


def test_level1_large_device_resident():
    """BASELINE config sizes with operands resident on the device: 2^26 doubles for dot/axpy/nrm2,
    2^28 for idamax; checked through size-independent properties and against torch fp64."""
    import torch
    lib = g.load()
    n = 1 << 26
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    y = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    torch.cuda.synchronize()
    r = f77(lib, "ddot_", n, x, 1, y, 1, restype=ctypes.c_double)
    ref = torch.dot(x, y).item(); sabs = torch.dot(x.abs(), y.abs()).item()
    assert abs(r - ref) <= 2 * n * 2.0 ** -53 * sabs
    # determinism: bit-identical on repeat
    assert r == f77(lib, "ddot_", n, x, 1, y, 1, restype=ctypes.c_double)
    nr = f77(lib, "dnrm2_", n, x, 1, restype=ctypes.c_double)
    assert abs(nr - torch.linalg.norm(x).item()) <= 4 * 2.0 ** -53 * 26 * nr
    # nrm2(x)^2 == dot(x,x) to rounding
    assert abs(nr * nr - f77(lib, "ddot_", n, x, 1, x, 1, restype=ctypes.c_double)) <= 1e-12 * nr * nr
    y0 = y.clone()
    torch.cuda.synchronize()      # the library runs on its own stream: the clone must have finished reading y
    f77(lib, "daxpy_", n, 0.7, x, 1, y, 1); f77(lib, "daxpy_", n, -0.7, x, 1, y, 1)
    assert (y - y0).abs().max().item() <= 4 * 2.0 ** -53 * 2
    del y0
    n = 1 << 28
    z = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    i1, i2, i3 = 1234567, (1 << 27) + 99, n - 5
    z[i1] = 1.5; z[i2] = 1.5; z[i3] = -1.5
    torch.cuda.synchronize()
    assert f77(lib, "idamax_", n, z, 1, restype=ctypes.c_int) == i1 + 1
    z[i1] = 0.25
    torch.cuda.synchronize()
    assert f77(lib, "idamax_", n, z, 1, restype=ctypes.c_int) == i2 + 1
    z[0] = -1.5
    torch.cuda.synchronize()
    assert f77(lib, "idamax_", n, z, 1, restype=ctypes.c_int) == 1


def test_dgemv_large_device_resident():
    import torch
    lib = g.load()
    n = 8192
    gen = torch.Generator(device="cuda").manual_seed(8)
    A = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1   # row-major == column-major A^T
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    for trans, ref in [("N", A.T @ x), ("T", A @ x)]:
        f77(lib, "dgemv_", trans, n, n, 1.0, A, n, x, 1, 0.0, y, 1)
        gb = (A.abs().T if trans == "N" else A.abs()) @ x.abs()
        assert ((y - ref).abs() / (2.0 ** -53 * gb)).max().item() < 16


@pytest.mark.parametrize("p", ["s", "d"])
def test_level2_more_vs_oracle(p):
    """GER / SYR / SYMV / TRMV / ROT / ROTG through the Fortran symbols against the oracle: ragged
    sizes spanning several column chunks, odd leading dimensions, positive and negative increments, untouched padding."""
    lib = g.load(); dt = DT[p]
    tol = 64 * EPS[p]
    for (m, n) in [(1, 1), (7, 5), (300, 129), (1030, 700)]:
        for (ix, iy) in [(1, 1), (2, -3)]:
            x = strided(41, m, ix, dt); y = strided(42, n, iy, dt)
            A0 = splitmix_uniform(43, (m + 1, n), dt); A, R = A0.copy(order="F"), A0.copy(order="F")
            f77(lib, p + "ger_", m, n, 0.7, x, ix, y, iy, A, m + 1)
            assert oracle_call(p + "ger", m, n, 0.7, x, ix, y, iy, R, m + 1) == 0
            assert np.array_equal(A[m:], A0[m:]) and np.allclose(A, R, rtol=tol, atol=tol)
    for n in [1, 5, 130, 777]:
        S0 = splitmix_uniform(44, (n + 2, n), dt)
        for uplo in "UL":
            tri = np.triu(np.ones((n, n), bool)) if uplo == "U" else np.tril(np.ones((n, n), bool))
            for (ix, iy) in [(1, 1), (-2, 3)]:
                x = strided(45, n, ix, dt); y0 = strided(46, n, iy, dt)
                y, r = y0.copy(), y0.copy()
                f77(lib, p + "symv_", uplo, n, 0.7, S0, n + 2, x, ix, 1.3, y, iy)
                assert oracle_call(p + "symv", uplo, n, 0.7, S0, n + 2, x, ix, 1.3, r, iy) == 0
                assert np.allclose(y, r, rtol=tol * n, atol=tol * n)
                S, R = S0.copy(order="F"), S0.copy(order="F")
                f77(lib, p + "syr_", uplo, n, 0.7, x, ix, S, n + 2)
                assert oracle_call(p + "syr", uplo, n, 0.7, x, ix, R, n + 2) == 0
                full = np.zeros((n + 2, n), bool); full[:n] = tri
                assert np.array_equal(S[~full], S0[~full]) and np.allclose(S, R, rtol=tol, atol=tol)
                for tr in "NTC":
                    for diag in "NU":
                        xv, rv = x.copy(), x.copy()
                        f77(lib, p + "trmv_", uplo, tr, diag, n, S0, n + 2, xv, ix)
                        assert oracle_call(p + "trmv", uplo, tr, diag, n, S0, n + 2, rv, ix) == 0
                        assert np.allclose(xv, rv, rtol=tol * n, atol=tol * n), (p, n, uplo, tr, diag, ix)
        x0 = strided(47, n, 2, dt); y0 = strided(48, n, -1, dt)
        x, y, xr, yr = x0.copy(), y0.copy(), x0.copy(), y0.copy()
        f77(lib, p + "rot_", n, x, 2, y, -1, dt(0.6), dt(0.8)); oracle_call(p + "rot", n, xr, 2, yr, -1, dt(0.6), dt(0.8), restype=None)
        assert np.allclose(x, xr, rtol=tol, atol=tol) and np.allclose(y, yr, rtol=tol, atol=tol)
    for (a, b) in [(3.0, 4.0), (-3.0, 4.0), (4.0, -3.0), (0.0, 2.0), (0.0, 0.0)]:
        v1 = [np.array([v], dtype=dt) for v in (a, b, 0, 0)]; v2 = [np.array([v], dtype=dt) for v in (a, b, 0, 0)]
        f77(lib, p + "rotg_", *v1); oracle_call(p + "rotg", *v2, restype=None)
        assert np.allclose([v[0] for v in v1], [v[0] for v in v2], rtol=4 * EPS[p])


def test_pageable_operands_through_the_bounce_ring():
    """Pageable host operands of >= 4 MiB travel through pinned bounce slots packed by host threads (csrc/host_stager.cu; the
    reference's miss path is a whole-array blocking cudaMemcpy, runtime-mem.hpp:84-112).  Covered here: a flat vector larger than
    one 32 MiB slot (in and in/out), a strided in/out vector (element-wise write-back stays), and a 2-D matrix with an odd leading
    dimension that needs several slots.  Level-1 results must be bit-identical to the same calls on device-resident copies."""
    import torch
    lib = g.load()
    n = (5 << 20) + 3                                     # 40 MiB of doubles: two slot-sized chunks
    x = splitmix_uniform(91, (n,)); y0 = splitmix_uniform(92, (n,))
    xd = torch.from_numpy(x).cuda(); yd = torch.from_numpy(y0).cuda()
    y = y0.copy()
    s0 = g.stats()
    f77(lib, "daxpy_", n, -0.37, x, 1, y, 1)
    s1 = g.stats()
    f77(lib, "daxpy_", n, -0.37, xd, 1, yd, 1)
    torch.cuda.synchronize()
    assert np.array_equal(y, yd.cpu().numpy())
    assert s1["h2d_bytes"] - s0["h2d_bytes"] == 2 * n * 8 and s1["d2h_bytes"] - s0["d2h_bytes"] == n * 8
    r = f77(lib, "ddot_", n, x, 1, y, 1, restype=ctypes.c_double)
    rd = f77(lib, "ddot_", n, xd, 1, yd, 1, restype=ctypes.c_double)
    assert r == rd
    # strided in/out vector: the gaps keep their contents
    m = 700001
    v0 = splitmix_uniform(93, (1 + (m - 1) * 2,)); v = v0.copy()
    f77(lib, "dscal_", m, 3.0, v, 2)
    assert np.array_equal(v[::2], v0[::2] * 3.0) and np.array_equal(v[1::2], v0[1::2])
    # 2-D: 3000 x 2500 inside lda = 3001 (57 MiB of columns -> several slots), dgemv both ways and dger in place
    rows, cols, lda = 3000, 2500, 3001
    A0 = np.asfortranarray(splitmix_uniform(94, (lda, cols))); A = A0.copy(order="F")
    u = splitmix_uniform(95, (rows,)); w = splitmix_uniform(96, (cols,))
    Ad = torch.from_numpy(np.ascontiguousarray(A0.T)).cuda()            # row-major (cols, lda) == column-major lda x cols
    ud = torch.from_numpy(u).cuda(); wd = torch.from_numpy(w).cuda()
    for trans, nx, ny, xin, xin_d in [("N", cols, rows, w, wd), ("T", rows, cols, u, ud)]:
        out = np.zeros(ny); outd = torch.zeros(ny, dtype=torch.float64, device="cuda")
        f77(lib, "dgemv_", trans, rows, cols, 1.0, A, lda, xin, 1, 0.0, out, 1)
        f77(lib, "dgemv_", trans, rows, cols, 1.0, Ad, lda, xin_d, 1, 0.0, outd, 1)
        torch.cuda.synchronize()
        # (the staged copy has a different leading dimension, so the kernel may take another vector width: GEMV bound, not bits)
        gb = (np.abs(A0[:rows]) @ np.abs(w)) if trans == "N" else (np.abs(A0[:rows]).T @ np.abs(u))
        ref = (A0[:rows] @ w) if trans == "N" else (A0[:rows].T @ u)
        assert (np.abs(out - ref) / (2.0 ** -53 * gb)).max() < 16 and (np.abs(outd.cpu().numpy() - ref) / (2.0 ** -53 * gb)).max() < 16, trans
    f77(lib, "dger_", rows, cols, 0.5, u, 1, w, 1, A, lda)
    f77(lib, "dger_", rows, cols, 0.5, ud, 1, wd, 1, Ad, lda)
    torch.cuda.synchronize()
    got_d = Ad.cpu().numpy().T
    assert np.array_equal(A[:rows], got_d[:rows]) and np.allclose(A[:rows], A0[:rows] + 0.5 * np.outer(u, w), rtol=0, atol=4 * 2.0 ** -53)
    assert np.array_equal(A[rows:], A0[rows:])                           # the padding row of lda is not written
