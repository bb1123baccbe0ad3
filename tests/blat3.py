"""A Python restatement of the netlib Level-3 BLAS test programs the reference runs under its
interposer (tests/netlib/{s,d,c,z}blat3.f with input.{s,d,c,z}blat3, launched by
tests/netlib/test.py:28 as LD_PRELOAD=<lib> <exe> < input).  gfortran is not in this image, so the
program is restated here; each piece cites the Fortran it follows (line numbers of dblat3.f; the
s/c/z programs are the same code with the type changed, zblat3.f for the complex generator).

  Beg        DBEG (dblat3.f:2673-2718) / ZBEG (zblat3.f:3287-3340): the LCG  I = I*891 mod 1000
             (J = J*457 mod 1000 for the imaginary stream), every 5th value skipped, (I-500)/1001
  make       DMAKE (dblat3.f:2344-2456): 'GE' / 'SY' / 'TR' storage with rogue -1e10 padding in rows
             M+1..LDA and in the unreferenced triangle (and on a unit diagonal)
  mmch       DMMCH (dblat3.f:2457-2578): max_ij |ct - cc| / (eps * g), g = |alpha| sum|a||b| + |beta||c|
  chk_gemm   DCHK1 (dblat3.f:356-636)     chk_trxm  DCHK3 (dblat3.f:907-1211)
  chk_syrk   DCHK4 (dblat3.f:1212-1486)   chke      DCHKE (dblat3.f:1801-2343) for the routines built
Input values: input.dblat3:9-14 / input.zblat3:9-14  (N in {0,1,2,3,5,9}, alpha {0,1,0.7}, beta {0,1,1.3};
complex alpha {0,1,0.7-0.9i}, beta {0,1,1.3-1.1i}); threshold 16 (input.dblat3:8); NMAX = 65.

`call(name, *args)` is how the program reaches the library under test: the same Fortran-ABI call
is issued on the oracle (CPU tests) or on libb200blas.so (GPU tests).
"""
import numpy as np

NMAX = 65
IDIM = [0, 1, 2, 3, 5, 9]
THRESH = 16.0
ROGUE = -1.0e10
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": float(np.finfo(np.float32).eps) / 2, "d": float(np.finfo(np.float64).eps) / 2}
EPS["c"], EPS["z"] = EPS["s"], EPS["d"]
# netlib uses EPS = DDIFF(1+eps,1) search result = 2^-52 for d (relative machine precision as found by the program:
# halving until 1+eps == 1 gives 2^-53, then EPS = EPS+EPS = 2^-52), dblat3.f:239-245
EPS = {k: 2 * v for k, v in EPS.items()}


def alphas(p):
    return [0.0, 1.0, 0.7] if p in "sd" else [0.0 + 0j, 1.0 + 0j, 0.7 - 0.9j]


def betas(p):
    return [0.0, 1.0, 1.3] if p in "sd" else [0.0 + 0j, 1.0 + 0j, 1.3 - 1.1j]


class Beg:
    def __init__(self, cplx):
        self.cplx = cplx
        self.reset()

    def reset(self):
        self.i, self.j, self.ic = 7, 7, 0

    def __call__(self):
        self.ic += 1
        while True:
            self.i = (self.i * 891) % 1000
            self.j = (self.j * 457) % 1000
            if self.ic >= 5:
                self.ic = 0
                continue
            break
        re = (self.i - 500) / 1001.0
        return complex(re, (self.j - 500) / 1001.0) if self.cplx else re


def make(p, beg, typ, uplo, diag, m, n, lda):
    """returns (A, AA): the mathematical matrix A (m x n, zeros/symmetric filled like DMAKE does) and its
    stored form AA (lda x n, Fortran order) with rogue values where the routine must not look."""
    dt = DT[p]
    rogue = dt(ROGUE) if p in "sd" else dt(complex(ROGUE, -ROGUE))
    gen, sym, her, tri = typ == "GE", typ == "SY", typ == "HE", typ == "TR"
    upper = (sym or her or tri) and uplo == "U"
    lower = (sym or her or tri) and uplo == "L"
    unit = tri and diag == "U"
    A = np.zeros((max(m, 1), max(n, 1)), dtype=dt, order="F")
    for j in range(n):
        for i in range(m):
            if gen or (upper and i <= j) or (lower and i >= j):
                A[i, j] = dt(beg())
                if i != j:
                    if n > 3 and (j + 1) == n // 2:
                        A[i, j] = 0
                    if her:
                        A[j, i] = np.conj(A[i, j])
                    elif sym:
                        A[j, i] = A[i, j]
                    elif tri:
                        A[j, i] = 0
        if her:
            A[j, j] = A[j, j].real
        if tri:
            A[j, j] = A[j, j] + dt(1)
        if unit:
            A[j, j] = dt(1)
    AA = np.full((max(lda, 1), max(n, 1)), rogue, dtype=dt, order="F")
    if gen:
        AA[:m, :n] = A[:m, :n]
    else:
        for j in range(n):
            if upper:
                ibeg, iend = 0, (j - 1 if unit else j)
            else:
                ibeg, iend = (j + 1 if unit else j), n - 1
            AA[ibeg:iend + 1, j] = A[ibeg:iend + 1, j]
            if her and not unit:
                AA[j, j] = complex(AA[j, j].real, ROGUE)    # zblat3 ZMAKE: imaginary part of a Hermitian diagonal is rogue
    return A[:m, :n], AA


def opm(x, t):
    return x if t == "N" else (x.T if t == "T" else x.conj().T)


def mmch(p, ta, tb, alpha, A, B, beta, C, CC):
    """DMMCH: CC is the computed result; returns the test ratio."""
    hi = np.complex128 if p in "cz" else np.float64
    a, b = opm(A.astype(hi), ta), opm(B.astype(hi), tb)
    if p in "cz":
        abs1 = lambda z: np.abs(z.real) + np.abs(z.imag)     # zblat3 ABS1
    else:
        abs1 = np.abs
    ct = alpha * (a @ b) + beta * C.astype(hi)
    g = abs1(np.asarray(alpha)) * (abs1(a) @ abs1(b)) + abs1(np.asarray(beta)) * abs1(C.astype(hi))
    erri = abs1(ct - CC.astype(hi)) / EPS[p]
    nz = g != 0
    erri[nz] = erri[nz] / g[nz]
    return float(erri.max()) if erri.size else 0.0


def same_outside(typ, uplo, m, n, before, after):
    """LDERES (dblat3.f:2611-2672): everything outside the part the routine may write is bit-identical."""
    mask = np.ones(before.shape, dtype=bool)
    if typ == "GE":
        mask[:m, :n] = False
    else:
        for j in range(n):
            if uplo == "U":
                mask[: j + 1, j] = False
            else:
                mask[j:n, j] = False
    return np.array_equal(before[mask], after[mask])


def _ld(rows):
    return rows + 1 if rows < NMAX else rows


def chk_gemm(p, call):
    """DCHK1.  Returns (ncalls, errmax)."""
    beg = Beg(p in "cz"); nc, errmax = 0, 0.0
    for m in IDIM:
        for n in IDIM:
            ldc = _ld(m); null = n <= 0 or m <= 0
            for k in IDIM:
                for ta in "NTC":
                    ma, na = (k, m) if ta != "N" else (m, k)
                    lda = _ld(ma)
                    A, AA = make(p, beg, "GE", " ", " ", ma, na, lda)
                    for tb in "NTC":
                        mb, nb = (n, k) if tb != "N" else (k, n)
                        ldb = _ld(mb)
                        B, BB = make(p, beg, "GE", " ", " ", mb, nb, ldb)
                        for alpha in alphas(p):
                            for beta in betas(p):
                                C, CC = make(p, beg, "GE", " ", " ", m, n, ldc)
                                nc += 1
                                AS, BS, CS = AA.copy(order="F"), BB.copy(order="F"), CC.copy(order="F")
                                call(p + "gemm_", ta, tb, m, n, k, alpha, AA, lda, BB, ldb, beta, CC, ldc)
                                assert np.array_equal(AS, AA) and np.array_equal(BS, BB), "input operand changed"
                                if null:
                                    assert np.array_equal(CS, CC)
                                else:
                                    assert same_outside("GE", " ", m, n, CS, CC), "rogue padding of C touched"
                                    err = mmch(p, ta, tb, alpha, A, B, beta, C, CC[:m, :n])
                                    errmax = max(errmax, err)
                                    assert err < THRESH, (p, "gemm", ta, tb, m, n, k, alpha, beta, err)
    return nc, errmax


def chk_trxm(p, call, which):
    """DCHK3 for which in {'trmm','trsm'}."""
    beg = Beg(p in "cz"); nc, errmax = 0, 0.0
    one = 1.0 if p in "sd" else 1.0 + 0j
    zero = 0.0 * one
    for m in IDIM:
        for n in IDIM:
            ldb = _ld(m); null = m <= 0 or n <= 0
            for side in "LR":
                na = m if side == "L" else n
                lda = _ld(na)
                for uplo in "UL":
                    for ta in "NTC":
                        for diag in "UN":
                            for alpha in alphas(p):
                                A, AA = make(p, beg, "TR", uplo, diag, na, na, lda)
                                B, BB = make(p, beg, "GE", " ", " ", m, n, ldb)
                                nc += 1
                                AS, BS = AA.copy(order="F"), BB.copy(order="F")
                                call(p + which + "_", side, uplo, ta, diag, m, n, alpha, AA, lda, BB, ldb)
                                assert np.array_equal(AS, AA), "triangular operand changed"
                                if null:
                                    assert np.array_equal(BS, BB)
                                    continue
                                assert same_outside("GE", " ", m, n, BS, BB), "rogue padding of B touched"
                                Z = np.zeros((m, n), dtype=DT[p])
                                if which == "trmm":
                                    err = mmch(p, ta, "N", alpha, A, B, zero, Z, BB[:m, :n]) if side == "L" else \
                                        mmch(p, "N", ta, alpha, B, A, zero, Z, BB[:m, :n])
                                else:
                                    # the computed X must reproduce alpha*B:  op(A) X  (or X op(A))  vs  alpha*B
                                    X = BB[:m, :n].copy(); aB = (alpha * B.astype(np.complex128 if p in "cz" else np.float64)).astype(DT[p])
                                    err = mmch(p, ta, "N", one, A, X, zero, Z, aB) if side == "L" else \
                                        mmch(p, "N", ta, one, X, A, zero, Z, aB)
                                errmax = max(errmax, err)
                                assert err < THRESH, (p, which, side, uplo, ta, diag, m, n, alpha, err)
    return nc, errmax


def chk_syrk(p, call):
    """DCHK4 (SYRK part): only the referenced triangle is computed and compared."""
    beg = Beg(p in "cz"); nc, errmax = 0, 0.0
    for n in IDIM:
        ldc = _ld(n)
        for k in IDIM:
            for trans in ("NTC" if p in "sd" else "NT"):
                ma, na = (k, n) if trans != "N" else (n, k)
                lda = _ld(ma)
                A, AA = make(p, beg, "GE", " ", " ", ma, na, lda)
                for uplo in "UL":
                    for alpha in alphas(p):
                        for beta in betas(p):
                            C, CC = make(p, beg, "SY", uplo, " ", n, n, ldc)
                            nc += 1
                            AS, CS = AA.copy(order="F"), CC.copy(order="F")
                            call(p + "syrk_", uplo, trans, n, k, alpha, AA, lda, beta, CC, ldc)
                            assert np.array_equal(AS, AA)
                            if n <= 0:
                                assert np.array_equal(CS, CC)
                                continue
                            assert same_outside("SY", uplo, n, n, CS, CC), "unreferenced triangle / padding of C touched"
                            tt = "T" if trans != "N" else "N"
                            for j in range(n):
                                rows = slice(0, j + 1) if uplo == "U" else slice(j, n)
                                # column j of C against op(A) rows x op(A)^T column j
                                a = opm(A, tt)                       # n x k
                                err = mmch(p, "N", "T", alpha, a[rows, :], a[j:j + 1, :], beta, C[rows, j:j + 1], CC[rows, j:j + 1])
                                errmax = max(errmax, err)
                                assert err < THRESH, (p, "syrk", uplo, trans, n, k, alpha, beta, j, err)
    return nc, errmax


def chk_symm(p, call, which="symm"):
    """DCHK2 (dblat3.f:637-906) / ZCHK2: SYMM and, for complex types, HEMM."""
    beg = Beg(p in "cz"); nc, errmax = 0, 0.0
    typ = "HE" if which == "hemm" else "SY"
    for m in IDIM:
        for n in IDIM:
            ldc = _ld(m); ldb = _ld(m); null = n <= 0 or m <= 0
            B, BB = make(p, beg, "GE", " ", " ", m, n, ldb)
            for side in "LR":
                na = m if side == "L" else n
                lda = _ld(na)
                for uplo in "UL":
                    A, AA = make(p, beg, typ, uplo, " ", na, na, lda)
                    for alpha in alphas(p):
                        for beta in betas(p):
                            C, CC = make(p, beg, "GE", " ", " ", m, n, ldc)
                            nc += 1
                            AS, BS, CS = AA.copy(order="F"), BB.copy(order="F"), CC.copy(order="F")
                            call(p + which + "_", side, uplo, m, n, alpha, AA, lda, BB, ldb, beta, CC, ldc)
                            assert np.array_equal(AS, AA) and np.array_equal(BS, BB), "input operand changed"
                            if null:
                                assert np.array_equal(CS, CC)
                                continue
                            assert same_outside("GE", " ", m, n, CS, CC), "rogue padding of C touched"
                            err = mmch(p, "N", "N", alpha, A, B, beta, C, CC[:m, :n]) if side == "L" else \
                                mmch(p, "N", "N", alpha, B, A, beta, C, CC[:m, :n])
                            errmax = max(errmax, err)
                            assert err < THRESH, (p, which, side, uplo, m, n, alpha, beta, err)
    return nc, errmax


def _rk_scalars(p, herm):
    """alpha/beta sets of DCHK4/DCHK5; the Hermitian routines take real beta (and HERK a real alpha): the real parts
    of the complex input values, as ZCHK4/ZCHK5 do (RALPHA = DBLE(ALPHA), RBETA = DBLE(BETA))."""
    return alphas(p), betas(p)


def chk_herk(p, call):
    """ZCHK4 (zblat3.f), HERK part: real alpha/beta, trans in {N,C}, imaginary diagonal of the result is zero."""
    assert p in "cz"
    beg = Beg(True); nc, errmax = 0, 0.0
    rt = np.float32 if p == "c" else np.float64
    for n in IDIM:
        ldc = _ld(n)
        for k in IDIM:
            for trans in "NC":
                ma, na = (k, n) if trans != "N" else (n, k)
                lda = _ld(ma)
                A, AA = make(p, beg, "GE", " ", " ", ma, na, lda)
                for uplo in "UL":
                    for alpha in alphas(p):
                        ralpha = rt(alpha.real)
                        for beta in betas(p):
                            rbeta = rt(beta.real)
                            C, CC = make(p, beg, "HE", uplo, " ", n, n, ldc)
                            nc += 1
                            AS, CS = AA.copy(order="F"), CC.copy(order="F")
                            call(p + "herk_", uplo, trans, n, k, ralpha, AA, lda, rbeta, CC, ldc)
                            assert np.array_equal(AS, AA)
                            if n <= 0:
                                assert np.array_equal(CS, CC)
                                continue
                            null = (ralpha == 0 or k <= 0) and rbeta == 1
                            assert same_outside("SY", uplo, n, n, CS, CC), "unreferenced triangle / padding of C touched"
                            if null:
                                assert np.array_equal(CS, CC)
                                continue
                            a = A if trans == "N" else A.conj().T            # n x k
                            for j in range(n):
                                rows = slice(0, j + 1) if uplo == "U" else slice(j, n)
                                err = mmch(p, "N", "C", complex(ralpha), a[rows, :], a[j:j + 1, :], complex(rbeta), C[rows, j:j + 1], CC[rows, j:j + 1])
                                errmax = max(errmax, err)
                                assert err < THRESH, (p, "herk", uplo, trans, n, k, ralpha, rbeta, j, err)
                                assert CC[j, j].imag == 0
    return nc, errmax


def chk_r2k(p, call, which="syr2k"):
    """DCHK5 (dblat3.f:1487-1800) / ZCHK5: SYR2K and, for complex types, HER2K (real beta, conj(alpha) on the second term)."""
    herm = which == "her2k"
    beg = Beg(p in "cz"); nc, errmax = 0, 0.0
    hi = np.complex128 if p in "cz" else np.float64
    rt = np.float32 if p in "sc" else np.float64
    transes = "NC" if herm else ("NTC" if p in "sd" else "NT")
    for n in IDIM:
        ldc = _ld(n)
        for k in IDIM:
            for trans in transes:
                ma, na = (k, n) if trans != "N" else (n, k)
                lda = _ld(ma)
                A, AA = make(p, beg, "GE", " ", " ", ma, na, lda)
                B, BB = make(p, beg, "GE", " ", " ", ma, na, lda)
                for uplo in "UL":
                    for alpha in alphas(p):
                        for beta in betas(p):
                            bet = rt(beta.real) if herm else beta
                            C, CC = make(p, beg, "HE" if herm else "SY", uplo, " ", n, n, ldc)
                            nc += 1
                            AS, BS, CS = AA.copy(order="F"), BB.copy(order="F"), CC.copy(order="F")
                            call(p + which + "_", uplo, trans, n, k, alpha, AA, lda, BB, lda, bet, CC, ldc)
                            assert np.array_equal(AS, AA) and np.array_equal(BS, BB)
                            if n <= 0:
                                assert np.array_equal(CS, CC)
                                continue
                            assert same_outside("SY", uplo, n, n, CS, CC), "unreferenced triangle / padding of C touched"
                            null = (alpha == 0 or k <= 0) and bet == 1
                            if null:
                                assert np.array_equal(CS, CC)
                                continue
                            # C = alpha a b^X + alpha' b a^X + beta C  ==  [alpha a | alpha' b] [b | a]^X   (X = T or H), the DCHK5 trick
                            a = (A if trans == "N" else (A.conj().T if herm else A.T)).astype(hi)      # n x k
                            b = (B if trans == "N" else (B.conj().T if herm else B.T)).astype(hi)
                            al2 = np.conj(alpha) if herm else alpha
                            W = np.concatenate([alpha * a, al2 * b], axis=1)                           # n x 2k
                            Z = np.concatenate([b, a], axis=1)
                            one = 1.0 + 0j if p in "cz" else 1.0
                            for j in range(n):
                                rows = slice(0, j + 1) if uplo == "U" else slice(j, n)
                                err = mmch(p, "N", "C" if herm else "T", one, W[rows, :], Z[j:j + 1, :], complex(bet) if p in "cz" else bet,
                                           C[rows, j:j + 1], CC[rows, j:j + 1])
                                errmax = max(errmax, err)
                                assert err < THRESH, (p, which, uplo, trans, n, k, alpha, beta, j, err)
                                if herm:
                                    assert CC[j, j].imag == 0
    return nc, errmax


def chke(p, call_capture):
    """DCHKE / ZCHKE restated: every illegal argument must report its INFO through XERBLA under the routine's SRNAME and
    touch nothing.  `call_capture(name, *args)` performs the call and returns the list of (srname, info) pairs XERBLA saw.
    Returns the number of checks: 162 for s/d (netlib DCHKE: 162), 291 for c/z (netlib ZCHKE's 288 + '/' as the illegal TRANS of
    ?SYRK, ?SYR2K, ?HER2K, where netlib only tries the letter that is legal for the sister routine)."""
    dt = DT[p]
    one = 1.0 if p in "sd" else 1.0 + 0j
    A = np.zeros((2, 2), dtype=dt, order="F"); B = np.zeros((2, 2), dtype=dt, order="F"); C = np.zeros((2, 2), dtype=dt, order="F")
    n_checked = 0

    def expect(name, info, *args):
        nonlocal n_checked
        seen = call_capture(name, *args)
        assert seen == [((name[:-1]).upper().ljust(6), info)], (name, info, seen)
        n_checked += 1

    # netlib ?CHKE walks every legal transpose letter: N, T for the real programs (dblat3.f DCHKE), N, C, T for the complex
    # ones (zblat3.f ZCHKE) -- 162 checks in s/d, 288 in c/z
    ops = "NT" if p in "sd" else "NCT"
    g = p + "gemm_"
    for tb in ops:
        expect(g, 1, "/", tb, 0, 0, 0, one, A, 1, B, 1, one, C, 1)
    for ta in ops:
        expect(g, 2, ta, "/", 0, 0, 0, one, A, 1, B, 1, one, C, 1)
    for ta in ops:
        for tb in ops:
            expect(g, 3, ta, tb, -1, 0, 0, one, A, 1, B, 1, one, C, 1)
            expect(g, 4, ta, tb, 0, -1, 0, one, A, 1, B, 1, one, C, 1)
            expect(g, 5, ta, tb, 0, 0, -1, one, A, 1, B, 1, one, C, 1)
            # LDA: op(A) is m x k -- 'N' stores m rows, otherwise k rows
            if ta == "N":
                expect(g, 8, ta, tb, 2, 0, 0, one, A, 1, B, 1, one, C, 2)
            else:
                expect(g, 8, ta, tb, 0, 0, 2, one, A, 1, B, 2 if tb == "N" else 1, one, C, 1)
            # LDB: op(B) is k x n -- 'N' stores k rows, otherwise n rows
            if tb == "N":
                expect(g, 10, ta, tb, 0, 0, 2, one, A, 1 if ta == "N" else 2, B, 1, one, C, 1)
            else:
                expect(g, 10, ta, tb, 0, 2, 0, one, A, 1, B, 1, one, C, 1)
            expect(g, 13, ta, tb, 2, 0, 0, one, A, 2 if ta == "N" else 1, B, 1, one, C, 1)
    for r in ("trmm_", "trsm_"):
        t = p + r
        expect(t, 1, "/", "U", "N", "N", 0, 0, one, A, 1, B, 1)
        expect(t, 2, "L", "/", "N", "N", 0, 0, one, A, 1, B, 1)
        expect(t, 3, "L", "U", "/", "N", 0, 0, one, A, 1, B, 1)
        expect(t, 4, "L", "U", "N", "/", 0, 0, one, A, 1, B, 1)
        for side in "LR":
            for uplo in "UL":
                for ta in ops:
                    expect(t, 5, side, uplo, ta, "N", -1, 0, one, A, 1, B, 1)
                    expect(t, 6, side, uplo, ta, "N", 0, -1, one, A, 1, B, 1)
        for uplo in "UL":
            for ta in ops:
                expect(t, 9, "L", uplo, ta, "N", 2, 0, one, A, 1, B, 2)
                expect(t, 9, "R", uplo, ta, "N", 0, 2, one, A, 1, B, 1)
                expect(t, 11, "L", uplo, ta, "N", 2, 0, one, A, 2, B, 1)
                expect(t, 11, "R", uplo, ta, "N", 2, 0, one, A, 1, B, 1)
    s = p + "syrk_"
    expect(s, 1, "/", "N", 0, 0, one, A, 1, one, C, 1)
    expect(s, 2, "U", "/", 0, 0, one, A, 1, one, C, 1)
    if p in "cz":
        expect(s, 2, "U", "C", 0, 0, one, A, 1, one, C, 1)      # zblat3 ZCHKE: 'C' is illegal for ZSYRK
    for uplo in "UL":
        for tr in "NT":
            expect(s, 3, uplo, tr, -1, 0, one, A, 1, one, C, 1)
            expect(s, 4, uplo, tr, 0, -1, one, A, 1, one, C, 1)
        expect(s, 7, uplo, "N", 2, 0, one, A, 1, one, C, 2)
        expect(s, 7, uplo, "T", 0, 2, one, A, 1, one, C, 1)
        expect(s, 10, uplo, "N", 2, 0, one, A, 2, one, C, 1)
        expect(s, 10, uplo, "T", 2, 0, one, A, 1, one, C, 1)
    # ---- DCHKE/ZCHKE blocks for SYMM/HEMM, SYR2K/HER2K, HERK ----
    rt = np.float32 if p in "sc" else np.float64
    for r in (("symm_", "hemm_") if p in "cz" else ("symm_",)):
        t = p + r
        expect(t, 1, "/", "U", 0, 0, one, A, 1, B, 1, one, C, 1)
        expect(t, 2, "L", "/", 0, 0, one, A, 1, B, 1, one, C, 1)
        for side in "LR":
            for uplo in "UL":
                expect(t, 3, side, uplo, -1, 0, one, A, 1, B, 1, one, C, 1)
                expect(t, 4, side, uplo, 0, -1, one, A, 1, B, 1, one, C, 1)
        for uplo in "UL":
            expect(t, 7, "L", uplo, 2, 0, one, A, 1, B, 2, one, C, 2)
            expect(t, 7, "R", uplo, 0, 2, one, A, 1, B, 1, one, C, 1)
            expect(t, 9, "L", uplo, 2, 0, one, A, 2, B, 1, one, C, 2)
            expect(t, 9, "R", uplo, 2, 0, one, A, 1, B, 1, one, C, 2)
            expect(t, 12, "L", uplo, 2, 0, one, A, 2, B, 2, one, C, 1)
            expect(t, 12, "R", uplo, 2, 0, one, A, 1, B, 2, one, C, 1)
    for r, tr2, bet in ((("syr2k_", "T", one),) + ((("her2k_", "C", rt(1)),) if p in "cz" else ())):
        t = p + r
        expect(t, 1, "/", "N", 0, 0, one, A, 1, B, 1, bet, C, 1)
        expect(t, 2, "U", "/", 0, 0, one, A, 1, B, 1, bet, C, 1)
        if p in "cz":
            expect(t, 2, "U", "C" if r == "syr2k_" else "T", 0, 0, one, A, 1, B, 1, bet, C, 1)
        for uplo in "UL":
            for tr in ("N", tr2):
                expect(t, 3, uplo, tr, -1, 0, one, A, 1, B, 1, bet, C, 1)
                expect(t, 4, uplo, tr, 0, -1, one, A, 1, B, 1, bet, C, 1)
            expect(t, 7, uplo, "N", 2, 0, one, A, 1, B, 1, bet, C, 2)
            expect(t, 7, uplo, tr2, 0, 2, one, A, 1, B, 1, bet, C, 1)
            expect(t, 9, uplo, "N", 2, 0, one, A, 2, B, 1, bet, C, 2)
            expect(t, 9, uplo, tr2, 0, 2, one, A, 2, B, 1, bet, C, 1)
            expect(t, 12, uplo, "N", 2, 0, one, A, 2, B, 2, bet, C, 1)
            expect(t, 12, uplo, tr2, 2, 0, one, A, 1, B, 1, bet, C, 1)
    if p in "cz":
        t = p + "herk_"
        expect(t, 1, "/", "N", 0, 0, rt(1), A, 1, rt(1), C, 1)
        expect(t, 2, "U", "T", 0, 0, rt(1), A, 1, rt(1), C, 1)
        for uplo in "UL":
            for tr in "NC":
                expect(t, 3, uplo, tr, -1, 0, rt(1), A, 1, rt(1), C, 1)
                expect(t, 4, uplo, tr, 0, -1, rt(1), A, 1, rt(1), C, 1)
            expect(t, 7, uplo, "N", 2, 0, rt(1), A, 1, rt(1), C, 2)
            expect(t, 7, uplo, "C", 0, 2, rt(1), A, 1, rt(1), C, 1)
            expect(t, 10, uplo, "N", 2, 0, rt(1), A, 2, rt(1), C, 1)
            expect(t, 10, uplo, "C", 2, 0, rt(1), A, 1, rt(1), C, 1)
    assert not A.any() and not B.any() and not C.any()
    return n_checked
