"""Shared plumbing of the banded / packed / Hermitian Level-2 tests.

  * storage builders: dense logical matrix -> band / packed / full arrays, column-major (netlib) and row-major (CBLAS)
  * a numpy model of every routine on the dense logical matrix (float64 / complex128) -- the expectation for the
    row-major CBLAS forms, which the Fortran-style oracle does not cover
  * one case list used by every leg, so the oracle-vs-OpenBLAS pin, the CPU emulation of the kernels' index logic and the
    GPU parity tests all run the *same* calls.

Nothing here is imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

from helpers import ROOT, splitmix_uniform

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": 2.0 ** -24, "d": 2.0 ** -53, "c": 2.0 ** -24, "z": 2.0 ** -53}
CREAL = {"s": ctypes.c_float, "d": ctypes.c_double, "c": ctypes.c_float, "z": ctypes.c_double}
K_FULL_GEN, K_FULL_TRI, K_BAND_GEN, K_BAND_TRI, K_PACKED = 0, 1, 2, 3, 4          # structured.cuh Kind
R_GERU, R_GERC, R_SYR2, R_HER, R_HER2, R_SYR = 0, 1, 2, 3, 4, 5                    # structured.cuh RankMode
ROGUE = -1.0e10     # netlib xBLAT2 fills unreferenced storage with a rogue value; here it also proves it is never read


def cplx(p):
    return p in "cz"


def rnd(seed, shape, p):
    return splitmix_uniform(seed, shape, DT[p])


def vec(seed, n, inc, p):
    return rnd(seed, (1 + (max(n, 1) - 1) * abs(inc),), p)


def logical(v, n, inc):
    e = v[:: abs(inc)][:n]
    return e if inc > 0 else e[::-1]


# ---------------------------------------------------------------- storage builders
def band_gen(M, kl, ku, lda, rowmajor=False):
    """GBMV storage.  column-major (netlib): AB[ku+i-j, j]; row-major (CBLAS): a[i*lda + kl+j-i]."""
    m, n = M.shape
    if not rowmajor:
        ab = np.full((lda, n), ROGUE, dtype=M.dtype, order="F")
        for j in range(n):
            for i in range(max(0, j - ku), min(m, j + kl + 1)):
                ab[ku + i - j, j] = M[i, j]
        return ab
    ab = np.full((m, lda), ROGUE, dtype=M.dtype, order="C")
    for i in range(m):
        for j in range(max(0, i - kl), min(n, i + ku + 1)):
            ab[i, kl + j - i] = M[i, j]
    return ab


def band_tri(M, uplo, k, lda, rowmajor=False):
    """SBMV/HBMV/TBMV/TBSV storage of one triangle with k off-diagonals."""
    n = M.shape[0]
    if not rowmajor:
        ab = np.full((lda, n), ROGUE, dtype=M.dtype, order="F")
        for j in range(n):
            rows = range(max(0, j - k), j + 1) if uplo == "U" else range(j, min(n, j + k + 1))
            for i in rows:
                ab[(k + i - j) if uplo == "U" else (i - j), j] = M[i, j]
        return ab
    ab = np.full((n, lda), ROGUE, dtype=M.dtype, order="C")
    for i in range(n):
        cols = range(i, min(n, i + k + 1)) if uplo == "U" else range(max(0, i - k), i + 1)
        for j in cols:
            ab[i, (j - i) if uplo == "U" else (k + j - i)] = M[i, j]
    return ab


def packed(M, uplo, rowmajor=False):
    n = M.shape[0]
    if not rowmajor:
        cols = [M[: j + 1, j] if uplo == "U" else M[j:, j] for j in range(n)]
    else:
        cols = [M[i, i:] if uplo == "U" else M[i, : i + 1] for i in range(n)]
    return np.ascontiguousarray(np.concatenate(cols)) if n else np.zeros(0, M.dtype)


def full_tri(M, uplo, lda, rowmajor=False):
    """full storage with only one triangle meaningful (the other is rogue)"""
    n = M.shape[0]
    keep = np.triu(np.ones((n, n), bool)) if uplo == "U" else np.tril(np.ones((n, n), bool))
    D = np.where(keep, M, M.dtype.type(ROGUE))
    if not rowmajor:
        a = np.full((lda, n), ROGUE, dtype=M.dtype, order="F"); a[:n, :] = D
    else:
        a = np.full((n, lda), ROGUE, dtype=M.dtype, order="C"); a[:, :n] = D
    return a


def tri_mask(n, uplo, k=None):
    i, j = np.indices((n, n))
    m = (i <= j) if uplo == "U" else (i >= j)
    if k is not None:
        m &= np.abs(i - j) <= k
    return m


def sym_from_tri(T, uplo, herm):
    """dense symmetric / Hermitian matrix whose `uplo` triangle is T's (Hermitian: real diagonal)"""
    n = T.shape[0]
    S = np.where(tri_mask(n, uplo), T, 0)
    strict = S - np.diag(np.diag(S))
    d = np.diag(np.diag(S).real.astype(T.dtype)) if herm else np.diag(np.diag(S))
    return strict + (strict.conj().T if herm else strict.T) + d


def opmat(M, trans):
    return M if trans == "N" else (M.T if trans == "T" else M.conj().T)


def wide(p):
    return np.complex128 if cplx(p) else np.float64


# ---------------------------------------------------------------- the emulation harness (tests/drivers/struct_emul.cpp)
_emu = None


def load_emul():
    global _emu
    if _emu is None:
        src = os.path.join(ROOT, "tests", "drivers", "struct_emul.cpp")
        hdr = os.path.join(ROOT, "libgpublas_b200", "csrc", "structured.cuh")
        out = os.path.join(ROOT, "tests", "drivers", "_build", "libstruct_emul.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            # -ffp-contract=off: plain mul+add on the host; the GPU contracts to FMA, both within the stated tolerances
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-ffp-contract=off", "-I/usr/local/cuda/include",
                                   "-o", out, src])
        _emu = ctypes.CDLL(out)
    return _emu


def _sp(p, v):
    """pointer to a scalar of precision p"""
    r = CREAL[p]
    if cplx(p):
        v = complex(v)
        return (r * 2)(v.real, v.imag)
    return (r * 1)(float(v))


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def _ch(c):
    return ctypes.c_char(c.encode())


def emu_call(name, *args, rowmajor=False):
    """Run Fortran-style routine `name` (e.g. 'zhbmv') with the argument list of its Fortran form through the CPU
    emulation of the kernels.  rowmajor=True gives the CBLAS row-major meaning to the same argument list."""
    lib = load_emul()
    p, r = name[0], name[1:]
    rmj = int(rowmajor)
    fn = lambda suffix: getattr(lib, "emu_" + p + suffix)
    if r == "gbmv":
        tr, m, n, kl, ku, al, a, lda, x, ix, be, y, iy = args
        fn("gbmv")(rmj, _ch(tr), m, n, kl, ku, _sp(p, al), _ptr(a), lda, _ptr(x), ix, _sp(p, be), _ptr(y), iy)
    elif r in ("sbmv", "hbmv"):
        ul, n, k, al, a, lda, x, ix, be, y, iy = args
        fn("symv_like")(K_BAND_TRI, int(r[0] == "h"), rmj, _ch(ul), n, k, _sp(p, al), _ptr(a), lda, _ptr(x), ix, _sp(p, be), _ptr(y), iy)
    elif r in ("spmv", "hpmv"):
        ul, n, al, a, x, ix, be, y, iy = args
        fn("symv_like")(K_PACKED, int(r[0] == "h"), rmj, _ch(ul), n, 0, _sp(p, al), _ptr(a), 1, _ptr(x), ix, _sp(p, be), _ptr(y), iy)
    elif r in ("hemv", "symv"):
        ul, n, al, a, lda, x, ix, be, y, iy = args
        fn("symv_like")(K_FULL_TRI, int(r == "hemv"), rmj, _ch(ul), n, 0, _sp(p, al), _ptr(a), lda, _ptr(x), ix, _sp(p, be), _ptr(y), iy)
    elif r in ("tbmv", "tbsv"):
        ul, tr, dg, n, k, a, lda, x, ix = args
        fn("tri")(K_BAND_TRI, int(r == "tbsv"), rmj, _ch(ul), _ch(tr), _ch(dg), n, k, _ptr(a), lda, _ptr(x), ix)
    elif r in ("tpmv", "tpsv"):
        ul, tr, dg, n, a, x, ix = args
        fn("tri")(K_PACKED, int(r == "tpsv"), rmj, _ch(ul), _ch(tr), _ch(dg), n, 0, _ptr(a), 1, _ptr(x), ix)
    elif r in ("trmv", "trsv"):
        ul, tr, dg, n, a, lda, x, ix = args
        fn("tri")(K_FULL_TRI, int(r == "trsv"), rmj, _ch(ul), _ch(tr), _ch(dg), n, 0, _ptr(a), lda, _ptr(x), ix)
    elif r in ("geru", "gerc", "ger"):
        m, n, al, x, ix, y, iy, a, lda = args
        fn("ger")(int(r == "gerc"), rmj, m, n, _sp(p, al), _ptr(x), ix, _ptr(y), iy, _ptr(a), lda)
    elif r in ("syr2", "her2"):
        ul, n, al, x, ix, y, iy, a, lda = args
        fn("rank_sym")(K_FULL_TRI, R_SYR2 if r == "syr2" else R_HER2, rmj, _ch(ul), n, _sp(p, al), _ptr(x), ix, _ptr(y), iy, _ptr(a), lda)
    elif r in ("her", "syr"):
        ul, n, al, x, ix, a, lda = args
        fn("rank_sym")(K_FULL_TRI, R_HER if r == "her" else R_SYR, rmj, _ch(ul), n, _sp(p, al), _ptr(x), ix, _ptr(None), 1, _ptr(a), lda)
    elif r in ("spr", "hpr"):
        ul, n, al, x, ix, a = args
        fn("rank_sym")(K_PACKED, R_SYR if r == "spr" else R_HER, rmj, _ch(ul), n, _sp(p, al), _ptr(x), ix, _ptr(None), 1, _ptr(a), 1)
    elif r in ("spr2", "hpr2"):
        ul, n, al, x, ix, y, iy, a = args
        fn("rank_sym")(K_PACKED, R_SYR2 if r == "spr2" else R_HER2, rmj, _ch(ul), n, _sp(p, al), _ptr(x), ix, _ptr(y), iy, _ptr(a), 1)
    else:
        raise KeyError(name)


# ---------------------------------------------------------------- CBLAS caller (product library)
CBLAS_ENUM = {"R": 101, "Cm": 102, "N": 111, "T": 112, "C": 113, "U": 121, "L": 122, "dN": 131, "dU": 132}


def cblas_call(lib, name, order, *args):
    """Call cblas_<name> with the Fortran-form argument list: characters become enums (uplo / trans / diag by position in
    the routine's signature), real scalars go by value, complex scalars by pointer."""
    p, r = name[0], name[1:]
    real = CREAL[p]
    fn = getattr(lib, "cblas_" + name)
    fn.restype = None
    cargs, keep = [ctypes.c_int(CBLAS_ENUM["R" if order == "R" else "Cm"])], []
    chars = [a for a in args if isinstance(a, str)]
    # character roles by routine family
    if r == "gbmv":
        roles = ["trans"]
    elif r in ("tbmv", "tbsv", "tpmv", "tpsv", "trmv", "trsv"):
        roles = ["uplo", "trans", "diag"]
    elif r in ("geru", "gerc", "ger"):
        roles = []
    else:
        roles = ["uplo"]
    assert len(chars) == len(roles)
    ci = 0
    for a in args:
        if isinstance(a, str):
            role = roles[ci]; ci += 1
            cargs.append(ctypes.c_int(CBLAS_ENUM[("d" + a) if role == "diag" else a]))
        elif isinstance(a, (bool, int, np.integer)):
            cargs.append(ctypes.c_int(int(a)))
        elif isinstance(a, float):
            cargs.append(real(a))                       # real scalar by value (also the real alpha of her / hpr)
        elif isinstance(a, complex):
            c = (real * 2)(a.real, a.imag); keep.append(c); cargs.append(ctypes.byref(c))
        else:
            cargs.append(_ptr(a))
    fn(*cargs)


# ---------------------------------------------------------------- case list
def scal(p, v):
    """a scalar literal of the routine's type: complex for c/z"""
    return complex(v) if cplx(p) else float(np.real(v))


ALPHA = {"s": 0.7, "d": 0.7, "c": 0.7 - 0.9j, "z": 0.7 - 0.9j}       # the netlib testers' values (input.dblat3 / input.zblat3)
BETA = {"s": 1.3, "d": 1.3, "c": 1.3 - 1.1j, "z": 1.3 - 1.1j}


def well_conditioned_tri(seed, n, p):
    """dense n x n whose triangles both give well-conditioned triangular systems"""
    M = rnd(seed, (n, n), p).astype(wide(p))
    M = M / max(n, 1)
    M[np.arange(n), np.arange(n)] = 2.0 + np.abs(rnd(seed + 1, (n,), p).real)
    if cplx(p):
        M[np.arange(n), np.arange(n)] = M[np.arange(n), np.arange(n)] * np.exp(1j * 0.3)
    return M.astype(DT[p])


class Case:
    """One call: Fortran-form argument list, index of the output array in it, and the model's expectation for that array."""

    def __init__(self, name, args, out, expect, tol, tag):
        self.name, self.args, self.out, self.expect, self.tol, self.tag = name, args, out, expect, tol, tag

    def fresh_args(self):
        return [a.copy(order="K") if isinstance(a, np.ndarray) else a for a in self.args]


def _put(v, n, inc, vals):
    """copy of strided vector v with its n logical elements replaced"""
    out = v.copy()
    idx = np.arange(n) * abs(inc)
    out[idx if inc > 0 else idx[::-1]] = vals
    return out


def cases(p, rowmajor=False, sizes=(1, 2, 5, 33, 70), big=()):
    """Every routine of level2_struct.cu for precision p.  Column-major cases follow the netlib ?BLAT2 recipe (rogue padding,
    LDA = rows + 1, ragged n, k / kl / ku from 0 to 'everything', positive and negative increments, alpha = 0 and
    beta = 0 / 1 corner cases); row-major cases build the CBLAS row-major storage explicitly."""
    dt, W, eps = DT[p], wide(p), EPS[p]
    al, be = ALPHA[p], BETA[p]
    is_c = cplx(p)
    out = []
    transs = "NTC"
    incs = [(1, 1), (2, -3), (-1, 2)]
    seed = [100]

    def nxt():
        seed[0] += 7
        return seed[0]

    def add(name, args, oi, expect, n, tag):
        out.append(Case(p + name, args, oi, np.asarray(expect).astype(dt), 16 * eps * max(n, 4), "%s%s %s" % (p, name, tag)))

    # ---- GBMV
    shapes = [(1, 1, 0, 0), (7, 5, 2, 1), (5, 7, 0, 3), (33, 40, 5, 7), (40, 33, 39, 32)] + [(n, n + 3, n // 3, 2) for n in big]
    for (m, n, kl, ku) in shapes:
        M = rnd(nxt(), (m, n), p).astype(W)
        i, j = np.indices((m, n)); M = np.where((i - j <= kl) & (j - i <= ku), M, 0)
        lda = kl + ku + 2
        A = band_gen(M.astype(dt), kl, ku, lda, rowmajor)
        for tr in transs:
            for (ix, iy) in incs[: (3 if m < 20 else 1)]:
                for (a_, b_) in ([(al, be), (0.0, be), (al, 0.0), (al, 1.0)] if m < 10 else [(al, be)]):
                    lx, ly = (n, m) if tr == "N" else (m, n)
                    x = vec(nxt(), lx, ix, p); y = vec(nxt(), ly, iy, p)
                    r = scal(p, a_) * (opmat(M, tr) @ logical(x, lx, ix).astype(W)) + (scal(p, b_) * logical(y, ly, iy).astype(W) if b_ != 0 else 0)
                    add("gbmv", [tr, m, n, kl, ku, scal(p, a_), A, lda, x, ix, scal(p, b_), y, iy], 11, _put(y, ly, iy, r), max(m, n),
                        "%s m=%d n=%d kl=%d ku=%d inc=%d,%d a=%s b=%s" % (tr, m, n, kl, ku, ix, iy, a_, b_))
    for n in list(sizes) + list(big):
        ks = sorted({0, 1, min(4, max(n - 1, 0)), max(n - 1, 0)})
        for ul in "UL":
            T = rnd(nxt(), (n, n), p).astype(W)
            x0 = {inc: vec(nxt(), n, inc, p) for inc in (1, 2, -1, -3)}
            # ---- SBMV / HBMV
            for k in ks:
                S = sym_from_tri(np.where(tri_mask(n, ul, k), T, 0), ul, is_c)
                lda = k + 2
                A = band_tri(np.where(tri_mask(n, ul, k), T, 0).astype(dt), ul, k, lda, rowmajor)
                for (ix, iy) in incs[: (2 if n < 40 else 1)]:
                    for (a_, b_) in ([(al, be), (0.0, be), (al, 0.0)] if n <= 5 else [(al, be)]):
                        x = x0[ix]; y = vec(nxt(), n, iy, p)
                        r = scal(p, a_) * (S @ logical(x, n, ix).astype(W)) + (scal(p, b_) * logical(y, n, iy).astype(W) if b_ != 0 else 0)
                        add("hbmv" if is_c else "sbmv", [ul, n, k, scal(p, a_), A, lda, x, ix, scal(p, b_), y, iy], 9, _put(y, n, iy, r), n,
                            "%s n=%d k=%d inc=%d,%d a=%s b=%s" % (ul, n, k, ix, iy, a_, b_))
            # ---- SPMV / HPMV / HEMV
            Tt = np.where(tri_mask(n, ul), T, 0)
            S = sym_from_tri(Tt, ul, is_c)
            AP = packed(Tt.astype(dt), ul, rowmajor)
            AF = full_tri(Tt.astype(dt), ul, n + 1, rowmajor)
            for (ix, iy) in incs[: (3 if n < 40 else 1)]:
                x = x0[ix]; y = vec(nxt(), n, iy, p)
                r = scal(p, al) * (S @ logical(x, n, ix).astype(W)) + scal(p, be) * logical(y, n, iy).astype(W)
                add("hpmv" if is_c else "spmv", [ul, n, scal(p, al), AP, x, ix, scal(p, be), y, iy], 7, _put(y, n, iy, r), n, "%s n=%d inc=%d,%d" % (ul, n, ix, iy))
                add("hemv" if is_c else "symv", [ul, n, scal(p, al), AF, n + 1, x, ix, scal(p, be), y, iy], 8, _put(y, n, iy, r), n, "%s n=%d inc=%d,%d" % (ul, n, ix, iy))
            # ---- rank updates on the stored triangle
            for (ix, iy) in incs[: (3 if n < 40 else 1)]:
                x = x0[ix]; y = vec(nxt(), n, iy, p)
                xl, yl = logical(x, n, ix).astype(W), logical(y, n, iy).astype(W)
                keep = tri_mask(n, ul)
                if is_c:
                    ra = 0.7     # real alpha of HER / HPR
                    U1 = np.where(keep, Tt + ra * np.outer(xl, xl.conj()), 0)
                    U2 = np.where(keep, Tt + al * np.outer(xl, yl.conj()) + np.conj(al) * np.outer(yl, xl.conj()), 0)
                    for U in (U1, U2):
                        U[np.arange(n), np.arange(n)] = U[np.arange(n), np.arange(n)].real
                    add("her", [ul, n, ra, x, ix, AF, n + 1], 5, full_tri(U1.astype(dt), ul, n + 1, rowmajor), n, "%s n=%d inc=%d" % (ul, n, ix))
                    add("her2", [ul, n, scal(p, al), x, ix, y, iy, AF, n + 1], 7, full_tri(U2.astype(dt), ul, n + 1, rowmajor), n, "%s n=%d inc=%d,%d" % (ul, n, ix, iy))
                    add("hpr", [ul, n, ra, x, ix, AP], 5, packed(U1.astype(dt), ul, rowmajor), n, "%s n=%d inc=%d" % (ul, n, ix))
                    add("hpr2", [ul, n, scal(p, al), x, ix, y, iy, AP], 7, packed(U2.astype(dt), ul, rowmajor), n, "%s n=%d inc=%d,%d" % (ul, n, ix, iy))
                else:
                    U1 = np.where(keep, Tt + al * np.outer(xl, xl), 0)
                    U2 = np.where(keep, Tt + al * np.outer(xl, yl) + al * np.outer(yl, xl), 0)
                    add("syr2", [ul, n, al, x, ix, y, iy, AF, n + 1], 7, full_tri(U2.astype(dt), ul, n + 1, rowmajor), n, "%s n=%d inc=%d,%d" % (ul, n, ix, iy))
                    add("syr", [ul, n, al, x, ix, AF, n + 1], 5, full_tri(U1.astype(dt), ul, n + 1, rowmajor), n, "%s n=%d inc=%d" % (ul, n, ix))
                    add("spr", [ul, n, al, x, ix, AP], 5, packed(U1.astype(dt), ul, rowmajor), n, "%s n=%d inc=%d" % (ul, n, ix))
                    add("spr2", [ul, n, al, x, ix, y, iy, AP], 7, packed(U2.astype(dt), ul, rowmajor), n, "%s n=%d inc=%d,%d" % (ul, n, ix, iy))
            # ---- triangular products and solves
            G = well_conditioned_tri(nxt(), n, p).astype(W)
            for dg in "NU":
                for tr in transs:
                    for ix in ((1, -3) if n < 40 else (2,)):
                        x = x0[ix] if ix in x0 else vec(nxt(), n, ix, p)
                        xl = logical(x, n, ix).astype(W)
                        for k in ks + [None]:
                            Tk = np.where(tri_mask(n, ul, k), G, 0)
                            if dg == "U":
                                Tk[np.arange(n), np.arange(n)] = 1.0
                            prod = opmat(Tk, tr) @ xl
                            sol = np.linalg.solve(opmat(Tk, tr), xl) if n else xl
                            stored = np.where(tri_mask(n, ul, k), G, 0).astype(dt)
                            if dg == "U":   # a unit diagonal is never read: store rogue there
                                stored[np.arange(n), np.arange(n)] = ROGUE
                            tag = "%s%s%s n=%d k=%s inc=%d" % (ul, tr, dg, n, k, ix)
                            if k is not None:
                                A = band_tri(stored, ul, k, k + 2, rowmajor)
                                add("tbmv", [ul, tr, dg, n, k, A, k + 2, x, ix], 7, _put(x, n, ix, prod), n, tag)
                                add("tbsv", [ul, tr, dg, n, k, A, k + 2, x, ix], 7, _put(x, n, ix, sol), 4 * n, tag)
                            else:
                                AP2 = packed(stored, ul, rowmajor)
                                add("tpmv", [ul, tr, dg, n, AP2, x, ix], 5, _put(x, n, ix, prod), n, tag)
                                add("tpsv", [ul, tr, dg, n, AP2, x, ix], 5, _put(x, n, ix, sol), 4 * n, tag)
                                AF2 = full_tri(stored, ul, n + 1, rowmajor)
                                add("trmv", [ul, tr, dg, n, AF2, n + 1, x, ix], 6, _put(x, n, ix, prod), n, tag)
                                if is_c:   # (real TRSV is level2.cu's blocked solve, tested in test_level12_gpu.py)
                                    add("trsv", [ul, tr, dg, n, AF2, n + 1, x, ix], 6, _put(x, n, ix, sol), 4 * n, tag)
    # ---- GER / GERU / GERC
    if True:
        for (m, n) in [(1, 1), (7, 5), (33, 70)] + [(b, b + 5) for b in big]:
            for (ix, iy) in incs[: (3 if m < 40 else 1)]:
                x = vec(nxt(), m, ix, p); y = vec(nxt(), n, iy, p)
                M = rnd(nxt(), (m, n), p).astype(W)
                xl, yl = logical(x, m, ix).astype(W), logical(y, n, iy).astype(W)
                for nm, U in ((("geru", M + al * np.outer(xl, yl)), ("gerc", M + al * np.outer(xl, yl.conj()))) if is_c else (("ger", M + al * np.outer(xl, yl)),)):
                    if not rowmajor:
                        A = np.full((m + 1, n), ROGUE, dtype=dt, order="F"); A[:m] = M; E = A.copy(order="F"); E[:m] = U; lda = m + 1
                    else:
                        A = np.full((m, n + 1), ROGUE, dtype=dt, order="C"); A[:, :n] = M; E = A.copy(order="C"); E[:, :n] = U; lda = n + 1
                    add(nm, [m, n, scal(p, al), x, ix, y, iy, A, lda], 7, E, max(m, n), "m=%d n=%d inc=%d,%d" % (m, n, ix, iy))
    return out


def compare(got, case):
    """max abs error relative to the case's tolerance; rogue / padding positions must be bit-identical"""
    exp = case.expect
    assert got.shape == exp.shape, (case.tag, got.shape, exp.shape)
    rogue = exp == exp.dtype.type(ROGUE)
    if not np.array_equal(got[rogue], exp[rogue]):
        return float("inf")
    d = np.abs(got.astype(np.complex128) - exp.astype(np.complex128))
    scale = max(1.0, float(np.abs(exp[~rogue]).max()) if (~rogue).any() else 1.0)
    return float(d.max() / (case.tol * scale)) if d.size else 0.0
