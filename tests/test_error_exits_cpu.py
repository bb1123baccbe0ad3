"""Argument checking of the product library without a GPU.

Every entry point validates its arguments the way netlib does and reports through xerbla_ BEFORE its first CUDA call
(reference pattern: gemm.cc:87-127 `gemm_check`, runtime-blas.c:34-52 `runtime_blas_xerbla`), so the INFO numbering of the
whole exported surface can be pinned on a machine with no device:
  * Level 3: the restated netlib ?CHKE tables of tests/blat3.py (the same tables run on the GPU in test_blat3_gpu.py and
    against the oracle in test_oracle.py);
  * Level 2: one table generated from each routine's signature -- a valid call, then one illegal value per checked argument;
    the expected INFO is that argument's position (netlib checks in argument order) -- through the product AND the oracle.
"""
import ctypes

import numpy as np
import pytest

import blat3
import libgpublas_b200 as g
from helpers import f77, oracle_call

CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t)


class Capture:
    def __enter__(self):
        self.lib, self.seen = g.load(), []
        self.cb = CB(lambda name, info, ln: self.seen.append((name[:6].decode(), info[0])))
        self.lib.b200blas_set_xerbla(self.cb)
        return self

    def __exit__(self, *exc):
        self.lib.b200blas_set_xerbla(CB(0))

    def call(self, name, *args):
        self.seen.clear()
        f77(self.lib, name, *args)
        return list(self.seen)


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_level3_error_exits_without_a_device(p):
    with Capture() as cap:
        assert blat3.chke(p, cap.call) == (162 if p in "sd" else 291)      # netlib DCHKE: 162; ZCHKE: 288 (+3, see blat3.chke)


# Level-2 signatures: argument kinds in netlib order.  Checked kinds and their illegal value:
#   uplo/trans/diag -> 'X';  m/n/k/kl/ku -> -1;  lda -> 1 (every valid call below needs lda >= 2);  incx/incy -> 0
SIG = {
    "gemv": ["trans", "m", "n", "alpha", "A", "lda", "x", "incx", "beta", "y", "incy"],
    "gbmv": ["trans", "m", "n", "kl", "ku", "alpha", "A", "lda", "x", "incx", "beta", "y", "incy"],
    "symv": ["uplo", "n", "alpha", "A", "lda", "x", "incx", "beta", "y", "incy"],
    "hemv": ["uplo", "n", "alpha", "A", "lda", "x", "incx", "beta", "y", "incy"],
    "sbmv": ["uplo", "n", "k", "alpha", "A", "lda", "x", "incx", "beta", "y", "incy"],
    "hbmv": ["uplo", "n", "k", "alpha", "A", "lda", "x", "incx", "beta", "y", "incy"],
    "spmv": ["uplo", "n", "alpha", "A", "x", "incx", "beta", "y", "incy"],
    "hpmv": ["uplo", "n", "alpha", "A", "x", "incx", "beta", "y", "incy"],
    "trmv": ["uplo", "trans", "diag", "n", "A", "lda", "x", "incx"],
    "trsv": ["uplo", "trans", "diag", "n", "A", "lda", "x", "incx"],
    "tbmv": ["uplo", "trans", "diag", "n", "k", "A", "lda", "x", "incx"],
    "tbsv": ["uplo", "trans", "diag", "n", "k", "A", "lda", "x", "incx"],
    "tpmv": ["uplo", "trans", "diag", "n", "A", "x", "incx"],
    "tpsv": ["uplo", "trans", "diag", "n", "A", "x", "incx"],
    "ger": ["m", "n", "alpha", "x", "incx", "y", "incy", "A", "lda"],
    "geru": ["m", "n", "alpha", "x", "incx", "y", "incy", "A", "lda"],
    "gerc": ["m", "n", "alpha", "x", "incx", "y", "incy", "A", "lda"],
    "syr": ["uplo", "n", "alpha", "x", "incx", "A", "lda"],
    "her": ["uplo", "n", "ralpha", "x", "incx", "A", "lda"],
    "spr": ["uplo", "n", "alpha", "x", "incx", "A"],
    "hpr": ["uplo", "n", "ralpha", "x", "incx", "A"],
    "syr2": ["uplo", "n", "alpha", "x", "incx", "y", "incy", "A", "lda"],
    "her2": ["uplo", "n", "alpha", "x", "incx", "y", "incy", "A", "lda"],
    "spr2": ["uplo", "n", "alpha", "x", "incx", "y", "incy", "A"],
    "hpr2": ["uplo", "n", "alpha", "x", "incx", "y", "incy", "A"],
}
REAL_ONLY = {"symv", "sbmv", "spmv", "ger", "syr", "spr", "syr2", "spr2"}
CPLX_ONLY = {"hemv", "hbmv", "hpmv", "geru", "gerc", "her", "hpr", "her2", "hpr2"}
BAD = {"uplo": "X", "trans": "X", "diag": "X", "m": -1, "n": -1, "k": -1, "kl": -1, "ku": -1, "lda": 1, "incx": 0, "incy": 0}
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def _valid(kind, p, arrays):
    cplx = p in "cz"
    return {"uplo": "U", "trans": "N", "diag": "N", "m": 2, "n": 2, "k": 1, "kl": 1, "ku": 1, "lda": 3, "incx": 1, "incy": 1,
            "alpha": (1 + 0j) if cplx else 1.0, "beta": (1 + 0j) if cplx else 1.0, "ralpha": 1.0,
            "A": arrays[0], "x": arrays[1], "y": arrays[2]}[kind]


def _routines(p):
    for r in SIG:
        if (r in REAL_ONLY and p in "cz") or (r in CPLX_ONLY and p in "sd"):
            continue
        yield r


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_level2_info_numbering_product_and_oracle(p):
    checks = 0
    with Capture() as cap:
        for r in _routines(p):
            name = p + r
            arrays = [np.full((3, 3), 5.0, dtype=DT[p], order="F"), np.full(4, 7.0, dtype=DT[p]), np.full(4, 9.0, dtype=DT[p])]
            before = [a.copy() for a in arrays]
            valid = [_valid(k, p, arrays) for k in SIG[r]]
            for pos, kind in enumerate(SIG[r], start=1):
                if kind not in BAD:
                    continue
                args = list(valid); args[pos - 1] = BAD[kind]
                seen = cap.call(name + "_", *args)
                assert seen == [((name.upper() + "      ")[:6], pos)], (name, kind, seen)
                assert oracle_call(name, *args) == pos, (name, kind)
                checks += 1
            # an illegal call touches nothing (netlib: return right after XERBLA)
            assert all(np.array_equal(a, b) for a, b in zip(arrays, before)), name
    assert checks >= 80


QUICK_RETURN_SCRIPT = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import libgpublas_b200 as g
from helpers import f77
from test_error_exits_cpu import SIG, DT, _valid, _routines
lib = g.load()
n_calls = 0
for p in "sdcz":
    cplx = p in "cz"
    one, zero = ((1 + 0j), 0j) if cplx else (1.0, 0.0)
    A = np.full((3, 3), 5.0, dtype=DT[p], order="F"); B = A.copy(order="F"); C = A.copy(order="F")
    x = np.full(4, 7.0, dtype=DT[p]); y = np.full(4, 9.0, dtype=DT[p])
    # Level 2: a zero dimension; alpha == 0 with beta == 1 (matrix-vector) or alpha == 0 (rank updates)
    for r in _routines(p):
        valid = [_valid(k, p, [A, x, y]) for k in SIG[r]]
        for dim in ("m", "n"):
            if dim in SIG[r]:
                args = list(valid); args[SIG[r].index(dim)] = 0
                f77(lib, p + r + "_", *args); n_calls += 1
        for ak in ("alpha", "ralpha"):
            if ak in SIG[r] and (("beta" in SIG[r]) or r in ("ger", "geru", "gerc", "syr", "her", "spr", "hpr", "syr2", "her2", "spr2", "hpr2")):
                args = list(valid); args[SIG[r].index(ak)] = zero if ak == "alpha" else 0.0
                f77(lib, p + r + "_", *args); n_calls += 1
    # Level 3 (netlib quick returns: a zero dimension; (alpha == 0 or k == 0) with beta == 1)
    f77(lib, p + "gemm_", "N", "N", 0, 2, 2, one, A, 3, B, 3, one, C, 3); f77(lib, p + "gemm_", "N", "N", 2, 0, 2, one, A, 3, B, 3, one, C, 3)
    f77(lib, p + "gemm_", "N", "N", 2, 2, 0, one, A, 3, B, 3, one, C, 3); f77(lib, p + "gemm_", "N", "T", 2, 2, 2, zero, A, 3, B, 3, one, C, 3)
    f77(lib, p + "syrk_", "U", "N", 0, 2, one, A, 3, one, C, 3); f77(lib, p + "syrk_", "L", "T", 2, 0, one, A, 3, one, C, 3)
    f77(lib, p + "syr2k_", "U", "N", 0, 2, one, A, 3, B, 3, one, C, 3); f77(lib, p + "symm_", "L", "U", 0, 2, one, A, 3, B, 3, one, C, 3)
    f77(lib, p + "symm_", "R", "L", 2, 2, zero, A, 3, B, 3, one, C, 3)
    f77(lib, p + "trsm_", "L", "U", "N", "N", 0, 2, one, A, 3, B, 3); f77(lib, p + "trmm_", "R", "L", "T", "U", 2, 0, one, A, 3, B, 3)
    n_calls += 11
    if cplx:
        rone = 1.0
        f77(lib, p + "hemm_", "L", "U", 2, 0, one, A, 3, B, 3, one, C, 3); f77(lib, p + "herk_", "U", "N", 0, 2, rone, A, 3, rone, C, 3)
        f77(lib, p + "herk_", "U", "N", 2, 0, rone, A, 3, rone, C, 3); f77(lib, p + "her2k_", "L", "C", 0, 2, one, A, 3, B, 3, rone, C, 3)
        n_calls += 4
    # Level 1 with n <= 0
    import ctypes
    f77(lib, p + "axpy_", 0, one, x, 1, y, 1); f77(lib, p + "scal_", 0, one, x, 1); f77(lib, p + "copy_", 0, x, 1, y, 1); f77(lib, p + "swap_", -1, x, 1, y, 1)
    assert f77(lib, "i" + p + "amax_", 0, x, 1, restype=ctypes.c_int) == 0 and f77(lib, "i" + p + "amin_", 3, x, 0, restype=ctypes.c_int) == 0
    n_calls += 6
    assert np.all(A == 5) and np.all(B == 5) and np.all(C == 5) and np.all(x == 7) and np.all(y == 9), p
assert f77(lib, "ddot_", 0, np.ones(1), 1, np.ones(1), 1, restype=ctypes.c_double) == 0.0
assert f77(lib, "dnrm2_", 0, np.ones(1), 1, restype=ctypes.c_double) == 0.0 and f77(lib, "dasum_", 3, np.ones(3), 0, restype=ctypes.c_double) == 0.0
st = g.stats()
print("QUICK_RETURNS_OK", n_calls, st["calls"] if isinstance(st, dict) else st)
'''


def test_quick_returns_need_no_device():
    """netlib quick returns (a zero dimension, alpha == 0 with beta == 1, n <= 0 in Level 1) are decided on the host: the call
    returns without touching its operands and without bringing a device up.  Run in a child process: a quick return that did
    reach the device would abort there ("no CUDA device: libb200blas has no CPU fallback")."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", QUICK_RETURN_SCRIPT, root], capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0 and "QUICK_RETURNS_OK" in out.stdout, (out.stdout[-2000:], out.stderr[-2000:])
