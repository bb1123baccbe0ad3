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
        assert blat3.chke(p, cap.call) > 100


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
