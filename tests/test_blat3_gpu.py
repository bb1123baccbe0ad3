"""The restated netlib ?blat3 conformance programs (tests/blat3.py) run against libb200blas.so on the
GPU -- the reference's own correctness gate (tests/netlib/test.py), same shapes, same generator, same
ratio test < 16, same untouched-padding and error-exit checks."""
import ctypes

import pytest

import blat3
import libgpublas_b200 as g
from helpers import f77

pytestmark = pytest.mark.gpu


def call(name, *args):
    return f77(g.load(), name, *args)


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_blat3_gemm(p):
    nc, err = blat3.chk_gemm(p, call)
    assert nc == 17496 and err < blat3.THRESH


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_blat3_syrk(p):
    nc, err = blat3.chk_syrk(p, call)
    assert nc > 0 and err < blat3.THRESH


@pytest.mark.parametrize("which", ["trmm", "trsm"])
@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_blat3_trxm(p, which):
    nc, err = blat3.chk_trxm(p, call, which)
    assert nc == 2592 and err < blat3.THRESH


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_blat3_error_exits(p):
    lib = g.load()
    seen = []
    CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t)
    cb = CB(lambda name, info, ln: seen.append((name[:6].decode(), info[0])))
    lib.b200blas_set_xerbla(cb)

    def cap(name, *args):
        seen.clear()
        f77(lib, name, *args)
        return list(seen)
    try:
        assert blat3.chke(p, cap) > 100
    finally:
        lib.b200blas_set_xerbla(CB(0))


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_blat3_symm_syr2k(p):
    """SURVEY 8(f) rank 1 on the GPU: DCHK2 (SYMM) and DCHK5 (SYR2K)."""
    nc, err = blat3.chk_symm(p, call, "symm")
    assert nc == 1296 and err < blat3.THRESH
    nc, err = blat3.chk_r2k(p, call, "syr2k")
    assert nc > 0 and err < blat3.THRESH


@pytest.mark.parametrize("p", ["c", "z"])
def test_blat3_hermitian_family(p):
    """ZCHK2 (HEMM), ZCHK4 (HERK), ZCHK5 (HER2K)."""
    assert blat3.chk_symm(p, call, "hemm")[1] < blat3.THRESH
    assert blat3.chk_herk(p, call)[1] < blat3.THRESH
    assert blat3.chk_r2k(p, call, "her2k")[1] < blat3.THRESH
