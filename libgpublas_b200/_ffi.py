"""Generic Fortran-ABI caller: every argument by reference, exactly what a Fortran caller of
the interposed symbols does (reference blas.h:202-314)."""
import ctypes

import numpy as np

_PREC = {"s": ctypes.c_float, "d": ctypes.c_double, "c": ctypes.c_float, "z": ctypes.c_double}


def routine_prec(name):
    """precision letter of a BLAS routine name: dgemm_->d, idamax_->d, dznrm2_->z, scnrm2_->c,
    cblas_dgemm->d."""
    n = name.lower()
    if n.startswith("cblas_"):
        n = n[6:]
    if n[0] == "i":
        return n[1]
    if n[:2] in ("dz", "sc"):
        return n[1]
    return n[0]


class DevPtr:
    """A raw (device or managed) address passed through unchanged."""

    def __init__(self, addr):
        self.addr = int(addr)


def as_ptr(a):
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    if isinstance(a, DevPtr):
        return ctypes.c_void_p(a.addr)
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    raise TypeError(type(a))


def f77call(lib, name, *args, restype=None):
    """str -> CHARACTER*1, int -> INTEGER, float/complex -> scalar of the routine's precision,
    np.float32/np.float64 -> that exact real type, arrays/tensors/DevPtr -> address."""
    creal = _PREC[routine_prec(name)]
    fn = getattr(lib, name)
    fn.restype = restype
    keep, cargs = [], []
    for a in args:
        if isinstance(a, str):
            c = ctypes.c_char(a.encode())
        elif isinstance(a, (bool, int, np.integer)):
            c = ctypes.c_int(int(a))
        elif isinstance(a, np.float32):
            c = ctypes.c_float(float(a))
        elif isinstance(a, np.float64):
            c = ctypes.c_double(float(a))
        elif isinstance(a, float):
            c = creal(a)
        elif isinstance(a, complex):
            c = (creal * 2)(a.real, a.imag)
        else:
            cargs.append(as_ptr(a))
            continue
        keep.append(c)
        cargs.append(ctypes.byref(c))
    return fn(*cargs)
