"""Shared test plumbing: loaders for the oracle, the CPU BLAS the reference sits in front of
(OpenBLAS from this image) and the product library, plus one generic Fortran-ABI caller so a
parity test issues the *same* call on each of them.

Nothing here is imported by the product package.
"""
import ctypes
import glob
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ROOT, "libgpublas_b200", "libb200blas.so")


def build_oracle():
    so = os.path.join(ORACLE_DIR, "librefblas.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("refblas.c", "refblas_real.inc", "refblas_cplx.inc", "refblas_l2x.inc")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "librefblas.so"], stdout=subprocess.DEVNULL)
    return so


_oracle = None


def load_oracle():
    global _oracle
    if _oracle is None:
        _oracle = ctypes.CDLL(build_oracle())
    return _oracle


def find_openblas():
    """The CPU BLAS available offline in this image: OpenBLAS 0.3.15, LP64,
    plain Fortran symbols."""
    pats = []
    for p in sys.path:
        pats.append(os.path.join(p, "opencv_python_headless.libs", "libopenblasp-*.so"))
    for pat in pats:
        hits = sorted(glob.glob(pat))
        if hits:
            return hits[0]
    return None


_openblas = None


def load_openblas():
    global _openblas
    if _openblas is None:
        path = find_openblas()
        if path is None:
            return None
        # auto-detection on these Xeons falls back to SSE3 kernels
        os.environ.setdefault("OPENBLAS_CORETYPE", "SkylakeX")
        d = os.path.dirname(path)
        # its private libgfortran/libquadmath live beside it
        for dep in sorted(glob.glob(os.path.join(d, "libquadmath-*.so*"))) + sorted(
                glob.glob(os.path.join(d, "libgfortran-*.so*"))):
            ctypes.CDLL(dep, mode=ctypes.RTLD_GLOBAL)
        _openblas = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _openblas


from libgpublas_b200._ffi import DevPtr, as_ptr as _as_ptr, f77call, routine_prec  # noqa: E402,F401

_PREC = {"s": (ctypes.c_float, np.float32), "d": (ctypes.c_double, np.float64),
         "c": (ctypes.c_float, np.complex64), "z": (ctypes.c_double, np.complex128)}


def f77(lib, name, *args, restype=None):
    return f77call(lib, name, *args, restype=restype)


def oracle_call(name, *args, restype=ctypes.c_int):
    """Call oracle routine ref_<name> (by-value C signature; complex scalars by pointer)."""
    lib = load_oracle()
    prec = name[1] if name[0] == "i" else name[0]
    creal, _ = _PREC[prec]
    fn = getattr(lib, "ref_" + name)
    fn.restype = restype
    keep, cargs = [], []
    for a in args:
        if isinstance(a, str):
            cargs.append(ctypes.c_char(a.encode()))
        elif isinstance(a, (bool, int, np.integer)):
            cargs.append(ctypes.c_int(int(a)))
        elif isinstance(a, np.float32):
            cargs.append(ctypes.c_float(float(a)))
        elif isinstance(a, np.float64):
            cargs.append(ctypes.c_double(float(a)))
        elif isinstance(a, float):
            cargs.append(creal(a))
        elif isinstance(a, complex):
            c = (creal * 2)(a.real, a.imag)
            keep.append(c)
            cargs.append(ctypes.byref(c))
        else:
            cargs.append(_as_ptr(a))
    return fn(*cargs)


# ---------------------------------------------------------------------------------------------
# deterministic inputs: splitmix64 -> U(-1,1) (SURVEY.md section 8d "C1 synthetic input")
def splitmix_uniform(seed, shape, dtype=np.float64):
    n = int(np.prod(shape))
    cplx = np.issubdtype(dtype, np.complexfloating)
    cnt = 2 * n if cplx else n
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * np.arange(1, cnt + 1, dtype=np.uint64))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 2.0 - 1.0
    if cplx:
        u = u[0::2] + 1j * u[1::2]
    return np.asfortranarray(u.astype(dtype).reshape(shape, order="F"))


def fro(x):
    return float(np.linalg.norm(np.asarray(x, dtype=np.complex128 if np.iscomplexobj(x) else np.float64).ravel()))
