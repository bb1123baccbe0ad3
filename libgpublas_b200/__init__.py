"""libgpublas_b200 -- B200-native drop-in for the BLAS hot path of Prince781/libgpublas.

The product is the C-ABI shared object ``libb200blas.so`` built from ``csrc/`` (hand-written
sm_100a kernels + the Fortran/CBLAS/allocator interposition layer).  This Python package is
plumbing only: it builds the library, loads it with ctypes and mirrors the reference's
interface (Fortran BLAS names and argument order) for tests, benchmarks and embedding.

There is deliberately no CPU or PyTorch fallback: if the shared object is missing, loading
raises; if no CUDA device is present, the first BLAS call aborts inside the library.
"""
import ctypes
import os

from ._ffi import DevPtr, f77call, routine_prec  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200blas.so")
_lib = None


def load():
    """ctypes handle of libb200blas.so (RTLD_LOCAL: its malloc/free exports interpose nothing
    when loaded this way; LD_PRELOAD is what activates the allocator tracker)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libb200blas.so is not built: run `python -m libgpublas_b200.build` "
                "(there is no CPU fallback by design)")
        _lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_LOCAL)
        _lib.b200blas_last_variant.restype = ctypes.c_char_p
        _lib.b200blas_malloc_managed.restype = ctypes.c_void_p
        _lib.b200blas_malloc_managed.argtypes = [ctypes.c_size_t]
        _lib.b200blas_free_managed.argtypes = [ctypes.c_void_p]
        _lib.b200blas_is_tracked.argtypes = [ctypes.c_void_p]
        _lib.b200blas_set_stream.argtypes = [ctypes.c_void_p]
        _lib.b200blas_force_variant.argtypes = [ctypes.c_char_p]
        _lib.b200blas_set_options.argtypes = [ctypes.c_char_p]
    return _lib


def call(name, *args, restype=None):
    """Call the Fortran-ABI entry point `name` (e.g. "dgemm_") with the reference's argument
    order; numpy arrays are host operands, torch CUDA tensors / DevPtr are used in place."""
    return f77call(load(), name, *args, restype=restype)


def last_variant():
    return load().b200blas_last_variant().decode()


def force_variant(name):
    load().b200blas_force_variant(None if name in (None, "auto") else name.encode())


def set_sync(on):
    load().b200blas_set_sync(1 if on else 0)


def use_torch_stream():
    """Run this thread's BLAS calls on torch's current CUDA stream (so torch.cuda.Event timing
    and tensor lifetimes line up)."""
    import torch
    load().b200blas_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in
                ("hits", "misses", "calls", "h2d_bytes", "d2h_bytes", "prefetch_bytes",
                 "managed_allocs", "managed_frees", "managed_bytes_live")]


def stats():
    s = Stats()
    load().b200blas_get_stats(ctypes.byref(s))
    return {n: int(getattr(s, n)) for n, _ in Stats._fields_}
