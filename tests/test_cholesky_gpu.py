"""Blocked Cholesky workload (BASELINE.json configs[3]) on one GPU: parity with the oracle's restated unblocked
Cholesky and with OpenBLAS dpotrf_ at sizes they finish quickly, residual property at a larger size."""
import ctypes

import numpy as np
import pytest

import libgpublas_b200 as g
from libgpublas_b200.cholesky import blocked_cholesky, potrf_lower
from helpers import load_oracle, load_openblas, splitmix_uniform

pytestmark = pytest.mark.gpu
EPS = 2.0 ** -53


def spd(n, seed=9):
    """SURVEY 8d C4 input: symmetric U(-1,1) off-diagonal, diagonal n (diagonally dominant => SPD)."""
    A = splitmix_uniform(seed, (n, n))
    A = np.asfortranarray(np.tril(A, -1) + np.tril(A, -1).T + n * np.eye(n))
    return A


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 200, 513, 1024])
def test_potrf_lower_vs_oracle_and_openblas(n):
    A0 = spd(n)
    lda = n + 3
    A = np.zeros((lda, n), order="F"); A[:n] = A0; A[n:] = -1e10             # rogue padding must survive
    A[:n][np.triu_indices(n, 1)] = -1e10                                        # strictly upper: not referenced
    G = A.copy(order="F")
    assert potrf_lower(n, G, lda) == 0
    assert np.array_equal(G[n:], A[n:]) and np.array_equal(G[:n][np.triu_indices(n, 1)], A[:n][np.triu_indices(n, 1)])
    R = A.copy(order="F")
    lib = load_oracle()
    assert lib.ref_dpotrf_lower(ctypes.c_int(n), R.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(lda)) == 0
    L, Lr = np.tril(G[:n]), np.tril(R[:n])
    assert np.abs(L - Lr).max() <= 64 * n * EPS * np.abs(Lr).max()
    assert np.linalg.norm(L @ L.T - A0) <= 8 * n * EPS * np.linalg.norm(A0)
    ob = load_openblas()
    if ob is not None:
        O = A.copy(order="F"); info = ctypes.c_int(0)
        ob.dpotrf_(ctypes.c_char_p(b"L"), ctypes.byref(ctypes.c_int(n)), O.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ctypes.c_int(lda)),
                   ctypes.byref(info))
        assert info.value == 0 and np.abs(L - np.tril(O[:n])).max() <= 64 * n * EPS * np.abs(Lr).max()


def test_potrf_not_positive_definite_reports_the_minor():
    n = 300
    A = spd(n); A[137, 137] = -1.0
    assert potrf_lower(n, A, n) == 138
    assert g.load().b200blas_dpotrf_lower(-1, None, 1) == -1


@pytest.mark.parametrize("nb", [256, 1000])
def test_blocked_cholesky_workload_device_resident(nb):
    """DSYRK + DTRSM + diagonal blocks through the Fortran symbols on a device-resident matrix; checked with the
    random-vector probe ||A x - L (L^T x)|| and a leading block against the oracle."""
    import torch
    n = 3000
    A0 = spd(n, seed=19)
    At = torch.from_numpy(np.ascontiguousarray(A0.T)).cuda()      # row-major transposed == column-major A0 (symmetric anyway)
    torch.cuda.synchronize()
    assert blocked_cholesky(n, At, n, nb=nb) == 0
    torch.cuda.synchronize()
    L = np.tril(At.cpu().numpy().T)
    x = splitmix_uniform(5, (n,))
    assert np.linalg.norm(A0 @ x - L @ (L.T @ x)) <= 16 * n * EPS * np.linalg.norm(A0) * np.linalg.norm(x)
    k = 512
    R = np.asfortranarray(A0[:k, :k].copy())
    assert load_oracle().ref_dpotrf_lower(ctypes.c_int(k), R.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(k)) == 0
    assert np.abs(L[:k, :k] - np.tril(R)).max() <= 64 * k * EPS * np.abs(R).max()


def _chol_call(lib, n, ptr, lda, nb):
    lib.b200blas_cholesky_lower.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int]
    lib.b200blas_cholesky_lower.restype = ctypes.c_int
    return lib.b200blas_cholesky_lower(n, ctypes.c_void_p(ptr), lda, nb)


@pytest.mark.parametrize("n,nb", [(3000, 256), (2500, 1024), (1000, 2048), (130, 128)])
def test_cholesky_workload_single_call(n, nb):
    """b200blas_cholesky_lower: the whole blocked workload (diagonal-block factorisation, DTRSM panel, masked-GEMM trailing update,
    look-ahead 1) issued from C++ in one call -- here on one device.  Device-resident and host matrices, ragged last block,
    rogue padding and the strictly upper triangle untouched; residual ||A - L L^T|| and agreement with OpenBLAS dpotrf_."""
    import torch
    lib = g.load()
    A0 = spd(n, seed=23)
    lda = n + 2
    H = np.full((lda, n), -1e10, order="F"); H[:n] = A0
    H[:n][np.triu_indices(n, 1)] = -7e9
    # host matrix
    G = H.copy(order="F")
    assert _chol_call(lib, n, G.ctypes.data, lda, nb) == 0
    assert np.array_equal(G[n:], H[n:]) and np.array_equal(G[:n][np.triu_indices(n, 1)], H[:n][np.triu_indices(n, 1)])
    L = np.tril(G[:n])
    assert np.linalg.norm(L @ L.T - A0) <= 8 * n * EPS * np.linalg.norm(A0)
    # device-resident matrix: same bits
    D = torch.from_numpy(H.ravel(order="F").copy()).cuda()
    torch.cuda.synchronize()
    assert _chol_call(lib, n, D.data_ptr(), lda, nb) == 0
    torch.cuda.synchronize()
    assert np.array_equal(D.cpu().numpy().reshape((lda, n), order="F"), G)
    ob = load_openblas()
    if ob is not None:
        O = H.copy(order="F"); info = ctypes.c_int(0)
        ob.dpotrf_(ctypes.c_char_p(b"L"), ctypes.byref(ctypes.c_int(n)), O.ctypes.data_as(ctypes.c_void_p), ctypes.byref(ctypes.c_int(lda)), ctypes.byref(info))
        assert info.value == 0 and np.abs(L - np.tril(O[:n])).max() <= 64 * n * EPS * np.abs(L).max()
    # not positive definite: LAPACK info = order of the first failing leading minor
    Bad = H.copy(order="F"); Bad[n // 2 + 3, n // 2 + 3] = -1.0
    assert _chol_call(lib, n, Bad.ctypes.data, lda, nb) == n // 2 + 4
