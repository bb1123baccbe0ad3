"""GPU parity for the blocked Level-3 routines (SYRK / TRMM / TRSM) at sizes that exercise the
recursion (several diagonal blocks, ragged edges, odd leading dimensions), against the oracle."""
import numpy as np
import pytest

import libgpublas_b200 as g
from helpers import f77, fro, oracle_call, splitmix_uniform

pytestmark = pytest.mark.gpu
DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": 2.0 ** -24, "d": 2.0 ** -53, "c": 2.0 ** -24, "z": 2.0 ** -53}


def F(x):
    return np.array(x, order="F")


@pytest.mark.parametrize("p", ["d", "z", "s"])
def test_syrk_vs_oracle(p):
    lib = g.load(); dt = DT[p]
    al, be = ((0.7 - 0.9j), (1.3 - 1.1j)) if p in "cz" else (0.7, 1.3)
    for (n, k) in [(70, 33), (129, 200), (300, 64), (257, 1), (390, 300)]:   # the last one reaches the tensor-core SGEMM tiles
        for uplo in "UL":
            for tr in "NT":
                ra, ca = (n, k) if tr == "N" else (k, n)
                A = splitmix_uniform(1, (ra + 1, ca), dt); C0 = splitmix_uniform(2, (n + 3, n), dt)
                C, R = F(C0), F(C0)
                f77(lib, p + "syrk_", uplo, tr, n, k, al, A, ra + 1, be, C, n + 3)
                assert oracle_call(p + "syrk", uplo, tr, n, k, al, A, ra + 1, be, R, n + 3) == 0
                tri = np.triu(np.ones((n, n), bool)) if uplo == "U" else np.tril(np.ones((n, n), bool))
                full = np.zeros((n + 3, n), bool); full[:n] = tri
                assert np.array_equal(C[~full], C0[~full]), "outside the referenced triangle must be untouched"
                err = fro((C - R)[full]); bound = 4 * (k + 2) * EPS[p] * (abs(al) * fro(A) ** 2 + abs(be) * fro(C0))
                assert err <= bound, (p, n, k, uplo, tr, err, bound)


@pytest.mark.parametrize("which", ["trmm", "trsm"])
@pytest.mark.parametrize("p", ["d", "z", "s"])
def test_trxm_vs_oracle(p, which):
    lib = g.load(); dt = DT[p]
    al = (0.7 - 0.9j) if p in "cz" else 0.7
    hi = np.complex128 if p in "cz" else np.float64
    for (m, n) in [(65, 40), (200, 130), (300, 70), (520, 33), (33, 520)]:
        for side in "LR":
            na = m if side == "L" else n
            T = splitmix_uniform(3, (na + 1, na), dt); T[np.arange(na), np.arange(na)] += dt(na)   # well conditioned
            for uplo in "UL":
                for ta in "NTC":
                    for diag in "NU":
                        if which == "trsm" and diag == "U":
                            T2 = F(T); T2[:na] = T2[:na] / dt(na)      # small off-diagonals keep the unit system tame
                        else:
                            T2 = T
                        B0 = splitmix_uniform(4, (m + 2, n), dt); B, R = F(B0), F(B0)
                        f77(lib, p + which + "_", side, uplo, ta, diag, m, n, al, T2, na + 1, B, m + 2)
                        assert oracle_call(p + which, side, uplo, ta, diag, m, n, al, T2, na + 1, R, m + 2) == 0
                        assert np.array_equal(B[m:], B0[m:])
                        if which == "trmm":
                            err = fro(B[:m] - R[:m]); bound = 4 * (na + 2) * EPS[p] * abs(al) * fro(np.triu(T2[:na]) if uplo == "U" else np.tril(T2[:na])) * fro(B0[:m])
                            assert err <= bound, (p, which, side, uplo, ta, diag, m, n, err, bound)
                        else:
                            # backward form: ||op(A) X - alpha B||_F <= c m eps ||A||_F ||X||_F, c = 4
                            Tm = (np.triu(T2[:na]) if uplo == "U" else np.tril(T2[:na])).astype(hi)
                            if diag == "U":
                                Tm[np.arange(na), np.arange(na)] = 1
                            opT = Tm if ta == "N" else (Tm.T if ta == "T" else Tm.conj().T)
                            X = B[:m].astype(hi)
                            resid = (opT @ X if side == "L" else X @ opT) - complex(dt(al)) * B0[:m].astype(hi) if p in "cz" else \
                                (opT @ X if side == "L" else X @ opT) - float(dt(al)) * B0[:m].astype(hi)
                            assert fro(resid) <= 4 * na * EPS[p] * fro(Tm) * fro(X) + 1e-300, (p, side, uplo, ta, diag, m, n, fro(resid))
                            assert np.allclose(B[:m], R[:m], rtol=1e4 * EPS[p], atol=1e4 * EPS[p])


def test_dsyrk_dtrsm_large_device_property():
    """Config-4-shaped panels on the device: SYRK equals the masked GEMM; TRSM round trip X*L^T -> B."""
    import torch
    lib = g.load()
    n, k = 2048, 512
    gen = torch.Generator(device="cuda").manual_seed(9)
    A = torch.rand((k, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1   # column-major n x k
    C = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    f77(lib, "dsyrk_", "L", "N", n, k, 1.0, A, n, 0.0, C, n)
    ref = (A.T @ A)                               # column-major A A^T == row-major (A^T A) transposed; symmetric
    low = torch.triu(torch.ones((n, n), dtype=torch.bool, device="cuda"))    # row-major upper == column-major lower
    assert ((C - ref)[low].abs().max().item()) <= 16 * 2.0 ** -53 * k
    assert C[~low].abs().max().item() == 0.0     # strictly upper (column-major) never written
    # TRSM: B := B * L^{-T}, then multiply back
    L = torch.tril(torch.rand((k, k), dtype=torch.float64, device="cuda", generator=gen)) + k * torch.eye(k, dtype=torch.float64, device="cuda")
    Lcm = L.T.contiguous()                        # column-major storage of L
    B0 = torch.rand((k, n), dtype=torch.float64, device="cuda", generator=gen)           # column-major n x k
    B = B0.clone()
    torch.cuda.synchronize()
    f77(lib, "dtrsm_", "R", "L", "T", "N", n, k, 1.0, Lcm, k, B, n)
    X = B.T                                        # n x k
    back = X @ L.T
    assert (back - B0.T).abs().max().item() <= 64 * 2.0 ** -53 * k * X.abs().max().item() * k
