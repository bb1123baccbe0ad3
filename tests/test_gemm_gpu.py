"""GPU parity tests for the GEMM family through the C ABI (Fortran symbols), against the oracle.

Tolerances (SURVEY.md section 8c / BASELINE.json north_star):
  * Frobenius bound  ||C - C_ref||_F <= c * k * eps * ||A||_F * ||B||_F   with c = 2 (d, s), 4 (z, c)
  * netlib DMMCH element-wise ratio  max |C - C_ref| / (eps * (|alpha| |A||B| + |beta||C|)) < 16
"""
import ctypes

import numpy as np
import pytest

import libgpublas_b200 as g
from helpers import f77, fro, oracle_call, splitmix_uniform

pytestmark = pytest.mark.gpu

DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}
EPS = {"s": 2.0 ** -24, "d": 2.0 ** -53, "c": 2.0 ** -24, "z": 2.0 ** -53}
CFRO = {"s": 2, "d": 2, "c": 4, "z": 4}


def F(x):
    return np.array(x, order="F")


def op(x, t):
    return x if t == "N" else (x.T if t == "T" else x.conj().T)


def check_gemm(p, ta, tb, m, n, k, alpha, beta, A, B, C0, C):
    """C (result of the library) against the float64/complex128 numpy product and the oracle."""
    hi = np.complex128 if p in "cz" else np.float64
    a = op(A.astype(hi)[: (m if ta == "N" else k), : (k if ta == "N" else m)], ta)
    b = op(B.astype(hi)[: (k if tb == "N" else n), : (n if tb == "N" else k)], tb)
    ref = alpha * (a @ b) + (beta * C0.astype(hi)[:m, :n] if beta != 0 else 0)
    err = np.abs(C.astype(hi)[:m, :n] - ref)
    gbound = abs(alpha) * (np.abs(a) @ np.abs(b)) + abs(beta) * np.abs(C0.astype(hi)[:m, :n])
    ratio = float((err / (EPS[p] * np.maximum(gbound, np.finfo(np.float64).tiny))).max()) if err.size else 0.0
    frob = fro(err)
    bound = CFRO[p] * (k + 2) * EPS[p] * (abs(alpha) * fro(a) * fro(b) + abs(beta) * fro(C0[:m, :n]))
    assert frob <= bound, (frob, bound)
    assert ratio < 16.0, ratio
    # rows m..ldc-1 (padding) and columns >= n must be untouched (netlib LDERES)
    assert np.array_equal(C[m:, :], C0[m:, :])
    assert np.array_equal(C[:, n:], C0[:, n:])


SHAPES = [(1, 1, 1), (2, 3, 5), (9, 9, 9), (37, 29, 41), (64, 64, 64), (129, 127, 65), (200, 300, 17), (256, 128, 512)]


@pytest.mark.parametrize("variant", ["auto", "generic_tile", "dmma_tma", "dmma_ldg"])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T"), ("C", "N")])
def test_dgemm_host_operands(variant, ta, tb):
    lib = g.load()
    g.force_variant(variant)
    try:
        for (m, n, k) in SHAPES:
            for (alpha, beta) in [(1.0, 0.0), (0.7, 1.3)]:
                ra, ca = (m, k) if ta == "N" else (k, m)
                rb, cb = (k, n) if tb == "N" else (n, k)
                lda, ldb, ldc = ra + 1, rb + 3, m + 2
                A = splitmix_uniform(11, (lda, ca)); B = splitmix_uniform(12, (ldb, cb)); C0 = splitmix_uniform(13, (ldc, n + 1))
                C = F(C0)
                f77(lib, "dgemm_", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
                check_gemm("d", ta, tb, m, n, k, alpha, beta, A, B, C0, C)
                # and bit-level agreement in structure with the oracle's own restatement
                C2 = F(C0)
                assert oracle_call("dgemm", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C2, ldc) == 0
                assert np.allclose(C[:m, :n], C2[:m, :n], rtol=1e-11, atol=1e-11)
    finally:
        g.force_variant("auto")


@pytest.mark.parametrize("p", ["s", "c", "z"])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "C"), ("C", "T")])
def test_other_gemm_host_operands(p, ta, tb):
    lib = g.load()
    dt = DT[p]
    for (m, n, k) in [(1, 1, 1), (5, 9, 3), (37, 29, 41), (130, 70, 90)]:
        alpha, beta = ((0.7 - 0.9j), (1.3 - 1.1j)) if p in "cz" else (0.7, 1.3)
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        lda, ldb, ldc = ra + 1, rb + 1, m + 1
        A = splitmix_uniform(21, (lda, ca), dt); B = splitmix_uniform(22, (ldb, cb), dt); C0 = splitmix_uniform(23, (ldc, n), dt)
        C = F(C0)
        f77(lib, p + "gemm_", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
        check_gemm(p, ta, tb, m, n, k, alpha, beta, A, B, C0, C)


def test_dgemm_quick_returns_and_scaling():
    lib = g.load()
    m, n, k = 33, 17, 9
    A = splitmix_uniform(1, (m, k)); B = splitmix_uniform(2, (k, n)); C0 = splitmix_uniform(3, (m + 1, n))
    # alpha == 0: C := beta*C without reading A/B; beta == 0 must overwrite NaN
    C = F(C0); f77(lib, "dgemm_", "N", "N", m, n, k, 0.0, A, m, B, k, 1.3, C, m + 1)
    assert np.allclose(C[:m], 1.3 * C0[:m]) and np.array_equal(C[m:], C0[m:])
    C = F(C0); C[:m] = np.nan; f77(lib, "dgemm_", "N", "N", m, n, k, 0.0, A, m, B, k, 0.0, C, m + 1)
    assert np.all(C[:m] == 0.0)
    C = F(C0); C[:m] = np.nan; f77(lib, "dgemm_", "N", "N", m, n, k, 1.0, A, m, B, k, 0.0, C, m + 1)
    assert np.allclose(C[:m], A @ B)
    # k == 0 and beta == 1, m == 0, n == 0: untouched
    for (mm, nn, kk, be) in [(m, n, 0, 1.0), (0, n, k, 0.5), (m, 0, k, 0.5)]:
        C = F(C0); f77(lib, "dgemm_", "N", "N", mm, nn, kk, 1.0, A, m, B, k, be, C, m + 1)
        assert np.array_equal(C, C0)


def test_gemm_error_exits():
    """netlib DCHKE-style: each illegal argument reports its INFO through XERBLA and leaves C alone
    (reference gemm.cc:96-111 numbering; tests/netlib/dblat3.f DCHKE)."""
    lib = g.load()
    seen = []
    CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t)
    cb = CB(lambda name, info, ln: seen.append((name[:6].decode(), info[0])))
    lib.b200blas_set_xerbla(cb)
    try:
        A = np.zeros((4, 4), order="F"); C = np.ones((4, 4), order="F")
        cases = [(("X", "N", 2, 2, 2, 2, 2, 2), 1), (("N", "X", 2, 2, 2, 2, 2, 2), 2), (("N", "N", -1, 2, 2, 2, 2, 2), 3),
                 (("N", "N", 2, -1, 2, 2, 2, 2), 4), (("N", "N", 2, 2, -1, 2, 2, 2), 5), (("N", "N", 2, 2, 2, 1, 2, 2), 8),
                 (("T", "N", 2, 2, 3, 2, 3, 2), 8), (("N", "N", 2, 2, 2, 2, 1, 2), 10), (("N", "T", 2, 3, 2, 2, 2, 2), 10),
                 (("N", "N", 2, 2, 2, 2, 2, 1), 13)]
        for (ta, tb, m, n, k, lda, ldb, ldc), want in cases:
            for p in "sdcz":
                seen.clear()
                al = 1.0 if p in "sd" else 1.0 + 0j
                f77(lib, p + "gemm_", ta, tb, m, n, k, al, A, lda, A, ldb, al, C, ldc)
                assert seen == [((p + "gemm").upper().ljust(6), want)], (p, seen, want)
                assert np.all(C == 1.0)
    finally:
        lib.b200blas_set_xerbla(CB(0))


def test_dgemm_golden_fixture_1024():
    """BASELINE config 1 / reference tests/c/gemm.c:29-35 restated in f64: A[i,j]=i, B[i,j]=j,
    alpha=1, beta=0 => C[i,j] = k*i*j exactly (all partial sums < 2^53)."""
    lib = g.load()
    n = 1024
    i = np.arange(n, dtype=np.float64)
    A = F(np.repeat(i[:, None], n, axis=1)); B = F(np.repeat(i[None, :], n, axis=0)); C = np.zeros((n, n), order="F")
    f77(lib, "dgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n)
    assert g.last_variant() == "dmma_tma"                     # 64 tiles of 128x128 < 148 SMs: the small-tile DMMA variant
    assert np.array_equal(C, n * np.outer(i, i))
    g.force_variant("dmma_tma")                               # and the 128x128 TMA kernel on the same fixture
    try:
        C[:] = -1
        f77(lib, "dgemm_", "N", "N", n, n, n, 1.0, A, n, B, n, 0.0, C, n)
        assert g.last_variant() == "dmma_tma" and np.array_equal(C, n * np.outer(i, i))
    finally:
        g.force_variant("auto")
    # same through CBLAS, column- and row-major
    for order in (102, 101):
        C[:] = -1
        lib.cblas_dgemm(ctypes.c_int(order), ctypes.c_int(111), ctypes.c_int(111), n, n, n, ctypes.c_double(1.0),
                        A.ctypes.data_as(ctypes.c_void_p), n, B.ctypes.data_as(ctypes.c_void_p), n, ctypes.c_double(0.0),
                        C.ctypes.data_as(ctypes.c_void_p), n)
        want = n * np.outer(i, i) if order == 102 else (A.T @ B.T).T   # row-major view of the same buffers
        assert np.array_equal(C, want)


def test_dgemm_device_resident_vs_oracle_and_unaligned():
    """Operands already on the device (torch tensors) are used in place; odd lda / 8-byte-offset
    base pointers route to the LDG-staged DMMA variant and must agree."""
    import torch
    lib = g.load()
    m, n, k = 300, 260, 190
    for (ta, tb) in [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")]:
        for (lda_pad, off) in [(0, 0), (1, 0), (0, 1), (3, 1), (0, 0)]:
            forced = (lda_pad, off) == (0, 0) and ta == "T"      # aligned case: also run the 128x128 TMA kernel on these ragged shapes
            g.force_variant("dmma_tma" if forced else "auto")
            ra, ca = (m, k) if ta == "N" else (k, m)
            rb, cb = (k, n) if tb == "N" else (n, k)
            lda, ldb, ldc = ra + lda_pad, rb + lda_pad, m + lda_pad
            A = splitmix_uniform(31, (lda, ca)); B = splitmix_uniform(32, (ldb, cb)); C0 = splitmix_uniform(33, (ldc, n))
            dA = torch.zeros(lda * ca + 2, dtype=torch.float64, device="cuda"); dA[off:off + lda * ca] = torch.from_numpy(A.ravel(order="F")).cuda()
            dB = torch.zeros(ldb * cb + 2, dtype=torch.float64, device="cuda"); dB[off:off + ldb * cb] = torch.from_numpy(B.ravel(order="F")).cuda()
            dC = torch.zeros(ldc * n + 2, dtype=torch.float64, device="cuda"); dC[off:off + ldc * n] = torch.from_numpy(C0.ravel(order="F")).cuda()
            torch.cuda.synchronize()
            f77(lib, "dgemm_", ta, tb, m, n, k, 0.7, g.DevPtr(dA.data_ptr() + 8 * off), lda, g.DevPtr(dB.data_ptr() + 8 * off), ldb,
                1.3, g.DevPtr(dC.data_ptr() + 8 * off), ldc)
            g.force_variant("auto")
            want = "dmma_tma" if (lda_pad % 2 == 0 and off == 0) else "dmma_ldg"   # TMA needs 16-byte aligned base and pitch
            assert g.last_variant() == want, (g.last_variant(), want)
            C = dC[off:off + ldc * n].cpu().numpy().reshape((ldc, n), order="F")
            check_gemm("d", ta, tb, m, n, k, 0.7, 1.3, A, B, C0, C)


def test_dgemm_large_property_linearity():
    """Size-independent property at a size the oracle cannot reach: (A1+A2)B == A1 B + A2 B to
    rounding, and agreement with an independent fp64 product (torch/cuBLAS) within the bound."""
    import torch
    lib = g.load()
    n = 2048
    gen = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    B = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    C = torch.empty((n, n), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    # torch tensors are row-major: as column-major buffers they are A^T, B^T; C^T = B^T A^T
    f77(lib, "dgemm_", "N", "N", n, n, n, 1.0, B, n, A, n, 0.0, C, n)
    ref = A @ B
    err = torch.linalg.norm(C - ref).item()
    bound = 2 * n * 2.0 ** -53 * torch.linalg.norm(A).item() * torch.linalg.norm(B).item()
    assert err <= bound, (err, bound)


@pytest.mark.parametrize("ta", ["N", "T", "C"])
@pytest.mark.parametrize("tb", ["N", "T", "C"])
def test_zgemm_dmma_all_ops(ta, tb):
    """ZGEMM on the DMMA pipeline: every transpose/conjugate pair, ragged tiles and k tails,
    c = 4 Frobenius bound and the netlib ratio test; must agree with the generic tile kernel too."""
    lib = g.load()
    for (m, n, k) in [(64, 128, 8), (65, 129, 9), (130, 200, 77), (300, 70, 160)]:
        alpha, beta = (0.7 - 0.9j), (1.3 - 1.1j)                  # input.zblat3:11-14
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        lda, ldb, ldc = ra + 1, rb + 2, m + 3
        A = splitmix_uniform(41, (lda, ca), np.complex128); B = splitmix_uniform(42, (ldb, cb), np.complex128)
        C0 = splitmix_uniform(43, (ldc, n + 1), np.complex128)
        C = F(C0)
        f77(lib, "zgemm_", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
        assert g.last_variant() == "dmma_tma"
        check_gemm("z", ta, tb, m, n, k, alpha, beta, A, B, C0, C)
        g.force_variant("generic_tile")
        try:
            C2 = F(C0)
            f77(lib, "zgemm_", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C2, ldc)
            assert g.last_variant() == "generic_tile"
        finally:
            g.force_variant("auto")
        assert np.allclose(C[:m, :n], C2[:m, :n], rtol=1e-12, atol=1e-12)


def test_zgemm_large_device_resident():
    """BASELINE config 5 shape family at a size the host check finishes quickly: device-resident
    operands, beta = 0 must not read C (NaN-filled), spot-checked rows/columns against numpy."""
    import torch
    lib = g.load()
    n = 1536
    gen = torch.Generator(device="cuda").manual_seed(10)
    A = torch.rand((n, n), dtype=torch.complex128, device="cuda", generator=gen) - (0.5 + 0.5j)
    B = torch.rand((n, n), dtype=torch.complex128, device="cuda", generator=gen) - (0.5 + 0.5j)
    C = torch.full((n, n), float("nan"), dtype=torch.complex128, device="cuda")
    # torch tensors are row-major: the column-major view of A is A^T, so C_cm = A_cm * B_cm  <=>  C^T = A^T B^T
    f77(lib, "zgemm_", "N", "C", n, n, n, 1.0 + 0.0j, A, n, B, n, 0.0 + 0.0j, C, n)
    torch.cuda.synchronize()
    assert g.last_variant() == "dmma_tma"
    Acm, Bcm, Ccm = A.T, B.T, C.T
    ref = Acm @ Bcm.conj().T
    err = (Ccm - ref).abs().max().item()
    bound = 4 * n * 2.0 ** -53 * float(torch.linalg.norm(Acm[0]).item()) * float(torch.linalg.norm(Bcm[0]).item()) * 4
    assert err <= bound, (err, bound)


@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T"), ("C", "N")])
def test_sgemm_tf32x3_tcgen05(ta, tb):
    """SGEMM on tcgen05 (3xTF32 split, TMEM accumulators) must hold FP32 accuracy: the c=4 Frobenius
    bound with eps = 2^-24 AND the netlib DMMCH element-wise ratio test (< 16 eps-units of sum|a||b|)."""
    lib = g.load()
    g.force_variant("tf32x3_tcgen05")
    try:
        for (m, n, k) in [(128, 128, 32), (128, 128, 64), (129, 131, 40), (300, 260, 515), (64, 1000, 77)]:
            for (alpha, beta) in [(1.0, 0.0), (0.7, 1.3)]:
                ra, ca = (m, k) if ta == "N" else (k, m)
                rb, cb = (k, n) if tb == "N" else (n, k)
                lda, ldb, ldc = ra + 1, rb + 3, m + 2
                A = splitmix_uniform(51, (lda, ca), np.float32); B = splitmix_uniform(52, (ldb, cb), np.float32)
                C0 = splitmix_uniform(53, (ldc, n + 1), np.float32)
                C = F(C0)
                f77(lib, "sgemm_", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
                assert g.last_variant() == "tf32x3_tcgen05"
                check_gemm("s", ta, tb, m, n, k, alpha, beta, A, B, C0, C)
    finally:
        g.force_variant("auto")


@pytest.mark.parametrize("cfg", [0, 1, 2])
def test_sgemm_every_tile_configuration(cfg):
    """Each tcgen05 tile configuration forced in turn (option sgemm_cfg: 0 = 128x128 BK32 x3, 1 = 128x256 BK32 x2,
    2 = 128x256 BK16 x4 with the 64-byte swizzle -- the one behind the benchmarked 16384^3 number, which round 1 never compared
    with anything) on ragged shapes: partial tiles in m and n, k tails that are not a multiple of BK nor of the 128-deep TMEM
    chunk, odd leading dimensions, all transposes.  Same bar as every SGEMM: c = 4 Frobenius bound with eps = 2^-24 and the
    netlib ratio < 16, against the float64 numpy product."""
    lib = g.load()
    lib.b200blas_set_options(("sgemm_cfg=%d" % cfg).encode())
    g.force_variant("tf32x3_tcgen05")
    try:
        for (ta, tb) in [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")]:
            for (m, n, k) in [(128, 256, 16), (129, 257, 17), (300, 520, 515), (257, 300, 1000), (64, 1000, 130)]:
                for (alpha, beta) in [(1.0, 0.0), (0.7, 1.3)]:
                    ra, ca = (m, k) if ta == "N" else (k, m)
                    rb, cb = (k, n) if tb == "N" else (n, k)
                    lda, ldb, ldc = ra + 1, rb + 3, m + 2
                    A = splitmix_uniform(51, (lda, ca), np.float32); B = splitmix_uniform(52, (ldb, cb), np.float32)
                    C0 = splitmix_uniform(53, (ldc, n + 1), np.float32)
                    C = F(C0)
                    f77(lib, "sgemm_", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
                    assert g.last_variant() == "tf32x3_tcgen05"
                    check_gemm("s", ta, tb, m, n, k, alpha, beta, A, B, C0, C)
    finally:
        g.force_variant("auto")
        lib.b200blas_set_options(b"sgemm_cfg=-1")


def test_sgemm_wide_tile_selected_by_size():
    """A shape whose 128x256 tiling fills the SMs (16 x 19 = 304 tiles >= 148) so the size-based selector itself picks the
    wide configuration <256,16,4>; host operands, ragged n and k.  Checked against float64 numpy on all of C."""
    lib = g.load()
    m, n, k = 2048, 4864 - 7, 1000 + 3
    for (ta, tb, alpha, beta) in [("N", "N", 1.0, 0.0), ("T", "T", 0.7, 1.3)]:
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        lda, ldb, ldc = ra + 4, rb + 4, m + 4
        A = splitmix_uniform(54, (lda, ca), np.float32); B = splitmix_uniform(55, (ldb, cb), np.float32)
        C0 = splitmix_uniform(56, (ldc, n), np.float32)
        C = F(C0)
        f77(lib, "sgemm_", ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
        assert g.last_variant() == "tf32x3_tcgen05"
        check_gemm("s", ta, tb, m, n, k, alpha, beta, A, B, C0, C)


def test_sgemm_inf_nan_and_flt_max_inputs():
    """IEEE special values through the tensor-core SGEMM: the 3xTF32 split of +-Inf would meet the other operand's
    lo = 0 (Inf * 0 = NaN where sgemm_ returns Inf), and rounding FLT_MAX up to the tf32 grid would turn a finite input into Inf.
    Finite-but-huge inputs stay finite (hi is truncated when rounding would overflow); an operand holding Inf/NaN raises a device
    flag in the split pass, the tensor-core kernel stands down and the FFMA tile kernel computes the product -- so the result has
    the CPU BLAS's Inf/NaN pattern.  Compared with a float32 numpy product on the same inputs."""
    lib = g.load()
    g.force_variant("tf32x3_tcgen05")
    try:
        m, n, k = 260, 300, 200
        A0 = splitmix_uniform(57, (m, k), np.float32); B0 = np.abs(splitmix_uniform(58, (k, n), np.float32)) + np.float32(0.1)
        fmax = np.finfo(np.float32).max
        # 1. FLT_MAX times small numbers: finite everywhere, close to the float64 product
        A = A0.copy(order="F"); B = (B0 * np.float32(1e-3)).copy(order="F")
        A[7, 2] = fmax; A[100, 150] = -fmax
        B[2, :] = np.float32(2.0 ** -20); B[150, :] = np.float32(2.0 ** -21)
        C = np.zeros((m, n), np.float32, order="F")
        f77(lib, "sgemm_", "N", "N", m, n, k, 1.0, A, m, B, k, 0.0, C, m)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        assert np.all(np.isfinite(C))
        gb = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64))
        assert float((np.abs(C - ref) / (2.0 ** -24 * gb)).max()) < 16.0
        # 2. +-Inf and NaN in A, Inf in B: same non-finite pattern as float32 arithmetic, finite entries within the usual bound
        for (ta, tb) in [("N", "N"), ("T", "N"), ("N", "T")]:
            A = A0.copy(order="F"); B = B0.copy(order="F")
            A[3, 5] = np.inf; A[40, 9] = -np.inf; A[77, 11] = np.nan
            B[150, 20] = np.inf
            Ax = F(A.T) if ta == "T" else A
            Bx = F(B.T) if tb == "T" else B
            C = np.zeros((m, n), np.float32, order="F")
            f77(lib, "sgemm_", ta, tb, m, n, k, 1.0, Ax, Ax.shape[0], Bx, Bx.shape[0], 0.0, C, m)
            with np.errstate(invalid="ignore", over="ignore"):
                ref32 = A @ B
            assert np.array_equal(np.isnan(C), np.isnan(ref32)), (ta, tb)
            assert np.array_equal(np.isposinf(C), np.isposinf(ref32)) and np.array_equal(np.isneginf(C), np.isneginf(ref32)), (ta, tb)
            fin = np.isfinite(ref32)
            assert np.allclose(C[fin], ref32[fin], rtol=1e-4, atol=1e-4)
    finally:
        g.force_variant("auto")


def _rows_check(name, C_rows, ref_rows, k, eps, c, anorm_rows, bnorm):
    err = float(np.linalg.norm((C_rows - ref_rows).ravel()))
    bound = c * k * eps * anorm_rows * bnorm
    assert err <= bound, (name, err, bound)
    return err / bound


@pytest.mark.parametrize("which", ["dgemm_16384", "sgemm_16384", "zgemm_8192_NN", "zgemm_8192_NC"])
def test_gemm_at_the_benchmarked_sizes(which):
    """Parity AT the sizes bench.py quotes (BASELINE.json configs[1] and [4]; VERDICT r1 item 2b: the largest DGEMM checked against
    anything independent was n = 2048, SGEMM 515, ZGEMM 1536).  Device-resident operands with splitmix-free torch generators;
    128 random rows of C x ALL columns are compared with a float64 numpy panel product computed on the host from the same
    operands (A rows gathered on the device, B copied back whole), with the routine's Frobenius bound restricted to those rows:
    ||C_R - (A_R B)||_F <= c k eps ||A_R||_F ||B||_F, c = 2 (d), 4 (s: eps = 2^-24; z)."""
    import torch
    lib = g.load()
    p = which[0]
    n = 16384 if "16384" in which else 8192
    tb = "C" if which.endswith("NC") else "N"
    tdt = {"d": torch.float64, "s": torch.float32, "z": torch.complex128}[p]
    gen = torch.Generator(device="cuda").manual_seed(1234)
    def rnd():
        if p == "z":
            return torch.complex(torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1,
                                 torch.rand((n, n), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1)
        return torch.rand((n, n), dtype=tdt, device="cuda", generator=gen) * 2 - 1
    # torch tensors are row-major: At holds the column-major A as its transpose, i.e. A_cm[i, j] = At[j, i]
    At, Bt = rnd(), rnd()
    Ct = torch.full((n, n), float("nan"), dtype=tdt, device="cuda")
    alpha = (0.7 - 0.9j) if p == "z" else 1.0
    beta = 0.0 * alpha
    torch.cuda.synchronize()
    f77(lib, p + "gemm_", "N", tb, n, n, n, alpha, At, n, Bt, n, beta, Ct, n)
    torch.cuda.synchronize()
    want_variant = {"d": "dmma_tma", "s": "tf32x3_tcgen05", "z": "dmma_tma"}[p]
    assert g.last_variant() == want_variant, g.last_variant()
    rows = torch.from_numpy(np.random.default_rng(7).choice(n, size=128, replace=False)).cuda()
    hi = np.complex128 if p == "z" else np.float64
    A_rows = At[:, rows].T.contiguous().cpu().numpy().astype(hi)           # A_cm[rows, :]   (128 x k)
    B_cm = Bt.cpu().numpy().astype(hi).T                                    # k x n (N)  or  n x k stored, used as B^H (C)
    opB = B_cm.conj().T if tb == "C" else B_cm
    ref = alpha * (A_rows @ opB)
    C_rows = Ct[:, rows].T.contiguous().cpu().numpy().astype(hi)
    assert np.all(np.isfinite(C_rows))
    eps, c = {"d": (2.0 ** -53, 2), "s": (2.0 ** -24, 4), "z": (2.0 ** -53, 4)}[p]
    _rows_check(which, C_rows, ref, n, eps, c, abs(alpha) * float(np.linalg.norm(A_rows)), float(np.linalg.norm(opB)))
    # and the netlib element-wise ratio on a 128 x 512 corner of those rows (|A||B| panel product kept small)
    gb = abs(alpha) * (np.abs(A_rows) @ np.abs(opB[:, :512]))
    ratio = float((np.abs(C_rows[:, :512] - ref[:, :512]) / (eps * gb)).max())
    assert ratio < 16.0, ratio


@pytest.mark.parametrize("p", ["d", "s", "z"])
@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "T"), ("C", "T")])
def test_gemm_pipelined_staging_host_operands(p, ta, tb):
    """Host-resident operands above pipeline_min bytes are staged in k-chunks / column panels overlapped
    with the multiply (staged_gemm.cuh): ragged last chunk and panel, odd leading dimensions,
    beta != 0 (C goes in and out), same numerical bar as the plain path."""
    lib = g.load()
    dt = DT[p]
    lib.b200blas_set_options(b"pipeline_min=1000")
    try:
        m, n, k = 384, 2304, 2200
        alpha, beta = ((0.7 - 0.9j), (1.3 - 1.1j)) if p == "z" else (0.7, 1.3)
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        lda, ldb, ldc = ra + 1, rb + 3, m + 2
        A = splitmix_uniform(61, (lda, ca), dt); B = splitmix_uniform(62, (ldb, cb), dt); C0 = splitmix_uniform(63, (ldc, n + 1), dt)
        for (al, be) in [(alpha, beta), (alpha, 0.0 * alpha)]:
            C = F(C0)
            s0 = g.stats()
            f77(lib, p + "gemm_", ta, tb, m, n, k, al, A, lda, B, ldb, be, C, ldc)
            s1 = g.stats()
            check_gemm(p, ta, tb, m, n, k, al, be, A, B, C0, C)
            es = np.dtype(dt).itemsize
            assert s1["h2d_bytes"] - s0["h2d_bytes"] == (m * k + k * n + (m * n if be != 0 else 0)) * es
            assert s1["d2h_bytes"] - s0["d2h_bytes"] == m * n * es
    finally:
        lib.b200blas_set_options(b"pipeline_min=67108864")


def test_dgemm_pipelined_mixed_residency():
    """A on the device (used in place), B and C in host memory (staged): same schedule, no copies for A."""
    import torch
    lib = g.load()
    lib.b200blas_set_options(b"pipeline_min=1000")
    try:
        m, n, k = 256, 1536, 1280
        gen = torch.Generator(device="cuda").manual_seed(5)
        At = torch.rand((k, m), dtype=torch.float64, device="cuda", generator=gen) * 2 - 1     # row-major (k,m) == column-major m x k
        B = splitmix_uniform(71, (k, n)); C0 = splitmix_uniform(72, (m, n)); C = F(C0)
        torch.cuda.synchronize()
        s0 = g.stats()
        f77(lib, "dgemm_", "N", "N", m, n, k, 1.0, At, m, B, k, 0.5, C, m)
        s1 = g.stats()
        Ah = At.cpu().numpy().T                                                                   # m x k
        ref = Ah @ B + 0.5 * C0
        assert np.abs(C - ref).max() <= 16 * 2.0 ** -53 * k
        assert s1["h2d_bytes"] - s0["h2d_bytes"] == (k * n + m * n) * 8 and s1["hits"] - s0["hits"] == 1
    finally:
        lib.b200blas_set_options(b"pipeline_min=67108864")


def test_entry_points_are_thread_safe():
    """SURVEY 8(b) "Threading": entry points may be called from any thread.  Each thread gets its own stream and workspace
    (csrc/runtime.cu ThreadCtx); eight threads issue different staged GEMMs / reductions concurrently and every result
    must equal the single-threaded one."""
    import threading
    lib = g.load()
    rng = np.random.default_rng(11)
    jobs = []
    for t in range(8):
        m, n, k = 200 + 17 * t, 150 + 9 * t, 300 + 31 * t
        A = F(rng.uniform(-1, 1, (m, k))); B = F(rng.uniform(-1, 1, (k, n))); x = rng.uniform(-1, 1, 50000 + 1000 * t)
        jobs.append((m, n, k, A, B, x))
    want = [(A @ B, float(x @ x)) for (m, n, k, A, B, x) in jobs]
    got = [None] * 8
    errs = []

    def work(t):
        try:
            m, n, k, A, B, x = jobs[t]
            for _ in range(5):
                C = np.zeros((m, n), order="F")
                f77(lib, "dgemm_", "N", "N", m, n, k, 1.0, A, m, B, k, 0.0, C, m)
                d = f77(lib, "ddot_", x.size, x, 1, x, 1, restype=ctypes.c_double)
                got[t] = (C, d)
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join()
    assert not errs, errs
    for t in range(8):
        assert np.abs(got[t][0] - want[t][0]).max() <= 16 * 2.0 ** -53 * jobs[t][2]
        assert abs(got[t][1] - want[t][1]) <= 1e-12 * want[t][1]


@pytest.mark.parametrize("ta,tb", [("N", "N"), ("T", "N"), ("N", "T"), ("T", "T")])
def test_dgemm_first_touch_of_tracked_managed_operands(ta, tb):
    """The reference's hit path: matrices calloc'd under the interposer (tracked managed blocks), filled by the CPU, then dgemm_.
    The first call migrates them range by range (cudaMemPrefetchAsync of contiguous column ranges in the order the multiply needs
    them) with the multiply following behind, in place (staged_gemm.cuh: gemm_first_touch; the reference bulk-copies, then calls
    cuBLAS: runtime-mem.hpp:84-112).  Result within the GEMM bound of a float64 numpy product; the blocks are resident afterwards
    (a second call prefetches nothing) and both calls agree to the bound."""
    lib = g.load()
    lib.b200blas_set_options(b"pipeline_min=1000")
    try:
        m, n, k = 300, 2048 + 260, 2048 + 90
        ra, ca = (m, k) if ta == "N" else (k, m)
        rb, cb = (k, n) if tb == "N" else (n, k)
        lda, ldb, ldc = ra + 2, rb + 4, m + 6
        A = splitmix_uniform(81, (lda, ca)); B = splitmix_uniform(82, (ldb, cb)); C0 = splitmix_uniform(83, (ldc, n))
        ptrs, views = [], []
        for X in (A, B, C0):
            nb = X.size * 8
            ptr = lib.b200blas_malloc_managed(nb); assert ptr
            v = np.frombuffer((ctypes.c_char * nb).from_address(ptr), dtype=np.float64).reshape(X.shape, order="F")
            v[...] = X                                             # first touch on the CPU
            ptrs.append(ptr); views.append(v)
        s0 = g.stats()
        f77(lib, "dgemm_", ta, tb, m, n, k, 0.7, g.DevPtr(ptrs[0]), lda, g.DevPtr(ptrs[1]), ldb, 1.3, g.DevPtr(ptrs[2]), ldc)
        s1 = g.stats()
        C1 = np.array(views[2], order="F")
        assert s1["prefetch_bytes"] - s0["prefetch_bytes"] >= (ra * ca + rb * cb + m * n) * 8 * 0.95, "every operand range was migrated"
        assert s1["h2d_bytes"] == s0["h2d_bytes"] and s1["hits"] - s0["hits"] == 3
        views[2][...] = C0
        f77(lib, "dgemm_", ta, tb, m, n, k, 0.7, g.DevPtr(ptrs[0]), lda, g.DevPtr(ptrs[1]), ldb, 1.3, g.DevPtr(ptrs[2]), ldc)
        s2 = g.stats()
        C2 = np.array(views[2], order="F")
        assert s2["prefetch_bytes"] == s1["prefetch_bytes"], "resident now: nothing left to migrate"
        opA = A[:ra] if ta == "N" else A[:ra].T
        opB = B[:rb] if tb == "N" else B[:rb].T
        ref = 0.7 * (opA @ opB) + 1.3 * C0[:m]
        bound = 4 * (k + 2) * 2.0 ** -53 * (0.7 * np.linalg.norm(opA) * np.linalg.norm(opB) + 1.3 * np.linalg.norm(C0[:m]))
        for Cx in (C1, C2):
            assert np.linalg.norm(Cx[:m] - ref) <= bound and np.array_equal(Cx[m:], C0[m:])
        del views, v
        for ptr in ptrs:
            lib.b200blas_free_managed(ptr)
    finally:
        lib.b200blas_set_options(b"pipeline_min=67108864")
