"""CPU-only legs for the banded / packed / Hermitian / complex Level-2 routines:

  1. the oracle's restatements (oracle/refblas_l2x.inc) pinned against the CPU BLAS of this image (OpenBLAS) and against an
     independent numpy model on the dense logical matrix, on the shared case list (tests/l2x.py);
  2. the index logic and per-routine plans the CUDA kernels are built from (libgpublas_b200/csrc/structured.cuh), compiled
     for the host and walked grid by grid (tests/drivers/struct_emul.cpp), against the same expectations -- column-major
     (Fortran) and row-major (CBLAS) meanings;
  3. the product library's argument checks and quick returns for these routines, which run before any CUDA call.

The CUDA kernels themselves are checked on the GPU by tests/test_zz_level2_struct_gpu.py."""
import ctypes

import numpy as np
import pytest

import l2x
from helpers import f77, load_openblas, oracle_call


def _worst(cs, runner):
    worst, tag = 0.0, None
    for c in cs:
        args = c.fresh_args()
        runner(c, args)
        e = l2x.compare(args[c.out], c)
        if e > worst:
            worst, tag = e, c.tag
    return worst, tag


@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_oracle_l2x_vs_openblas_and_model(p):
    ob = load_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    cs = l2x.cases(p, sizes=(1, 2, 5, 33))

    def run_oracle(c, args):
        assert oracle_call(c.name, *args) == 0, c.tag

    def run_openblas(c, args):
        f77(ob, c.name + "_", *args)

    w, tag = _worst(cs, run_oracle)
    assert w < 1.0, ("oracle vs model", tag, w)
    w, tag = _worst(cs, run_openblas)
    assert w < 1.0, ("OpenBLAS vs model", tag, w)
    # and directly against each other, on the arrays as a whole (padding included)
    for c in cs[::3]:
        a1, a2 = c.fresh_args(), c.fresh_args()
        run_oracle(c, a1); run_openblas(c, a2)
        o1, o2 = a1[c.out].astype(np.complex128), a2[c.out].astype(np.complex128)
        assert np.abs(o1 - o2).max() <= c.tol * max(1.0, np.abs(c.expect[c.expect != c.expect.dtype.type(l2x.ROGUE)]).max()), c.tag


@pytest.mark.parametrize("rowmajor", [False, True], ids=["colmajor", "rowmajor"])
@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_kernel_index_logic_emulated(p, rowmajor):
    cs = l2x.cases(p, rowmajor=rowmajor)
    w, tag = _worst(cs, lambda c, args: l2x.emu_call(c.name, *args, rowmajor=rowmajor))
    assert w < 1.0, (tag, w)


def test_emulated_solves_cross_block_boundaries():
    """n > 32 with every reach: the diagonal-block solve + update chain of TBSV/TPSV over several blocks, both directions;
    n > 256: several panels with the grid-wide update between them (packed / wide band), and a band wider than the
    single-launch limit (reach > 1024)"""
    for p in "dz":
        cs = [c for c in l2x.cases(p, sizes=(31, 32, 33, 64, 97)) if c.name[1:] in ("tbsv", "tpsv")]
        assert len(cs) > 200
        w, tag = _worst(cs, lambda c, args: l2x.emu_call(c.name, *args))
        assert w < 1.0, (tag, w)
    cs = [c for c in l2x.cases("d", sizes=(255, 256, 257, 600)) if c.name[1:] in ("tbsv", "tpsv")]
    w, tag = _worst(cs, lambda c, args: l2x.emu_call(c.name, *args))
    assert w < 1.0, (tag, w)
    n, k = 1300, 1100
    G = l2x.well_conditioned_tri(3, n, "d")
    b = l2x.rnd(4, (n,), "d")
    for ul in "UL":
        Tk = np.where(l2x.tri_mask(n, ul, k), G, 0)
        i, j = np.indices((k + 1, n)); rows = (j - k + i) if ul == "U" else (j + i); ok = (rows >= 0) & (rows < n)
        ab = np.zeros((k + 1, n), order="F"); ab[ok] = Tk[rows[ok], j[ok]]
        for tr in "NT":
            x = b.copy()
            l2x.emu_call("dtbsv", ul, tr, "N", n, k, ab, k + 1, x, 1)
            assert np.allclose(x, np.linalg.solve(l2x.opmat(Tk, tr), b), rtol=1e-10, atol=1e-10), (ul, tr)


def test_emulated_gemv_conj_notrans():
    """CBLAS row-major ConjTrans GEMV: y = alpha*A^H*x + beta*y on a row-major m x n A"""
    lib = l2x.load_emul()
    for p in "cz":
        m, n = 37, 21
        A = l2x.rnd(1, (m, n + 2), p); A = np.ascontiguousarray(A)          # row-major, lda = n+2
        x = l2x.vec(2, m, 2, p); y = l2x.vec(3, n, -1, p)
        al, be = l2x.ALPHA[p], l2x.BETA[p]
        want = al * (A[:, :n].astype(np.complex128).conj().T @ l2x.logical(x, m, 2)) + be * l2x.logical(y, n, -1)
        # column-major view: n x m with ld = n+2
        getattr(lib, "emu_" + p + "gemv_conj")(n, m, l2x._sp(p, al), l2x._ptr(A), n + 2, l2x._ptr(x), 2, l2x._sp(p, be), l2x._ptr(y), -1)
        assert np.allclose(l2x.logical(y, n, -1), want, rtol=64 * l2x.EPS[p] * m, atol=64 * l2x.EPS[p] * m)


def test_struct_error_exits_and_quick_returns():
    """netlib INFO numbering of the new entry points and their quick returns: all decided on the host before the first CUDA
    call, so this runs without a GPU (reference: the ?BLAT2-style DCHKE tables; gemm.cc:87-127 shows the pattern)."""
    import libgpublas_b200 as g
    lib = g.load(); seen = []
    CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.c_size_t)
    cb = CB(lambda name, info, ln: seen.append((name[:6].decode(), info[0])))
    lib.b200blas_set_xerbla(cb)
    A = np.zeros((4, 4), order="F"); v = np.ones(8); Z = np.zeros((4, 4), dtype=np.complex128, order="F"); w = np.ones(8, dtype=np.complex128)
    try:
        table = [
            ("dgbmv_", [(("X", 2, 2, 1, 1, 1.0, A, 3, v, 1, 1.0, v, 1), 1), (("N", -1, 2, 1, 1, 1.0, A, 3, v, 1, 1.0, v, 1), 2), (("N", 2, -1, 1, 1, 1.0, A, 3, v, 1, 1.0, v, 1), 3),
                        (("N", 2, 2, -1, 1, 1.0, A, 3, v, 1, 1.0, v, 1), 4), (("N", 2, 2, 1, -1, 1.0, A, 3, v, 1, 1.0, v, 1), 5), (("N", 2, 2, 1, 1, 1.0, A, 2, v, 1, 1.0, v, 1), 8),
                        (("N", 2, 2, 1, 1, 1.0, A, 3, v, 0, 1.0, v, 1), 10), (("N", 2, 2, 1, 1, 1.0, A, 3, v, 1, 1.0, v, 0), 13)]),
            ("dsbmv_", [(("X", 2, 1, 1.0, A, 2, v, 1, 1.0, v, 1), 1), (("U", -1, 1, 1.0, A, 2, v, 1, 1.0, v, 1), 2), (("U", 2, -1, 1.0, A, 2, v, 1, 1.0, v, 1), 3),
                        (("U", 2, 1, 1.0, A, 1, v, 1, 1.0, v, 1), 6), (("U", 2, 1, 1.0, A, 2, v, 0, 1.0, v, 1), 8), (("U", 2, 1, 1.0, A, 2, v, 1, 1.0, v, 0), 11)]),
            ("dspmv_", [(("X", 2, 1.0, A, v, 1, 1.0, v, 1), 1), (("U", -1, 1.0, A, v, 1, 1.0, v, 1), 2), (("U", 2, 1.0, A, v, 0, 1.0, v, 1), 6), (("U", 2, 1.0, A, v, 1, 1.0, v, 0), 9)]),
            ("zhemv_", [(("X", 2, 1j, Z, 2, w, 1, 1j, w, 1), 1), (("U", -1, 1j, Z, 2, w, 1, 1j, w, 1), 2), (("U", 2, 1j, Z, 1, w, 1, 1j, w, 1), 5),
                        (("U", 2, 1j, Z, 2, w, 0, 1j, w, 1), 7), (("U", 2, 1j, Z, 2, w, 1, 1j, w, 0), 10)]),
            ("zhbmv_", [(("U", 2, 1, 1j, Z, 1, w, 1, 1j, w, 1), 6), (("U", 2, 1, 1j, Z, 2, w, 1, 1j, w, 0), 11)]),
            ("zhpmv_", [(("U", 2, 1j, Z, w, 0, 1j, w, 1), 6), (("U", 2, 1j, Z, w, 1, 1j, w, 0), 9)]),
            ("dtbmv_", [(("X", "N", "N", 2, 1, A, 2, v, 1), 1), (("U", "X", "N", 2, 1, A, 2, v, 1), 2), (("U", "N", "X", 2, 1, A, 2, v, 1), 3), (("U", "N", "N", -1, 1, A, 2, v, 1), 4),
                        (("U", "N", "N", 2, -1, A, 2, v, 1), 5), (("U", "N", "N", 2, 1, A, 1, v, 1), 7), (("U", "N", "N", 2, 1, A, 2, v, 0), 9)]),
            ("ztbsv_", [(("U", "N", "N", 2, 1, Z, 1, w, 1), 7), (("U", "N", "N", 2, 1, Z, 2, w, 0), 9)]),
            ("dtpmv_", [(("X", "N", "N", 2, A, v, 1), 1), (("U", "X", "N", 2, A, v, 1), 2), (("U", "N", "X", 2, A, v, 1), 3), (("U", "N", "N", -1, A, v, 1), 4), (("U", "N", "N", 2, A, v, 0), 7)]),
            ("ctpsv_", [(("U", "N", "N", 2, Z, w, 0), 7)]),
            ("ztrmv_", [(("U", "N", "N", 2, Z, 1, w, 1), 6), (("U", "N", "N", 2, Z, 2, w, 0), 8)]),
            ("zgeru_", [((-1, 2, 1j, w, 1, w, 1, Z, 2), 1), ((2, -1, 1j, w, 1, w, 1, Z, 2), 2), ((2, 2, 1j, w, 0, w, 1, Z, 2), 5), ((2, 2, 1j, w, 1, w, 0, Z, 2), 7), ((2, 2, 1j, w, 1, w, 1, Z, 1), 9)]),
            ("cgerc_", [((2, 2, 1j, w, 1, w, 1, Z, 1), 9)]),
            ("zher_", [(("X", 2, 1.0, w, 1, Z, 2), 1), (("U", -1, 1.0, w, 1, Z, 2), 2), (("U", 2, 1.0, w, 0, Z, 2), 5), (("U", 2, 1.0, w, 1, Z, 1), 7)]),
            ("zher2_", [(("U", 2, 1j, w, 1, w, 0, Z, 2), 7), (("U", 2, 1j, w, 1, w, 1, Z, 1), 9)]),
            ("dsyr2_", [(("X", 2, 1.0, v, 1, v, 1, A, 2), 1), (("U", 2, 1.0, v, 0, v, 1, A, 2), 5), (("U", 2, 1.0, v, 1, v, 0, A, 2), 7), (("U", 2, 1.0, v, 1, v, 1, A, 1), 9)]),
            ("dspr_", [(("X", 2, 1.0, v, 1, A), 1), (("U", -1, 1.0, v, 1, A), 2), (("U", 2, 1.0, v, 0, A), 5)]),
            ("dspr2_", [(("U", 2, 1.0, v, 1, v, 0, A), 7)]),
            ("zhpr_", [(("U", 2, 1.0, w, 0, Z), 5)]),
            ("zhpr2_", [(("U", 2, 1j, w, 0, w, 1, Z), 5), (("U", 2, 1j, w, 1, w, 0, Z), 7)]),
        ]
        for name, rows in table:
            for args, want in rows:
                seen.clear(); f77(lib, name, *args)
                assert seen == [((name[:-1].upper() + "      ")[:6], want)], (name, args[:3], seen)
        # quick returns: n == 0, alpha == 0 (& beta == 1): nothing is touched and no device is needed
        seen.clear()
        f77(lib, "dgbmv_", "N", 0, 2, 1, 1, 1.0, A, 3, v, 1, 1.0, v, 1); f77(lib, "dgbmv_", "N", 2, 2, 1, 1, 0.0, A, 3, v, 1, 1.0, v, 1)
        f77(lib, "dsbmv_", "U", 0, 1, 1.0, A, 2, v, 1, 0.0, v, 1); f77(lib, "zhpmv_", "L", 2, 0j, Z, w, 1, 1 + 0j, w, 1)
        f77(lib, "dtbsv_", "U", "N", "N", 0, 1, A, 2, v, 1); f77(lib, "ztpmv_", "U", "N", "N", 0, Z, w, 1)
        f77(lib, "zgeru_", 2, 2, 0j, w, 1, w, 1, Z, 2); f77(lib, "zher_", "U", 2, 0.0, w, 1, Z, 2); f77(lib, "dspr2_", "U", 2, 0.0, v, 1, v, 1, A)
        assert seen == [] and np.all(A == 0) and np.all(Z == 0) and np.all(v == 1) and np.all(w == 1)
    finally:
        lib.b200blas_set_xerbla(CB(0))


@pytest.mark.parametrize("p", ["d", "z"])
def test_rowmajor_model_vs_openblas_cblas(p):
    """The row-major expectations (explicit CBLAS row-major storage builders + numpy model) agree with the CPU BLAS's own
    cblas_* entry points -- so the GPU test's row-major leg compares against something pinned."""
    ob = load_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    for order, rowmajor in (("R", True), ("C", False)):
        cs = l2x.cases(p, rowmajor=rowmajor, sizes=(1, 2, 5, 33))
        w, tag = _worst(cs, lambda c, args: l2x.cblas_call(ob, c.name, order, *args))
        assert w < 1.0, (order, tag, w)


def _rotmg_inputs():
    return [(2.0, 3.0, 0.5, 0.7), (3.0, 2.0, 0.7, -0.5), (1.0, 1.0, 1.0, 1.0), (-1.0, 2.0, 1.0, 1.0), (2.0, 0.0, 1.0, 1.0), (2.0, 3.0, 1.0, 0.0),
            (1e-9, 2.0, 3.0, 1e-3), (2.0, 1e9, 1e-3, 5.0), (4.0e8, 3.0, 2.0, 1e-6), (1.0, -0.5, 1.0, 1.0), (0.5, 0.25, 3.0, -4.0), (1e10, 1e-10, 1e-4, 1e4)]


def test_level1_extras_oracle_vs_openblas():
    """ROTM / ROTMG / I?AMIN / DSDOT / SDSDOT / CSROT / ZDROT restatements against the CPU BLAS; and the product library's
    ROTMG, which is scalar host work (no device needed)."""
    import libgpublas_b200 as g
    ob = load_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    lib = g.load()
    for p, dt, tol in (("s", np.float32, 1e-6), ("d", np.float64, 1e-14)):
        for flag in (-2.0, -1.0, 0.0, 1.0):
            param = np.array([flag, 0.3, -0.4, 0.5, 0.6], dtype=dt)
            for n, ix, iy in [(1, 1, 1), (33, 2, -3), (100, -1, 1)]:
                x = l2x.vec(1, n, ix, p); y = l2x.vec(2, n, iy, p)
                x1, y1, x2, y2 = x.copy(), y.copy(), x.copy(), y.copy()
                f77(ob, p + "rotm_", n, x1, ix, y1, iy, param); oracle_call(p + "rotm", n, x2, ix, y2, iy, param, restype=None)
                assert np.allclose(x1, x2, rtol=tol, atol=tol) and np.allclose(y1, y2, rtol=tol, atol=tol), (p, flag, n)
        for (d1, d2, x1, y1) in _rotmg_inputs():
            outs = []
            for who in ("openblas", "oracle", "product"):
                a = [np.array([v], dtype=dt) for v in (d1, d2, x1)]; prm = np.full(5, 9.0, dtype=dt)
                if who == "openblas":
                    f77(ob, p + "rotmg_", a[0], a[1], a[2], np.array([y1], dtype=dt), prm)
                elif who == "product":
                    f77(lib, p + "rotmg_", a[0], a[1], a[2], np.array([y1], dtype=dt), prm)
                else:
                    oracle_call(p + "rotmg", a[0], a[1], a[2], dt(y1), prm, restype=None)
                outs.append(np.concatenate([a[0], a[1], a[2], prm]).astype(np.float64))
            assert np.allclose(outs[0], outs[1], rtol=64 * tol, atol=1e-30), (p, d1, d2, x1, y1, outs[0], outs[1])
            assert np.array_equal(outs[1], outs[2]), (p, d1, d2, x1, y1, outs[1], outs[2])
    for p in "sdcz":
        for n, inc in [(1, 1), (7, 2), (1000, 1), (1000, 3)]:
            x = l2x.vec(3, n, inc, p)
            if n > 5:
                x[0] = x[3 * inc] = x[5 * inc] * 0 + 1e-3      # planted ties for the minimum
                x[5 * inc] = 1e-3
            fn = getattr(ob, "i" + p + "amin_"); fn.restype = ctypes.c_int
            want = f77(ob, "i" + p + "amin_", n, x, inc, restype=ctypes.c_int)
            assert oracle_call("i" + p + "amin", n, x, inc) == want, (p, n, inc)
    for n, ix, iy in [(1, 1, 1), (33, 2, -3), (1001, 1, 1)]:
        x = l2x.vec(4, n, ix, "s"); y = l2x.vec(5, n, iy, "s")
        want = f77(ob, "dsdot_", n, x, ix, y, iy, restype=ctypes.c_double)
        got = oracle_call("dsdot", n, x, ix, y, iy, restype=ctypes.c_double)
        exact = float(np.dot(l2x.logical(x, n, ix).astype(np.float64), l2x.logical(y, n, iy).astype(np.float64)))
        assert abs(got - exact) <= 1e-13 * n      # netlib DSDOT: every product and the sum in double
        assert abs(got - want) <= 2e-7 * n        # OpenBLAS 0.3.15 sums blocks of its SDOT kernel in float: float-level agreement only
        sb = np.float32(0.37)
        want = f77(ob, "sdsdot_", n, sb, x, ix, y, iy, restype=ctypes.c_float)
        got = oracle_call("sdsdot", n, sb, x, ix, y, iy, restype=ctypes.c_float)
        assert abs(got - want) <= 2e-6 * max(1.0, abs(want))
    for p, rp, tol in (("c", "s", 1e-6), ("z", "d", 1e-14)):
        nm = "csrot" if p == "c" else "zdrot"
        for n, ix, iy in [(1, 1, 1), (33, 2, -3)]:
            x = l2x.vec(6, n, ix, p); y = l2x.vec(7, n, iy, p)
            x1, y1, x2, y2 = x.copy(), y.copy(), x.copy(), y.copy()
            c, s_ = l2x.DT[rp](0.6), l2x.DT[rp](0.8)
            f77(ob, nm + "_", n, x1, ix, y1, iy, c, s_); oracle_call(p + "srot", n, x2, ix, y2, iy, c, s_, restype=None)
            assert np.allclose(x1, x2, rtol=tol, atol=tol) and np.allclose(y1, y2, rtol=tol, atol=tol)


@pytest.mark.parametrize("p", ["c", "z"])
def test_cblas_complex_level3_case_list_on_openblas(p):
    """The numpy expectations of the GPU test for cblas_{c,z}symm / hemm / syr2k / herk / her2k (both layouts) hold for the CPU
    BLAS's own cblas_* entry points, so that GPU leg compares against something pinned."""
    import importlib
    ob = load_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    zz = importlib.import_module("test_zz_level2_struct_gpu")

    def call(name, *cargs):
        fn = getattr(ob, "cblas_" + name); fn.restype = None
        fn(*cargs)
    assert zz._cblas_l3_cases(call, p) < 1.0


def test_complex_rotg_host_path_vs_openblas():
    """CROTG / ZROTG are scalar host work in the product (no device): against the CPU BLAS, and the rotation must zero b."""
    import libgpublas_b200 as g
    ob = load_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    lib = g.load()
    for p, dt, rt, tol in (("c", np.complex64, np.float32, 2e-6), ("z", np.complex128, np.float64, 1e-14)):
        for (a, b) in [(3 + 4j, 1 - 2j), (-1 + 0.5j, 2 + 2j), (0j, 1 + 1j), (2 - 1j, 0j), (1e-3 + 0j, 5j)]:
            outs = []
            for L in (ob, lib):
                ca, cb = np.array([a], dtype=dt), np.array([b], dtype=dt); c = np.zeros(1, dtype=rt); s = np.zeros(1, dtype=dt)
                f77(L, p + "rotg_", ca, cb, c, s)
                outs.append((ca[0], c[0], s[0]))
            (r1, c1, s1), (r2, c2, s2) = outs
            assert abs(r1 - r2) <= tol * max(1, abs(r1)) and abs(c1 - c2) <= tol and abs(s1 - s2) <= tol, (p, a, b, outs)
            # [c s; -conj(s) c] (a, b)^T = (r, 0)
            assert abs(-np.conj(s2) * a + c2 * b) <= 8 * tol * max(1.0, abs(a) + abs(b))
            assert abs(c2 * a + s2 * b - r2) <= 8 * tol * max(1.0, abs(a) + abs(b))


def test_kernel_index_logic_fuzzed():
    """Random shapes, bandwidths, increments, triangles and operations through the emulated kernels against the oracle:
    the fixed case list walks a grid of corner cases, this walks the space between them (seeded: reproducible)."""
    rng = np.random.default_rng(20261017)
    worst = 0.0
    for it in range(300):
        p = "sdcz"[it % 4]
        dt, W = l2x.DT[p], l2x.wide(p)
        n = int(rng.integers(1, 140)); m = int(rng.integers(1, 140))
        ul = "UL"[int(rng.integers(2))]; tr = "NTC"[int(rng.integers(3))]; dg = "NU"[int(rng.integers(2))]
        ix = int(rng.choice([1, 2, -1, -3])); iy = int(rng.choice([1, 3, -2]))
        al, be = l2x.scal(p, l2x.ALPHA[p]), l2x.scal(p, l2x.BETA[p])
        kind = it % 5
        if kind == 0:
            kl = int(rng.integers(0, m)); ku = int(rng.integers(0, n)); lda = kl + ku + 1 + int(rng.integers(0, 3))
            A = l2x.rnd(it, (lda, n), p); lx, ly = (n, m) if tr == "N" else (m, n)
            x = l2x.vec(it + 1, lx, ix, p); y = l2x.vec(it + 2, ly, iy, p); y2 = y.copy()
            l2x.emu_call(p + "gbmv", tr, m, n, kl, ku, al, A, lda, x, ix, be, y, iy)
            assert oracle_call(p + "gbmv", tr, m, n, kl, ku, al, A, lda, x, ix, be, y2, iy) == 0
            got, want, scale = y, y2, max(m, n)
        elif kind == 1:
            k = int(rng.integers(0, n)); lda = k + 1 + int(rng.integers(0, 3))
            A = l2x.rnd(it, (lda, n), p); x = l2x.vec(it + 1, n, ix, p); y = l2x.vec(it + 2, n, iy, p); y2 = y.copy()
            nm = p + ("hbmv" if l2x.cplx(p) else "sbmv")
            l2x.emu_call(nm, ul, n, k, al, A, lda, x, ix, be, y, iy)
            assert oracle_call(nm, ul, n, k, al, A, lda, x, ix, be, y2, iy) == 0
            got, want, scale = y, y2, n
        elif kind == 2:
            k = int(rng.integers(0, n)); lda = k + 1 + int(rng.integers(0, 3))
            G = l2x.well_conditioned_tri(it, n, p)
            A = l2x.band_tri(np.where(l2x.tri_mask(n, ul, k), G, 0).astype(dt), ul, k, lda)
            x = l2x.vec(it + 1, n, ix, p); x2 = x.copy()
            nm = p + ("tbsv" if it % 2 else "tbmv")
            l2x.emu_call(nm, ul, tr, dg, n, k, A, lda, x, ix)
            assert oracle_call(nm, ul, tr, dg, n, k, A, lda, x2, ix) == 0
            got, want, scale = x, x2, 4 * n
        elif kind == 3:
            G = l2x.well_conditioned_tri(it, n, p)
            AP = l2x.packed(np.where(l2x.tri_mask(n, ul), G, 0).astype(dt), ul)
            x = l2x.vec(it + 1, n, ix, p); x2 = x.copy()
            nm = p + ("tpsv" if it % 2 else "tpmv")
            l2x.emu_call(nm, ul, tr, dg, n, AP, x, ix)
            assert oracle_call(nm, ul, tr, dg, n, AP, x2, ix) == 0
            got, want, scale = x, x2, 4 * n
        else:
            AP = l2x.rnd(it, (n * (n + 1) // 2,), p); AP2 = AP.copy()
            x = l2x.vec(it + 1, n, ix, p); y = l2x.vec(it + 2, n, iy, p)
            nm = p + ("hpr2" if l2x.cplx(p) else "spr2")
            l2x.emu_call(nm, ul, n, al, x, ix, y, iy, AP)
            assert oracle_call(nm, ul, n, al, x, ix, y, iy, AP2) == 0
            got, want, scale = AP, AP2, 4
        err = float(np.abs(got.astype(np.complex128) - want.astype(np.complex128)).max()) if got.size else 0.0
        tol = 16 * l2x.EPS[p] * max(scale, 4) * max(1.0, float(np.abs(want).max()) if want.size else 1.0)
        worst = max(worst, err / tol)
        assert err <= tol, (it, p, kind, n, m, ul, tr, dg, ix, iy, err, tol)
    assert worst < 1.0


@pytest.mark.parametrize("p", ["d", "z"])
def test_oracle_reproduces_the_committed_openblas_vectors(p):
    """tests/golden/level2_struct_openblas.npz (generator: tests/golden/make_golden_level2.py) holds what the CPU BLAS computed for a
    slice of the case list; the oracle must reproduce it -- no CPU BLAS needed at test time."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "level2_struct_openblas.npz"))
    cs = l2x.cases(p, sizes=(5, 33))[::7]
    assert [c.tag for c in cs] == list(gold[p + "/tags"])
    for i, c in enumerate(cs):
        args = c.fresh_args()
        assert oracle_call(c.name, *args) == 0
        want = gold["%s/%d" % (p, i)]
        rogue = want == want.dtype.type(l2x.ROGUE)
        assert np.array_equal(args[c.out][rogue], want[rogue])
        scale = max(1.0, float(np.abs(want[~rogue]).max()) if (~rogue).any() else 1.0)
        assert float(np.abs(args[c.out].astype(np.complex128) - want.astype(np.complex128)).max()) <= c.tol * scale, c.tag


def test_emulated_unit_diagonal_never_multiplies_the_diagonal():
    """A unit diagonal takes no part in the product: with Inf in x, x(j) is added as is (netlib adds it; a 0 * Inf term would be NaN),
    and whatever is stored on the diagonal -- NaN here -- is never used."""
    for p in "dz":
        n = 70
        G = l2x.well_conditioned_tri(9, n, p)
        for ul in "UL":
            T = np.where(l2x.tri_mask(n, ul), G, 0).astype(l2x.DT[p])
            T[np.arange(n), np.arange(n)] = np.nan
            AP = l2x.packed(T, ul)
            for tr in "NT":
                x = l2x.rnd(10, (n,), p); x[17] = np.inf
                x2 = x.copy()
                l2x.emu_call(p + "tpmv", ul, tr, "U", n, AP, x, 1)
                assert oracle_call(p + "tpmv", ul, tr, "U", n, AP, x2, 1) == 0
                assert np.isinf(x[17].real) and not np.isnan(x[17].real), (p, ul, tr, x[17])
                fin = np.isfinite(x2)
                assert np.array_equal(np.isfinite(x), fin) and np.allclose(x[fin], x2[fin], rtol=1e-12, atol=1e-12)


def test_cblas_rotmg_rotg_host_paths_vs_openblas():
    """cblas_?rotmg / cblas_?rotg (all four precisions of ROTG) are scalar host work in the product: compared with the CPU BLAS's
    cblas on the CPU."""
    import libgpublas_b200 as g
    ob = load_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    lib = g.load()
    for p, dt, ct, tol in (("s", np.float32, ctypes.c_float, 2e-6), ("d", np.float64, ctypes.c_double, 1e-14)):
        for (d1, d2, x1, y1) in _rotmg_inputs():
            outs = []
            for L in (ob, lib):
                a = [np.array([v], dtype=dt) for v in (d1, d2, x1)]; prm = np.full(5, 9.0, dtype=dt)
                fn = getattr(L, "cblas_" + p + "rotmg"); fn.restype = None
                fn(l2x._ptr(a[0]), l2x._ptr(a[1]), l2x._ptr(a[2]), ct(y1), l2x._ptr(prm))
                outs.append(np.concatenate(a + [prm]).astype(np.float64))
            assert np.allclose(outs[0], outs[1], rtol=64 * tol, atol=1e-30), (p, d1, d2, x1, y1, outs)
        for (a_, b_) in [(3.0, 4.0), (-3.0, 4.0), (4.0, -3.0), (0.0, 2.0), (2.0, 0.0)]:
            outs = []
            for L in (ob, lib):
                v = [np.array([x], dtype=dt) for x in (a_, b_, 0.0, 0.0)]
                fn = getattr(L, "cblas_" + p + "rotg"); fn.restype = None
                fn(*[l2x._ptr(x) for x in v])
                outs.append(np.array([x[0] for x in v], dtype=np.float64))
            assert np.allclose(outs[0], outs[1], rtol=8 * tol, atol=1e-30), (p, a_, b_, outs)
    for p, dt, rt, tol in (("c", np.complex64, np.float32, 2e-6), ("z", np.complex128, np.float64, 1e-14)):
        for (a_, b_) in [(3 + 4j, 1 - 2j), (-1 + 0.5j, 2 + 2j), (2 - 1j, 0j)]:
            outs = []
            for L in (ob, lib):     # OpenBLAS 0.3.15 exports no cblas_crotg / cblas_zrotg: its Fortran symbol is the expectation
                ca, cb = np.array([a_], dtype=dt), np.array([b_], dtype=dt); c = np.zeros(1, dtype=rt); s = np.zeros(1, dtype=dt)
                if L is ob:
                    f77(ob, p + "rotg_", ca, cb, c, s)
                else:
                    fn = getattr(L, "cblas_" + p + "rotg"); fn.restype = None
                    fn(l2x._ptr(ca), l2x._ptr(cb), l2x._ptr(c), l2x._ptr(s))
                outs.append(np.array([ca[0], c[0], s[0]], dtype=np.complex128))
            assert np.allclose(outs[0], outs[1], rtol=16 * tol, atol=16 * tol), (p, a_, b_, outs)


@pytest.mark.parametrize("p", ["d", "z", "s"])
def test_emulated_one_pass_symmetric_products(p):
    """SBMV/HBMV/SPMV/HPMV/SYMV/HEMV read the stored triangle ONCE (structured.cuh: sym_row -- every loaded element feeds its own
    row and, mirrored, its column; per-row-block column sums in tp2, sym_finish_elem).  Sizes that span several 128-row blocks and
    several column chunks, bands narrower and wider than a row block, a band too wide for the strip (the two-pass form), both
    storage orders -- the window / tp2 indexing is what this checks where there is no GPU (the device butterfly: GPU tests)."""
    names = ("sbmv", "hbmv", "spmv", "hpmv", "symv", "hemv")
    for rowmajor in (False, True):
        cs = [c for c in l2x.cases(p, rowmajor=rowmajor, sizes=(127, 129, 300), big=(1100,) if p == "d" and not rowmajor else ()) if c.name[1:] in names]
        assert len(cs) >= 30
        w, tag = _worst(cs, lambda c, args: l2x.emu_call(c.name, *args, rowmajor=rowmajor))
        assert w < 1.0, (tag, w)
