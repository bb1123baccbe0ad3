"""GPU parity tests for the banded / packed / Hermitian / complex Level-2 routines (SURVEY.md section 8(f) rank 3;
libgpublas_b200/csrc/level2_struct.cu) through the C ABI: Fortran symbols and cblas_* in both layouts, on the shared case
list of tests/l2x.py (netlib ?BLAT2-style: rogue padding that must come back bit-identical, LDA = rows + 1, k / kl / ku from
0 to n-1, positive and negative increments, alpha = 0 / beta = 0 / beta = 1 corner cases, unit diagonals never read).

Expectations: the numpy model on the dense logical matrix -- which tests/test_level2_struct_cpu.py pins against the oracle
and against OpenBLAS (Fortran and cblas_*, both layouts) -- plus, for the Fortran form, the oracle directly.
Tolerance (in the case list): 16 * eps * max(n, 4) * max|expected| for products and updates, 4x that for the solves
(well-conditioned triangles)."""
import numpy as np
import pytest

import l2x
import libgpublas_b200 as g
from helpers import f77, oracle_call

pytestmark = pytest.mark.gpu


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
def test_struct_fortran_symbols_vs_model_and_oracle(p):
    lib = g.load()
    cs = l2x.cases(p, big=(257,))
    w, tag = _worst(cs, lambda c, args: f77(lib, c.name + "_", *args))
    assert w < 1.0, (tag, w)
    assert g.last_variant() == "generic_tile"
    for c in cs[::5]:      # the oracle itself, on the same call
        a1, a2 = c.fresh_args(), c.fresh_args()
        f77(lib, c.name + "_", *a1)
        assert oracle_call(c.name, *a2) == 0
        o1, o2 = a1[c.out], a2[c.out]
        rogue = o2 == o2.dtype.type(l2x.ROGUE)
        assert np.array_equal(o1[rogue], o2[rogue]), c.tag
        scale = max(1.0, float(np.abs(o2[~rogue]).max()) if (~rogue).any() else 1.0)
        assert np.abs(o1.astype(np.complex128) - o2.astype(np.complex128)).max() <= c.tol * scale, c.tag


@pytest.mark.parametrize("order", ["C", "R"])
@pytest.mark.parametrize("p", ["s", "d", "c", "z"])
def test_struct_cblas_both_layouts(p, order):
    lib = g.load()
    cs = l2x.cases(p, rowmajor=(order == "R"), sizes=(1, 2, 5, 33, 70))
    w, tag = _worst(cs, lambda c, args: l2x.cblas_call(lib, c.name, order, *args))
    assert w < 1.0, (order, tag, w)


@pytest.mark.parametrize("p", ["c", "z"])
def test_cblas_complex_gemv_all_layouts(p):
    """cblas_{c,z}gemv: column-major and row-major, NoTrans / Trans / ConjTrans (row-major ConjTrans runs as
    conjugate-no-transpose on the transposed view)"""
    import ctypes
    lib = g.load()
    dt, eps = l2x.DT[p], l2x.EPS[p]
    al, be = l2x.ALPHA[p], l2x.BETA[p]
    fn = getattr(lib, "cblas_" + p + "gemv"); fn.restype = None
    r = l2x.CREAL[p]
    for (m, n) in [(1, 1), (7, 5), (130, 77), (300, 513)]:
        M = l2x.rnd(11, (m, n), p)
        for order in "CR":
            if order == "C":
                A = np.full((m + 1, n), l2x.ROGUE, dtype=dt, order="F"); A[:m] = M; lda = m + 1
            else:
                A = np.full((m, n + 1), l2x.ROGUE, dtype=dt, order="C"); A[:, :n] = M; lda = n + 1
            for tr in "NTC":
                for (ix, iy) in [(1, 1), (2, -3)]:
                    lx, ly = (n, m) if tr == "N" else (m, n)
                    x = l2x.vec(12, lx, ix, p); y = l2x.vec(13, ly, iy, p); y0 = y.copy()
                    want = al * (l2x.opmat(M.astype(np.complex128), tr) @ l2x.logical(x, lx, ix)) + be * l2x.logical(y0, ly, iy)
                    ca, cb = (r * 2)(al.real, al.imag), (r * 2)(be.real, be.imag)
                    fn(ctypes.c_int(l2x.CBLAS_ENUM["R" if order == "R" else "Cm"]), ctypes.c_int(l2x.CBLAS_ENUM[tr]), ctypes.c_int(m), ctypes.c_int(n),
                       ctypes.byref(ca), l2x._ptr(A), ctypes.c_int(lda), l2x._ptr(x), ctypes.c_int(ix), ctypes.byref(cb), l2x._ptr(y), ctypes.c_int(iy))
                    tol = 32 * eps * max(m, n)
                    assert np.allclose(l2x.logical(y, ly, iy), want, rtol=tol, atol=tol), (p, m, n, order, tr, ix, iy)


def test_struct_solves_many_blocks_device_resident():
    """TPSV / TBSV over many 32-blocks with the operands already on the device (no staging): n = 3000, reach 5 and full"""
    import torch
    lib = g.load()
    n = 3000
    for p in "dz":
        dt = l2x.DT[p]
        G = l2x.well_conditioned_tri(5, n, p)
        b = l2x.rnd(6, (n,), p)
        for ul in "UL":
            for tr in "NTC":
                T = np.where(l2x.tri_mask(n, ul), G, 0)
                want = np.linalg.solve(l2x.opmat(T.astype(l2x.wide(p)), tr), b.astype(l2x.wide(p)))
                ap = torch.from_numpy(l2x.packed(T.astype(dt), ul)).cuda(); x = torch.from_numpy(b.copy()).cuda()
                f77(lib, p + "tpsv_", ul, tr, "N", n, ap, x, 1)
                assert np.allclose(x.cpu().numpy(), want, rtol=1e-9, atol=1e-9), (p, ul, tr, "tpsv")
                k = 5
                Tk = np.where(l2x.tri_mask(n, ul, k), G, 0)
                wantk = np.linalg.solve(l2x.opmat(Tk.astype(l2x.wide(p)), tr), b.astype(l2x.wide(p)))
                i, j = np.indices((k + 1, n))       # vectorised band storage: AB[k+i-j, j] (upper) / AB[i-j, j] (lower)
                rows = (j - k + i) if ul == "U" else (j + i)
                ok = (rows >= 0) & (rows < n)
                ab = np.zeros((k + 1, n), dtype=dt, order="F"); ab[ok] = Tk[rows[ok], j[ok]]
                abd = torch.from_numpy(np.ascontiguousarray(ab.T)).cuda()      # (n, k+1) C-order == (k+1, n) F-order in memory
                xk = torch.from_numpy(b.copy()).cuda()
                f77(lib, p + "tbsv_", ul, tr, "N", n, k, abd, k + 1, xk, 1)
                assert np.allclose(xk.cpu().numpy(), wantk, rtol=1e-9, atol=1e-9), (p, ul, tr, "tbsv")


def test_level1_extras_gpu():
    """ROTM / CSROT / ZDROT / I?AMIN / DSDOT / SDSDOT on the GPU against the oracle (pinned against OpenBLAS on the CPU)."""
    import ctypes
    lib = g.load()
    for p, dt, tol in (("s", np.float32, 2e-6), ("d", np.float64, 1e-14)):
        for flag in (-2.0, -1.0, 0.0, 1.0):
            param = np.array([flag, 0.3, -0.4, 0.5, 0.6], dtype=dt)
            for n, ix, iy in [(1, 1, 1), (33, 2, -3), (100003, 1, 1), (5000, -1, 2)]:
                x = l2x.vec(1, n, ix, p); y = l2x.vec(2, n, iy, p)
                x1, y1, x2, y2 = x.copy(), y.copy(), x.copy(), y.copy()
                f77(lib, p + "rotm_", n, x1, ix, y1, iy, param); oracle_call(p + "rotm", n, x2, ix, y2, iy, param, restype=None)
                assert np.allclose(x1, x2, rtol=tol, atol=tol) and np.allclose(y1, y2, rtol=tol, atol=tol), (p, flag, n)
    for p, rp, tol in (("c", "s", 2e-6), ("z", "d", 1e-14)):
        nm = "csrot" if p == "c" else "zdrot"
        for n, ix, iy in [(1, 1, 1), (33, 2, -3), (70001, 1, 1)]:
            x = l2x.vec(6, n, ix, p); y = l2x.vec(7, n, iy, p)
            x1, y1, x2, y2 = x.copy(), y.copy(), x.copy(), y.copy()
            c, s_ = l2x.DT[rp](0.6), l2x.DT[rp](0.8)
            f77(lib, nm + "_", n, x1, ix, y1, iy, c, s_); oracle_call(p + "srot", n, x2, ix, y2, iy, c, s_, restype=None)
            assert np.allclose(x1, x2, rtol=tol, atol=tol) and np.allclose(y1, y2, rtol=tol, atol=tol)
    for p in "sdcz":
        for n, inc in [(1, 1), (7, 2), (1000, 1), (100003, 3), (1 << 20, 1)]:
            x = l2x.vec(3, n, inc, p)
            if n > 5:
                x[(n // 2) * inc] = 1e-9; x[(n - 1) * inc] = 1e-9; x[5 * inc] = 1e-9      # ties: the first one wins
            want = oracle_call("i" + p + "amin", n, x, inc)
            assert f77(lib, "i" + p + "amin_", n, x, inc, restype=ctypes.c_int) == want, (p, n, inc)
            fn = getattr(lib, "cblas_i" + p + "amin"); fn.restype = ctypes.c_size_t
            assert fn(ctypes.c_int(n), l2x._ptr(x), ctypes.c_int(inc)) == want - 1
    for n, ix, iy in [(1, 1, 1), (33, 2, -3), (1000003, 1, 1)]:
        x = l2x.vec(4, n, ix, "s"); y = l2x.vec(5, n, iy, "s")
        exact = float(np.dot(l2x.logical(x, n, ix).astype(np.float64), l2x.logical(y, n, iy).astype(np.float64)))
        got = f77(lib, "dsdot_", n, x, ix, y, iy, restype=ctypes.c_double)
        assert abs(got - exact) <= 1e-13 * n, (n, got, exact)
        sb = np.float32(0.37)
        got = f77(lib, "sdsdot_", n, sb, x, ix, y, iy, restype=ctypes.c_float)
        assert abs(got - (exact + float(sb))) <= 2e-7 * max(1.0, abs(exact)), (n, got, exact)


def _cblas_l3_cases(lib_call, p):
    """cblas_{c,z}symm / hemm / syr2k / herk / her2k in both layouts against numpy; lib_call(name, *cargs) issues the call."""
    import ctypes
    dt, eps, r = l2x.DT[p], l2x.EPS[p], l2x.CREAL[p]
    al, be = l2x.ALPHA[p], l2x.BETA[p]
    E = l2x.CBLAS_ENUM
    I = ctypes.c_int
    cs = lambda v: (r * 2)(v.real, v.imag)
    worst = 0.0
    for order in "CR":
        def store(M, ld_extra=1):
            rows, cols = M.shape
            if order == "C":
                A = np.full((rows + ld_extra, cols), l2x.ROGUE, dtype=dt, order="F"); A[:rows] = M; return A, rows + ld_extra
            A = np.full((rows, cols + ld_extra), l2x.ROGUE, dtype=dt, order="C"); A[:, :cols] = M; return A, cols + ld_extra
        def logical(A, rows, cols):
            return A[:rows, :cols]
        o = I(E["R" if order == "R" else "Cm"])
        for (m, n) in [(5, 7), (70, 33)]:
            for side in "LR":
                ka = m if side == "L" else n
                for ul in "UL":
                    T = l2x.rnd(21, (ka, ka), p).astype(np.complex128)
                    keep = l2x.tri_mask(ka, ul)
                    B = l2x.rnd(22, (m, n), p); C0 = l2x.rnd(23, (m, n), p)
                    for herm in (False, True):
                        S = l2x.sym_from_tri(np.where(keep, T, 0), ul, herm)
                        Astore, lda = store(np.where(keep, T, l2x.ROGUE).astype(dt))
                        Bs, ldb = store(B); Cs, ldc = store(C0)
                        want = al * (S @ B.astype(np.complex128) if side == "L" else B.astype(np.complex128) @ S) + be * C0
                        ca, cb = cs(al), cs(be)
                        lib_call(p + ("hemm" if herm else "symm"), o, I(141 if side == "L" else 142), I(E[ul]), I(m), I(n), ctypes.byref(ca), l2x._ptr(Astore), I(lda),
                                 l2x._ptr(Bs), I(ldb), ctypes.byref(cb), l2x._ptr(Cs), I(ldc))
                        worst = max(worst, float(np.abs(logical(Cs, m, n) - want).max()) / (64 * eps * ka))
                        assert np.array_equal(Cs == dt(l2x.ROGUE), store(C0)[0] == dt(l2x.ROGUE))
        for (n, k) in [(6, 4), (65, 37)]:
            for ul in "UL":
                keep = l2x.tri_mask(n, ul)
                for tr in "NC":      # herk / her2k: N or C;  syr2k: N or T
                    shp = (n, k) if tr == "N" else (k, n)
                    A = l2x.rnd(24, shp, p); B = l2x.rnd(25, shp, p); C0 = l2x.rnd(26, (n, n), p)
                    Aw, Bw = A.astype(np.complex128), B.astype(np.complex128)
                    Cw = C0.astype(np.complex128)
                    As, lda = store(A); Bs, ldb = store(B)
                    # HERK (real alpha, beta): the imaginary part of the diagonal of C is taken as zero on input and output
                    ra, rb = 0.7, 1.3
                    Ch = Cw.copy(); Ch[np.arange(n), np.arange(n)] = Ch.diagonal().real
                    want = ra * (Aw @ Aw.conj().T if tr == "N" else Aw.conj().T @ Aw) + rb * Ch
                    Cs, ldc = store(C0)
                    lib_call(p + "herk", o, I(E[ul]), I(E[tr]), I(n), I(k), r(ra), l2x._ptr(As), I(lda), r(rb), l2x._ptr(Cs), I(ldc))
                    got = logical(Cs, n, n)
                    worst = max(worst, float(np.abs(np.where(keep, got - want, 0)).max()) / (64 * eps * k))
                    assert np.array_equal(np.where(keep, 0, got), np.where(keep, 0, C0)), "herk wrote outside the triangle"
                    want = al * (Aw @ Bw.conj().T if tr == "N" else Aw.conj().T @ Bw); want = want + want.conj().T + rb * Ch
                    Cs, ldc = store(C0); ca = cs(al)
                    lib_call(p + "her2k", o, I(E[ul]), I(E[tr]), I(n), I(k), ctypes.byref(ca), l2x._ptr(As), I(lda), l2x._ptr(Bs), I(ldb), r(rb), l2x._ptr(Cs), I(ldc))
                    got = logical(Cs, n, n)
                    worst = max(worst, float(np.abs(np.where(keep, got - want, 0)).max()) / (64 * eps * k))
                    trs = "N" if tr == "N" else "T"
                    want = al * (Aw @ Bw.T + Bw @ Aw.T if tr == "N" else Aw.T @ Bw + Bw.T @ Aw) + be * Cw
                    Cs, ldc = store(C0); ca, cb = cs(al), cs(be)
                    lib_call(p + "syr2k", o, I(E[ul]), I(E[trs]), I(n), I(k), ctypes.byref(ca), l2x._ptr(As), I(lda), l2x._ptr(Bs), I(ldb), ctypes.byref(cb), l2x._ptr(Cs), I(ldc))
                    got = logical(Cs, n, n)
                    worst = max(worst, float(np.abs(np.where(keep, got - want, 0)).max()) / (64 * eps * k))
    return worst


@pytest.mark.parametrize("p", ["c", "z"])
def test_cblas_complex_level3_both_layouts(p):
    lib = g.load()

    def call(name, *cargs):
        fn = getattr(lib, "cblas_" + name); fn.restype = None
        fn(*cargs)
    assert _cblas_l3_cases(call, p) < 1.0


@pytest.mark.parametrize("p", ["d", "z"])
def test_struct_against_committed_openblas_vectors(p):
    """The GPU results against tests/golden/level2_struct_openblas.npz -- what the CPU BLAS behind the reference's interposer
    computed for a slice of the case list (generator: tests/golden/make_golden_level2.py)."""
    import os
    lib = g.load()
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "level2_struct_openblas.npz"))
    cs = l2x.cases(p, sizes=(5, 33))[::7]
    assert [c.tag for c in cs] == list(gold[p + "/tags"])
    for i, c in enumerate(cs):
        args = c.fresh_args()
        f77(lib, c.name + "_", *args)
        want = gold["%s/%d" % (p, i)]
        rogue = want == want.dtype.type(l2x.ROGUE)
        assert np.array_equal(args[c.out][rogue], want[rogue]), c.tag
        scale = max(1.0, float(np.abs(want[~rogue]).max()) if (~rogue).any() else 1.0)
        assert float(np.abs(args[c.out].astype(np.complex128) - want.astype(np.complex128)).max()) <= c.tol * scale, c.tag


def test_cblas_level1_wrappers_vs_cpu_blas():
    """cblas_ forms of the complex copy / swap / scal / asum, cblas_?rotm, cblas_csrot / zdrot, cblas_dsdot / sdsdot, cblas_i?amin
    (thin wrappers over the Fortran entry points tested above) against the same calls on the CPU BLAS's cblas."""
    import ctypes
    from helpers import load_openblas
    ob = load_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    lib = g.load()
    I, F, D = ctypes.c_int, ctypes.c_float, ctypes.c_double
    n, ix, iy = 1000, 2, -3

    def both(name, build_args, outs, restype=None, tol=0.0):
        """run cblas_<name> on fresh copies through OpenBLAS and through the library; compare the arrays in `outs` and the return value"""
        res = []
        for L in (ob, lib):
            arrs, cargs = build_args()
            fn = getattr(L, "cblas_" + name); fn.restype = restype
            r = fn(*cargs)
            res.append((r, [arrs[k].copy() for k in outs]))
        (r0, a0), (r1, a1) = res
        if restype is not None:
            assert abs(r0 - r1) <= tol * max(1.0, abs(r0)), (name, r0, r1)
        for u, v in zip(a0, a1):
            assert np.abs(u.astype(np.complex128) - v.astype(np.complex128)).max() <= tol * max(1.0, float(np.abs(u).max())), name

    for p, rt, R in (("c", F, "s"), ("z", D, "d")):
        eps = l2x.EPS[p]

        def xy():
            return {"x": l2x.vec(1, n, ix, p), "y": l2x.vec(2, n, iy, p)}
        both(p + "copy", lambda: (lambda a: (a, [I(n), l2x._ptr(a["x"]), I(ix), l2x._ptr(a["y"]), I(iy)]))(xy()), ["y"])
        both(p + "swap", lambda: (lambda a: (a, [I(n), l2x._ptr(a["x"]), I(ix), l2x._ptr(a["y"]), I(iy)]))(xy()), ["x", "y"])
        al = (l2x.CREAL[p] * 2)(0.7, -0.9)
        both(p + "scal", lambda: (lambda a: (a, [I(n), ctypes.byref(al), l2x._ptr(a["x"]), I(ix)]))(xy()), ["x"], tol=4 * eps)
        both(p + ("sscal" if p == "c" else "dscal"), lambda: (lambda a: (a, [I(n), rt(1.7), l2x._ptr(a["x"]), I(ix)]))(xy()), ["x"], tol=4 * eps)
        both(("scasum" if p == "c" else "dzasum"), lambda: (lambda a: (a, [I(n), l2x._ptr(a["x"]), I(ix)]))(xy()), [], restype=rt, tol=n * eps)
        both("i" + p + "amin", lambda: (lambda a: (a, [I(n), l2x._ptr(a["x"]), I(ix)]))(xy()), [], restype=ctypes.c_size_t)
    for p, rt in (("s", F), ("d", D)):
        eps = l2x.EPS[p]
        prm = np.array([-1.0, 0.3, -0.4, 0.5, 0.6], dtype=l2x.DT[p])

        def xy():
            return {"x": l2x.vec(3, n, ix, p), "y": l2x.vec(4, n, iy, p)}
        both(p + "rotm", lambda: (lambda a: (a, [I(n), l2x._ptr(a["x"]), I(ix), l2x._ptr(a["y"]), I(iy), l2x._ptr(prm)]))(xy()), ["x", "y"], tol=8 * eps)
        both("i" + p + "amin", lambda: (lambda a: (a, [I(n), l2x._ptr(a["x"]), I(ix)]))(xy()), [], restype=ctypes.c_size_t)

    def sxy():
        return {"x": l2x.vec(5, n, ix, "s"), "y": l2x.vec(6, n, iy, "s")}
    both("dsdot", lambda: (lambda a: (a, [I(n), l2x._ptr(a["x"]), I(ix), l2x._ptr(a["y"]), I(iy)]))(sxy()), [], restype=D, tol=1e-6)
    both("sdsdot", lambda: (lambda a: (a, [I(n), F(0.37), l2x._ptr(a["x"]), I(ix), l2x._ptr(a["y"]), I(iy)]))(sxy()), [], restype=F, tol=1e-5)


def test_aligned_allocators_under_preload(tmp_path):
    """posix_memalign / aligned_alloc / memalign / valloc under LD_PRELOAD: blocks >= the threshold whose managed base
    satisfies the alignment are tracked (in place for BLAS), contents survive realloc, malloc_usable_size answers from the
    registry, everything frees cleanly.  (Sorted here, after the established preload tests.)"""
    from test_preload import build_driver, fields, run
    exe = build_driver("aligned_allocs")
    out, _ = run(exe, preload=True, cwd=str(tmp_path), timeout=120)
    r = fields([l for l in out.splitlines() if l.startswith("RESULT")][0])
    assert r["ok"] == "1" and int(r["tracked"]) >= 20, out       # 70000 B / 1 MiB / 4 MiB requests at alignments <= 256 at the very least
    out, _ = run(exe, preload=True, env_extra={"BLAS2CUDA_OPTIONS": "heuristic=false"}, cwd=str(tmp_path), timeout=120)
    assert "RESULT ok=1 tracked=0" in out


def test_ref_micro_tests_under_preload(tmp_path):
    """The reference's tests/c micro-tests (copy, dsdot, csrot, sgbmv, strmv, dtrsm, sgemm, chemm; restated in tests/drivers/ref_micro.c)
    under LD_PRELOAD=libb200blas.so against the same binary on the CPU BLAS: every call is intercepted (its calloc'd matrix is a
    tracked managed block), results agree, and the closed forms hold exactly where the fill has one."""
    from test_preload import run_ref_micro
    cpu = run_ref_micro(tmp_path, preload=False)
    gpu = run_ref_micro(tmp_path, preload=True)
    for name in ("copy", "rot", "dsdot", "gemm"):
        assert float(gpu[name][1]["closed_form_err"]) == 0.0, (name, gpu[name][1])      # exact on these inputs (DSDOT: all in double)
    for name, (arr, f) in gpu.items():
        assert f["tracked"] == "1" or name == "dsdot", (name, f)        # dsdot's 12 KB vectors stay on the heap (below the threshold)
        ref = cpu[name][0]
        assert arr.shape == ref.shape
        scale = max(1.0, float(np.abs(ref).max()))
        # float sums of n terms in different orders: n * eps(float) = 400 * 6e-8 = 2.4e-5 of the largest result; f64 solve: 1e-12
        tol = {"copy": 0.0, "rot": 0.0, "dsdot": 1e-7, "gbmv": 1e-6, "trmv": 3e-5, "trsm": 1e-12, "gemm": 0.0, "hemm": 3e-5}[name]
        assert float(np.abs(arr.astype(np.complex128) - ref.astype(np.complex128)).max()) <= tol * scale, name


def test_threaded_allocator_churn_under_preload(tmp_path):
    """tests/drivers/allocs_mt.c with the default size heuristic: 8 threads, every 10th block >= 64 KiB comes from the managed
    allocator (registry insert / lookup / remove and cudaMallocManaged / cudaFree from several threads at once, blocks freed on
    a thread other than the allocating one)."""
    from test_preload import build_driver, fields, run
    exe = build_driver("allocs_mt")
    # 8 x 60 big blocks; the malloc / calloc / posix_memalign ones (3 of the 4 random kinds, counted by the driver as big=) must ALL be
    # managed -- also the ones requested while another thread is still bringing the device up (round 1: 54 of ~360) -- a realloc
    # that grows a small heap block stays on the heap like the reference's (lib/obj_tracker.c:902-946).
    # Second run: thread 3 makes the first big allocation while threads 0-2 and 4-7 are mid-loop.
    for args in ([8, 600, 10], [8, 600, 10, 3]):
        out, _ = run(exe, args, preload=True, cwd=str(tmp_path), timeout=300)
        r = fields([l for l in out.splitlines() if l.startswith("RESULT")][0])
        assert r["ok"] == "1" and int(r["big"]) > 300 and int(r["tracked_seen"]) >= int(r["big"]), out


def test_one_pass_symmetric_products_larger_device_resident():
    """SPMV / SYMV / SBMV / HEMV / HPMV / HBMV at sizes that span many row blocks and column chunks of the one-pass kernel
    (level2_struct.cu: sympart_kernel -- every stored element read once, mirrored sums by a transposing warp butterfly), operands
    on the device, against a float64 / complex128 numpy product.  Bar: |y - y_ref| <= 16 n eps (|alpha| |S| |x| + |beta| |y0|)
    element-wise (the GEMV bound with c = 16)."""
    import torch
    lib = g.load()
    rng = np.random.default_rng(11)
    for p, n, k in (("d", 3001, 200), ("z", 1409, 77), ("s", 2050, 5)):
        dt = {"d": np.float64, "z": np.complex128, "s": np.float32}[p]
        eps = 2.0 ** -24 if p == "s" else 2.0 ** -53
        cplx = p == "z"
        def rnd(shape):
            a = rng.uniform(-1, 1, shape)
            return (a + 1j * rng.uniform(-1, 1, shape)).astype(dt) if cplx else a.astype(dt)
        M = rnd((n, n))
        x = rnd(n); y0 = rnd(n)
        alpha, beta = ((0.7 - 0.9j), (1.3 - 1.1j)) if cplx else (0.7, 1.3)
        fl = np.float64 if p in "dz" else np.float32
        def dev(a):
            return torch.from_numpy(np.ascontiguousarray(a).view(fl).copy()).cuda()
        for ul in "UL":
            tri = np.triu(M) if ul == "U" else np.tril(M)
            if cplx:
                tri[np.diag_indices(n)] = tri[np.diag_indices(n)].real + 5j      # the imaginary part of a Hermitian diagonal is never used
                S = tri + tri.conj().T; S[np.diag_indices(n)] = tri[np.diag_indices(n)].real
            else:
                S = tri + tri.T - np.diag(np.diag(tri))
            hi = np.complex128 if cplx else np.float64
            ref = alpha * (S.astype(hi) @ x.astype(hi)) + beta * y0.astype(hi)
            bound = 16 * n * eps * (abs(alpha) * (np.abs(S).astype(np.float64) @ np.abs(x).astype(np.float64)) + abs(beta) * np.abs(y0))
            # full storage (lda = n + 3), the unreferenced triangle poisoned
            Af = np.full((n + 3, n), np.nan, dtype=dt, order="F"); Af[:n][(np.triu if ul == "U" else np.tril)(np.ones((n, n), bool))] = tri[(np.triu if ul == "U" else np.tril)(np.ones((n, n), bool))]
            # packed by columns
            ap = np.concatenate([tri[: j + 1, j] if ul == "U" else tri[j:, j] for j in range(n)]).astype(dt)
            xd = dev(x)
            for name, a_args in ((("hemv" if cplx else "symv"), (dev(Af.ravel(order="F")), n + 3)), (("hpmv" if cplx else "spmv"), (dev(ap),))):
                yd = dev(y0)
                f77(lib, p + name + "_", ul, n, alpha, *a_args, xd, 1, beta, yd, 1)
                torch.cuda.synchronize()
                got = yd.cpu().numpy().view(dt)
                assert np.all(np.abs(got.astype(hi) - ref) <= bound), (p, name, ul, float(np.abs(got - ref).max()))
            # band: keep k off-diagonals
            i, j = np.indices((n, n))
            bt = np.where(np.abs(i - j) <= k, tri, 0)
            Sb = np.where(np.abs(i - j) <= k, S, 0)
            AB = np.full((k + 2, n), np.nan, dtype=dt, order="F")
            for jj in range(n):
                if ul == "U":
                    lo = max(0, jj - k); AB[k + lo - jj: k + 1, jj] = bt[lo: jj + 1, jj]
                else:
                    hi_r = min(n, jj + k + 1); AB[0: hi_r - jj, jj] = bt[jj: hi_r, jj]
            refb = alpha * (Sb.astype(hi) @ x.astype(hi)) + beta * y0.astype(hi)
            yd = dev(y0)
            f77(lib, p + ("hbmv" if cplx else "sbmv") + "_", ul, n, k, alpha, dev(AB.ravel(order="F")), k + 2, xd, 1, beta, yd, 1)
            torch.cuda.synchronize()
            got = yd.cpu().numpy().view(dt)
            assert np.all(np.abs(got.astype(hi) - refb) <= bound), (p, "band", ul, float(np.abs(got - refb).max()))
