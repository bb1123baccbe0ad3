"""Host logic of the 2-D tile-partitioned GEMM (libgpublas_b200/multigpu.py) on CPU: world_size 2 and 4 over
gloo, the per-tile GEMM replaced by the oracle's ref_dgemm on raw pointers.  Checks grid choice, row/column
block ranges, the k-chunk broadcast schedule (ragged last chunk, double buffering) and the tile gather."""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import load_oracle  # noqa: E402
from libgpublas_b200.multigpu import TiledGemm, block_range, grid_for  # noqa: E402


def test_grid_and_blocks():
    assert grid_for(1) == (1, 1) and grid_for(2) == (1, 2) and grid_for(4) == (2, 2) and grid_for(8) == (2, 4)
    for total, parts in [(16384, 2), (16384, 4), (300, 2), (5, 4), (1000, 3), (129, 2)]:
        rs = [block_range(total, parts, i) for i in range(parts)]
        assert rs[0][0] == 0 and rs[-1][1] == total
        assert all(rs[i][1] == rs[i + 1][0] for i in range(parts - 1)) and all(lo <= hi for lo, hi in rs)
    assert block_range(16384, 4, 1) == (4096, 8192)


def _oracle_gemm(m, n, k, alpha, a_ptr, lda, b_ptr, ldb, beta, c_ptr, ldc):
    lib = load_oracle()
    lib.ref_dgemm.restype = ctypes.c_int
    rc = lib.ref_dgemm(ctypes.c_char(b"N"), ctypes.c_char(b"N"), m, n, k, ctypes.c_double(alpha), ctypes.c_void_p(a_ptr), lda,
                       ctypes.c_void_p(b_ptr), ldb, ctypes.c_double(beta), ctypes.c_void_p(c_ptr), ldc)
    assert rc == 0


def _worker(rank, world, port, m, n, k, kchunk, q):
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        tg = TiledGemm(m, n, k, torch.device("cpu"), rank, world, kchunk=kchunk, gemm=_oracle_gemm, c_return="sendrecv")
        A = B = None
        if rank == 0:
            gen = torch.Generator().manual_seed(3)
            A = torch.rand(m * k, dtype=torch.float64, generator=gen) * 2 - 1
            B = torch.rand(k * n, dtype=torch.float64, generator=gen) * 2 - 1
        tg.set_inputs(A, B)
        for _ in range(2):      # twice: buffers and events must be reusable
            tg.run()
        if rank == 0:
            a = A.numpy().reshape((m, k), order="F"); b = B.numpy().reshape((k, n), order="F")
            c = tg.home_c().numpy().reshape((m, n), order="F")
            err = np.abs(c - a @ b).max()
            q.put(("ok", float(err), tg.describe()))
    except Exception as e:   # surface the failure to the parent
        q.put(("fail", repr(e), ""))
        raise
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world,shape", [(2, (300, 260, 500)), (4, (257, 300, 130)), (2, (64, 5, 33))])
def test_tiled_gemm_gloo(world, shape):
    m, n, k = shape
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, n, k, 128, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, err, desc = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", err
    assert err < 1e-10, err
    assert "2d-tile" in desc


# ---------------------------------------------------------------------------------------------
# Blocked Cholesky (BASELINE.json configs[3]) over ranks: ownership, broadcast order and look-ahead ordering of
# libgpublas_b200/cholesky.py::TiledCholesky with the four block operations replaced by numpy on the shared buffer.
class _NumpyBlas:
    def __init__(self, A, n):
        self.M = A.numpy().reshape((n, n), order="F")        # shares memory with the torch buffer

    def potrf(self, j, jb):
        blk = self.M[j:j + jb, j:j + jb]
        try:
            L = np.linalg.cholesky(np.tril(blk) + np.tril(blk, -1).T)
        except np.linalg.LinAlgError:
            return 1
        blk[np.tril_indices(jb)] = L[np.tril_indices(jb)]
        return 0

    def trsm(self, j, jb, rest):
        L = np.tril(self.M[j:j + jb, j:j + jb])
        P = self.M[j + jb:j + jb + rest, j:j + jb]
        P[:] = np.linalg.solve(L, P.T).T

    def syrk(self, kcol, kb, j, jb):
        P = self.M[kcol:kcol + kb, j:j + jb]
        blk = self.M[kcol:kcol + kb, kcol:kcol + kb]
        upd = P @ P.T
        blk[np.tril_indices(kb)] -= upd[np.tril_indices(kb)]

    def gemm(self, kcol, kb, below, j, jb):
        self.M[kcol + kb:kcol + kb + below, kcol:kcol + kb] -= self.M[kcol + kb:kcol + kb + below, j:j + jb] @ self.M[kcol:kcol + kb, j:j + jb].T


def _chol_worker(rank, world, port, n, nb, q):
    from libgpublas_b200.cholesky import TiledCholesky
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        tc = TiledCholesky(n, nb, torch.device("cpu"), rank, world, blas=_NumpyBlas)
        gen = torch.Generator().manual_seed(9)
        G = torch.rand((n, n), dtype=torch.float64, generator=gen) * 2 - 1
        S = torch.tril(G, -1); S = S + S.T + n * torch.eye(n, dtype=torch.float64)
        A0 = S.numpy().copy()
        buf = S.T.contiguous().view(-1) if rank == 0 else torch.zeros(n * n, dtype=torch.float64)   # only rank 0 holds the input
        tc.set_matrix(buf)
        info = tc.run()
        L = np.tril(buf.numpy().reshape((n, n), order="F"))
        err = np.linalg.norm(L @ L.T - A0) / np.linalg.norm(A0)
        q.put(("ok", rank, info, float(err)))
    except Exception as e:
        q.put(("fail", rank, repr(e), 0.0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,nb", [(2, 300, 64), (3, 257, 50), (2, 64, 100)])
def test_tiled_cholesky_gloo(world, n, nb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_chol_worker, args=(r, world, port, n, nb, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for status, rank, info, err in res:      # every rank ends with the complete factor
        assert status == "ok", info
        assert info == 0 and err < 1e-13, (rank, info, err)


def test_syrk_strip_bounds_equal_area():
    """partitioned.PartitionedSyrk: column strips of the lower triangle with (nearly) equal areas, 128-aligned, covering n."""
    from libgpublas_b200.partitioned import strip_bounds
    for n, parts in [(32768, 8), (32768, 2), (16384, 4), (1000, 3), (128, 8)]:
        b = strip_bounds(n, parts)
        assert b[0] == 0 and b[-1] == n and len(b) == parts + 1 and all(b[i] <= b[i + 1] for i in range(parts))
        assert all(x % 128 == 0 for x in b[1:-1])
        if n >= 128 * parts * 4:
            areas = [(n - b[i]) ** 2 - (n - b[i + 1]) ** 2 for i in range(parts)]
            assert max(areas) <= 1.25 * (n * n / parts), (n, parts, b, areas)


# ---------------------------------------------------------------------------------------------------------------------
# Single-process partitioned GEMM behind the symbol (csrc/multi_gemm.cu, option devices=<n>): the hop list is pure host
# logic exported by the library (b200blas_mg_plan), so its routing is checked here without a GPU -- by EXECUTING the hops
# as numpy copies between per-device panel arrays and multiplying the panels (a CPU model of what the copy engines and
# the per-device kernels do; test infrastructure only).
def _plan(lib, ndev, m, n, host):
    cap = 4096
    buf = (ctypes.c_int * (7 * cap))()
    cnt = lib.b200blas_mg_plan(ndev, ctypes.c_longlong(m), ctypes.c_longlong(n), int(host), buf, cap)
    assert 0 <= cnt <= cap
    return [tuple(buf[7 * i + j] for j in range(7)) for i in range(cnt)]


def _geometry(lib, ndev, m, n, slot):
    out = (ctypes.c_longlong * 4)()
    lib.b200blas_mg_geometry(ndev, ctypes.c_longlong(m), ctypes.c_longlong(n), slot, out)
    return tuple(out)


@pytest.mark.parametrize("ndev", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("host", [False, True])
def test_partitioned_gemm_hop_list_delivers_every_panel_once(ndev, host):
    import libgpublas_b200 as g
    lib = g.load()
    P, Q = grid_for(ndev)
    for (m, n, k) in [(16384, 16384, 64), (9000, 20000, 32), (2048 * P + 77, 2048 * Q + 130, 16)]:
        plan = _plan(lib, ndev, m, n, host)
        geo = [_geometry(lib, ndev, m, n, s) for s in range(ndev)]
        # tiles partition C
        cover = np.zeros((m, n), dtype=np.int8) if m * n <= 4e8 else None
        assert sorted({(r0, r1) for (r0, r1, _, _) in geo})[0][0] == 0 and max(r1 for (_, r1, _, _) in geo) == m
        assert max(c1 for (_, _, _, c1) in geo) == n
        assert sum((r1 - r0) * (c1 - c0) for (r0, r1, c0, c1) in geo) == m * n
        rng = np.random.default_rng(ndev * 2 + host)
        A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n))
        panelA = [np.full((geo[s][1] - geo[s][0], k), np.nan) for s in range(ndev)]
        panelB = [np.full((k, geo[s][3] - geo[s][2]), np.nan) for s in range(ndev)]
        if not host:          # the home GPU uses its operands in place
            panelA[0][:] = A[geo[0][0]:geo[0][1]]; panelB[0][:] = B[:, geo[0][2]:geo[0][3]]
        have = set()          # (slot, kind, piece) delivered so far
        origin_elems, fwd_elems = 0, 0
        per_dev_origin = [0] * ndev
        first_kinds = []
        for (kind, gidx, piece, off, ln, src, dst) in plan:
            r0, r1, c0, c1 = geo[dst]
            assert (dst // Q if kind == 0 else dst % Q) == gidx, "the receiver must be a consumer of the piece"
            assert (dst, kind, piece) not in have, "delivered twice"
            assert not (dst == 0 and not host), "the home GPU never receives its own operands"
            if src >= 0:
                assert (src, kind, piece) in have, "forwarded before it arrived"
                blk = panelA[src][off:off + ln] if kind == 0 else panelB[src][:, off:off + ln]
                fwd_elems += blk.size
            else:
                blk = A[r0 + off:r0 + off + ln] if kind == 0 else B[:, c0 + off:c0 + off + ln]
                origin_elems += blk.size
                per_dev_origin[dst] += blk.size
            assert not np.isnan(blk).any()
            if kind == 0:
                panelA[dst][off:off + ln] = blk
            else:
                panelB[dst][:, off:off + ln] = blk
            have.add((dst, kind, piece))
            first_kinds.append(kind)
        assert first_kinds[0] == 1, "the first pieces are B band 0 (the kernel's tile schedule starts there)"
        for s in range(ndev):
            r0, r1, c0, c1 = geo[s]
            assert np.array_equal(panelA[s], A[r0:r1]) and np.array_equal(panelB[s], B[:, c0:c1]), (ndev, host, s)
        # every operand element leaves the origin at most once (exactly once, except the part only the home tile uses)
        if host:
            assert origin_elems == A.size + B.size
            shares = [x for x in per_dev_origin]
            if m == n:         # (ragged shapes with one or two bands per panel cannot split evenly)
                assert max(shares) <= 1.6 * (sum(shares) / ndev), ("every GPU pulls an equal share over its own PCIe link", shares)
        else:
            only_home_a = 0 if Q > 1 else (geo[0][1] - geo[0][0]) * k
            only_home_b = 0 if P > 1 else k * (geo[0][3] - geo[0][2])
            assert origin_elems == A.size + B.size - only_home_a - only_home_b
        naive = sum((geo[s][1] - geo[s][0]) * k + k * (geo[s][3] - geo[s][2]) for s in range(0 if host else 1, ndev))
        assert origin_elems + fwd_elems == naive, "total inbound bytes are what the tiles need -- only their route changed"
        if ndev == 8 and not host and m == n:
            assert origin_elems * 2.5 < naive, "home egress 4.3 GB instead of 11.3 GB at N = 8"
        # and the product assembled from the per-device tiles is the product
        if k <= 32:
            C = np.empty((m, n))
            for s in range(ndev):
                r0, r1, c0, c1 = geo[s]
                C[r0:r1, c0:c1] = panelA[s] @ panelB[s]
            assert np.allclose(C, A @ B, rtol=1e-12, atol=1e-12)


# ---------------------------------------------------------------------------------------------------------------------
# Partitioned ?syrk_ / ?trsm_ / ?trmm_ behind the symbol (csrc/multi_level3.cu): strip boundaries and hop lists are pure host
# logic exported by the library; executed here as numpy copies, with the per-device work done by numpy (test infrastructure).
def _ml3_plan(lib, which, ndev, n, host):
    cap = 8192
    buf = (ctypes.c_int * (7 * cap))()
    cnt = lib.b200blas_ml3_plan(which, ndev, ctypes.c_longlong(n), int(host), buf, cap)
    assert 0 <= cnt <= cap
    return [tuple(buf[7 * i + j] for j in range(7)) for i in range(cnt)]


def _ml3_strips(lib, n, ndev):
    out = (ctypes.c_longlong * (ndev + 1))()
    lib.b200blas_ml3_strips(ctypes.c_longlong(n), ndev, out)
    return list(out)


@pytest.mark.parametrize("ndev", [2, 3, 4, 8])
@pytest.mark.parametrize("host", [False, True])
def test_partitioned_syrk_strips_and_hop_list(ndev, host):
    import libgpublas_b200 as g
    lib = g.load()
    for n in [1024 * ndev, 16384, 8192 + 77]:
        b = _ml3_strips(lib, n, ndev)
        assert b[0] == 0 and b[-1] == n and all(b[i] < b[i + 1] for i in range(ndev))
        assert all(v % 128 == 0 for v in b[:-1])
        # referenced elements per strip (lower: column j holds n - j of them): equal shares within the rounding to whole CTA tiles
        area = [(b[s + 1] - b[s]) * n - (b[s + 1] * (b[s + 1] - 1) - b[s] * (b[s] - 1)) // 2 for s in range(ndev)]
        assert sum(area) == n * (n + 1) // 2
        if n >= 16384:
            assert max(area) <= 1.08 * (sum(area) / ndev), area
        plan = _ml3_plan(lib, 0, ndev, n, host)
        k = 8
        rng = np.random.default_rng(n + ndev)
        A = rng.standard_normal((n, k))
        panel = [np.full((n - b[s], k), np.nan) for s in range(ndev)]
        if not host:
            panel[0][:] = A
        have, origin, fwd = set(), 0, 0
        first_dst = []
        for (kind, strip, piece, off, ln, src, dst) in plan:
            assert kind == 0 and dst <= strip, "only devices 0..strip consume a piece of that strip"
            assert b[strip] <= off and off + ln <= b[strip + 1]
            assert (dst, piece) not in have
            assert not (dst == 0 and not host)
            if src >= 0:
                assert (src, piece) in have, "forwarded before it arrived"
                blk = panel[src][off - b[src]:off - b[src] + ln]; fwd += blk.size
            else:
                blk = A[off:off + ln]; origin += blk.size
                first_dst.append(dst)
            assert not np.isnan(blk).any()
            panel[dst][off - b[dst]:off - b[dst] + ln] = blk
            have.add((dst, piece))
        for s in range(ndev):
            assert np.array_equal(panel[s], A[b[s]:]), (ndev, host, s)
        # each row leaves the origin once (device-resident: rows of strip 0 never leave -- only the home GPU uses them)
        assert origin == (A.size if host else (n - b[1]) * k)
        assert origin + fwd == sum((n - b[s]) * k for s in range(0 if host else 1, ndev))
        # the last strip is queued first: the device with the shortest panel starts first
        assert plan[0][1] == ndev - 1
        if ndev >= 4:
            assert len(set(first_dst)) >= ndev - 1, "first receivers rotate over the devices"
        # lower and upper triangles assembled from the per-device trapezoids
        if n > 9000:
            continue
        full = A @ A.T
        L = np.zeros((n, n)); U = np.zeros((n, n))
        for s in range(ndev):
            w = b[s + 1] - b[s]
            P = panel[s]
            L[b[s]:, b[s]:b[s + 1]] = np.tril(P @ P[:w].T)
            U[b[s]:b[s + 1], b[s]:] = np.triu(P[:w] @ P.T)
        assert np.allclose(L, np.tril(full), rtol=1e-12, atol=1e-12) and np.allclose(U, np.triu(full), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("ndev", [2, 3, 4, 8])
@pytest.mark.parametrize("host", [False, True])
def test_partitioned_triangular_hop_list(ndev, host):
    """?trsm_/?trmm_: every device receives the referenced trapezoid of every column group of A exactly once."""
    import libgpublas_b200 as g
    lib = g.load()
    for na in [2048, 8192, 5000]:
        plan = _ml3_plan(lib, 1, ndev, na, host)
        have = set()
        cols = [np.zeros(na, dtype=np.int32) for _ in range(ndev)]
        origin_cols = 0
        firsts = []
        for (kind, gidx, piece, off, ln, src, dst) in plan:
            assert 0 <= off and off + ln <= na and ln > 0
            assert (dst, piece) not in have and not (dst == 0 and not host)
            if src >= 0:
                assert (src, piece) in have
            else:
                origin_cols += ln; firsts.append(dst)
            cols[dst][off:off + ln] += 1
            have.add((dst, piece))
        for s in range(0 if host else 1, ndev):
            assert (cols[s] == 1).all()
        assert origin_cols == na, "A leaves its origin once"
        if ndev >= 3:
            assert len(set(firsts)) == (ndev if host else ndev - 1)


# ---------------------------------------------------------------------------------------------------------------------
# The multi-device drivers under a CUDA stream / event simulator (tests/drivers/mgsim.cpp): multi_gemm.cu and multi_level3.cu are
# host code without kernels, compiled here unchanged with g++ against a stand-in runtime in which nothing executes when it is
# queued and a seeded policy (random / kernels first / copies first) picks the next runnable operation among all streams of all
# devices -- so a missing event wait, a buffer reused too early or a flag protocol error shows up as a wrong result or a deadlock
# on the CPU.  The single-GPU kernels are OpenBLAS calls.  This is the one-process counterpart of a world_size-N gloo test.
def test_multi_device_drivers_under_the_stream_simulator(tmp_path):
    import subprocess
    from helpers import find_openblas
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ob = find_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    drv = os.path.join(ROOT, "tests", "drivers")
    build = os.path.join(drv, "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "mgsim")
    csrc = os.path.join(ROOT, "libgpublas_b200", "csrc")
    srcs = [os.path.join(drv, "mgsim.cpp"), os.path.join(csrc, "multi_gemm.cu"), os.path.join(csrc, "multi_level3.cu")]
    deps = srcs + [os.path.join(drv, "simcuda.inc")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".h", ".cuh"))]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-Wall", "-I/usr/local/cuda/include", "-o", exe, srcs[0], "-x", "c++", srcs[1], srcs[2], "-ldl", "-lpthread"])
    env = dict(os.environ, MGSIM_OPENBLAS=ob, OPENBLAS_CORETYPE="SkylakeX", OPENBLAS_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    env["LD_LIBRARY_PATH"] = os.path.dirname(ob) + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    last = out.stdout.strip().splitlines()[-1]
    assert last.startswith("RESULT cases=") and " failed=0 " in last, last
    assert int(last.split("cases=")[1].split()[0]) >= 140
    assert "DEADLOCK" not in out.stderr
    # every device count, every routine, the three residencies were exercised
    for needle in ("devices=2", "devices=3", "devices=4", "devices=6", "devices=8", "dgemm NN", "dsyrk LN", "dtrsm LLNN", "dtrmm LUNN", "cholesky",
                   "operands=device", "operands=pinned", "operands=pageable", "operands=managed", "policy=1", "policy=2"):
        assert needle in out.stdout, needle
