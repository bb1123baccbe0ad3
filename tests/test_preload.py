"""The drop-in boundary itself: C programs that call BLAS and malloc through the PLT, run (a) plain on the
CPU BLAS they are linked with and (b) under LD_PRELOAD=libb200blas.so (reference scripts/blas2cuda.sh:27,
tests/netlib/test.py:28).  CPU-only part: the drivers build and run on the CPU BLAS, and the library
exports every symbol include/b200blas.h declares."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helpers import LIB_PATH, ROOT, find_openblas

HAS_GPU = os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")

DRV = os.path.join(ROOT, "tests", "drivers")
BUILD = os.path.join(DRV, "_build")


def build_driver(name):
    ob = find_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, name)
    src = os.path.join(DRV, name + ".c")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        d = os.path.dirname(ob)
        subprocess.check_call(["gcc", "-O2", "-pthread", "-o", exe, src, "-L" + d, "-l:" + os.path.basename(ob), "-Wl,-rpath," + d, "-Wl,--disable-new-dtags",
                               "-Wl,--allow-shlib-undefined", "-ldl", "-lm"])
    return exe


def run(exe, args=(), preload=False, env_extra=None, cwd=None, timeout=300):
    env = dict(os.environ)
    env.setdefault("OPENBLAS_CORETYPE", "SkylakeX")
    env["OPENBLAS_NUM_THREADS"] = str(os.cpu_count() or 1)
    ob = find_openblas()
    if ob:   # OpenBLAS's private libgfortran / libquadmath live beside it
        env["LD_LIBRARY_PATH"] = os.path.dirname(ob) + ":" + env.get("LD_LIBRARY_PATH", "")
    if preload:
        env["LD_PRELOAD"] = LIB_PATH
    if env_extra:
        env.update(env_extra)
    out = subprocess.run([exe] + [str(a) for a in args], env=env, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert out.returncode == 0, (out.returncode, out.stdout[-2000:], out.stderr[-2000:])
    return out.stdout, out.stderr


def fields(line):
    return dict(kv.split("=", 1) for kv in line.split()[1:])


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200blas.h")).read()
    declared = set(re.findall(r"^B200_API [^;(]*?(\w+)\(", hdr, re.M)) | {"malloc", "calloc", "realloc", "free"}
    assert len(declared) > 100
    syms = subprocess.check_output(["nm", "-D", "--defined-only", LIB_PATH], text=True)
    exported = {l.split()[-1] for l in syms.splitlines() if " T " in l or " W " in l}
    missing = sorted(declared - exported)
    assert not missing, missing
    # and it loads (no unresolved dependencies) on a machine without a GPU
    import libgpublas_b200 as g
    assert g.load().b200blas_version() >= 100


def test_coverage_document_matches_the_built_library(tmp_path):
    """COVERAGE.md is generated from the library's exports (tools/gen_coverage.py): regenerating it must reproduce the committed file."""
    committed = open(os.path.join(ROOT, "COVERAGE.md")).read()
    try:
        subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_coverage.py")], stdout=subprocess.DEVNULL)
        assert open(os.path.join(ROOT, "COVERAGE.md")).read() == committed
    finally:
        open(os.path.join(ROOT, "COVERAGE.md"), "w").write(committed)


def test_public_header_is_plain_c(tmp_path):
    """include/b200blas.h is the C-ABI contract: it must compile on its own as C99 (pedantic) and as C++11."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "b200blas.h"\nint main(void) { return (int)sizeof(struct b200blas_stats) == 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I" + inc, str(src)])
    cpp = tmp_path / "hdr.cpp"
    cpp.write_text(src.read_text())
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I" + inc, str(cpp)])


def test_so_is_executable_and_prints_help():
    """reference entry.c:4-11 / meson.build:25: running the .so prints the option help."""
    out = subprocess.run([LIB_PATH], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "BLAS2CUDA_OPTIONS" in out.stderr and "heuristic=" in out.stderr


def test_gemm_fixture_on_cpu_blas():
    exe = build_driver("gemm_fixture")
    out, _ = run(exe, [256, 3])
    res = [fields(l) for l in out.splitlines() if l.startswith("RESULT")]
    assert len(res) == 2 and all(float(r["max_abs_err"]) == 0.0 for r in res)
    assert "STATS none" in out


def test_cg_chain_on_cpu_blas():
    exe = build_driver("cg_chain")
    out, _ = run(exe, [512, 10])
    r = fields([l for l in out.splitlines() if l.startswith("RESULT")][0])
    assert float(r["rnorm"]) < 1e-10


def test_allocs_on_glibc():
    exe = build_driver("allocs")
    out, _ = run(exe)
    assert "RESULT ok=1 tracked=0" in out


def test_aligned_allocators_fall_through_without_a_device(tmp_path):
    """posix_memalign / aligned_alloc / memalign / valloc / malloc_usable_size are interposed (the reference leaves them to
    glibc).  With heuristic=false nothing is ever placed in managed memory, so the interposers' pass-through path
    runs under LD_PRELOAD on a machine without a GPU; the managed path is tests/test_zz_level2_struct_gpu.py's."""
    exe = build_driver("aligned_allocs")
    out, _ = run(exe)
    assert "RESULT ok=1 tracked=0" in out
    out, _ = run(exe, preload=True, env_extra={"BLAS2CUDA_OPTIONS": "heuristic=false"}, cwd=str(tmp_path), timeout=90)
    assert "RESULT ok=1 tracked=0" in out


def test_threaded_allocator_churn_passes_through_without_a_device(tmp_path):
    """16 threads x 20000 malloc / calloc / realloc / posix_memalign / free with blocks freed on other threads, under LD_PRELOAD
    with heuristic=false: the interposers' bookkeeping (re-entrancy slots, range pre-filter, pass-through) under contention.
    The managed path of the same driver runs in the GPU suite."""
    exe = build_driver("allocs_mt")
    out, _ = run(exe, [16, 20000, 50])
    assert "RESULT ok=1" in out
    out, _ = run(exe, [16, 20000, 50], preload=True, env_extra={"BLAS2CUDA_OPTIONS": "heuristic=false"}, cwd=str(tmp_path), timeout=300)
    assert "RESULT ok=1 threads=16 iters=20000 tracked_seen=0" in out


def build_tracker_mock():
    """tracker.cpp + tests/drivers/tracker_mock.cpp (a stand-in for cudaMallocManaged / cudaFree) -> _build/libtracker_mock.so"""
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libtracker_mock.so")
    srcs = [os.path.join(ROOT, "libgpublas_b200", "csrc", "tracker.cpp"), os.path.join(DRV, "tracker_mock.cpp")]
    deps = srcs + [os.path.join(ROOT, "libgpublas_b200", "csrc", "tracker.h")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-I/usr/local/cuda/include", "-o", out] + srcs
                              + ["-ldl", "-lpthread"])
    return out


def test_managed_path_bookkeeping_with_mocked_cuda(tmp_path):
    """The tracker's MANAGED path without a GPU: tracker.cpp linked against a mock of the CUDA calls it makes (mmap for
    cudaMallocManaged), preloaded into the allocator drivers.  Same expectations as the GPU tests of the real library:
    blocks >= 64 KiB are tracked, interior pointers resolve, realloc keeps contents, aligned allocators hand out tracked blocks
    whose base satisfies the alignment, and 16 threads x 20000 allocations with cross-thread frees keep the registry consistent
    (heuristic=true: every one of the 320000 blocks goes through registry insert / remove)."""
    mock = build_tracker_mock()

    def run_mock(name, args=(), heuristic=None):
        env = dict(os.environ, LD_PRELOAD=mock)
        if heuristic:
            env["TRACKER_MOCK_HEURISTIC"] = heuristic
        out = subprocess.run([build_driver(name)] + [str(a) for a in args], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, (name, out.stdout[-1000:], out.stderr[-1000:])
        return fields([l for l in out.stdout.splitlines() if l.startswith("RESULT")][0])

    r = run_mock("allocs")
    assert r["ok"] == "1" and int(r["tracked"]) > 300
    assert run_mock("allocs", heuristic="true")["tracked"] == "512" and run_mock("allocs", heuristic="false")["tracked"] == "0"
    r = run_mock("aligned_allocs")
    assert r["ok"] == "1" and int(r["tracked"]) >= 20
    r = run_mock("allocs_mt", [8, 600, 10])
    assert r["ok"] == "1" and int(r["big"]) > 300 and int(r["tracked_seen"]) >= int(r["big"])      # EVERY direct big request is tracked
    r = run_mock("allocs_mt", [16, 20000, 50], heuristic="true")
    assert r["ok"] == "1" and r["tracked_seen"] == "320000"


def test_bring_up_window_tracks_every_qualifying_allocation(tmp_path):
    """Round-1 defect (GPUTEST_r01: tracked_seen=54 of ~360): application threads that asked for a qualifying block while another
    thread was bringing the device up were silently served from the heap, and threads created during the window were never tracked.
    The mock's bring-up takes 700 ms and, like the CUDA driver, creates a helper thread (which creates another) that allocates
    large blocks.  Required: every big block requested directly by an application thread is managed -- whichever thread triggers the
    bring-up, also when the other threads are mid-loop at that moment -- and none of the bring-up's own allocations is
    (reference: once tracking is on every qualifying allocation is tracked, lib/obj_tracker.c:789-840; CUDA's own allocations
    are excluded, :352-424)."""
    mock = build_tracker_mock()
    exe = build_driver("allocs_mt")
    for args in ([8, 600, 10], [8, 600, 10, 3], [8, 600, 10, 0], [32, 100, 5, 31]):
        env = dict(os.environ, LD_PRELOAD=mock, TRACKER_MOCK_INIT_MS="700")
        out = subprocess.run([exe] + [str(a) for a in args], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, (args, out.stdout[-1000:], out.stderr[-1000:])
        r = fields([l for l in out.stdout.splitlines() if l.startswith("RESULT")][0])
        assert r["ok"] == "1" and int(r["big"]) > 100 and int(r["tracked_seen"]) >= int(r["big"]), (args, out.stdout)
        assert "MOCK helper_tracked=0 helper_child_tracked=0" in out.stdout, out.stdout


def test_no_device_is_not_fatal_for_the_allocator(tmp_path):
    """A process that only allocates must survive the preload on a machine without a GPU (`LD_PRELOAD=libb200blas.so
    python3 -c 'bytearray(1<<20)'` aborted): the failed bring-up switches tracking off and the heap serves the process.  Checked
    with the mock (bring-up reports failure) and, when this machine has no GPU, with the real library; a BLAS call without a
    device stays fatal (no CPU fallback)."""
    mock = build_tracker_mock()
    env = dict(os.environ, LD_PRELOAD=mock, TRACKER_MOCK_NO_DEVICE="1")
    out = subprocess.run([build_driver("allocs_mt"), "8", "600", "10"], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "RESULT ok=1" in out.stdout and "tracked_seen=0" in out.stdout, (out.stdout, out.stderr)
    if not HAS_GPU:
        env = dict(os.environ, LD_PRELOAD=LIB_PATH)
        out = subprocess.run([build_driver("allocs_mt"), "8", "600", "10"], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "RESULT ok=1" in out.stdout and "tracked_seen=0" in out.stdout, (out.stdout, out.stderr)
        out = subprocess.run([sys.executable, "-c", "bytearray(1<<20); print('alive')"], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "alive" in out.stdout, (out.stdout, out.stderr)


def test_options_grammar_without_a_device(tmp_path):
    """BLAS2CUDA_OPTIONS keeps the reference's grammar (blas2cuda.c:59-124): ';'-separated keys, `help` prints the option list,
    an unknown key is reported and ignored, an unknown heuristic is fatal.  heuristic=false keeps every allocation on the heap,
    so the preloaded program runs to completion on a machine without a GPU."""
    exe = build_driver("allocs")
    env = dict(os.environ, LD_PRELOAD=LIB_PATH, BLAS2CUDA_OPTIONS="help;heuristic=false;no_such_option;threshold=4096")
    out = subprocess.run([exe], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "RESULT ok=1 tracked=0" in out.stdout
    assert "heuristic=<val>" in out.stderr and "unknown option 'no_such_option'" in out.stderr and "selecting heuristic false" in out.stderr
    env["BLAS2CUDA_OPTIONS"] = "heuristic=sometimes"
    out = subprocess.run([exe], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert out.returncode != 0 and "unsupported heuristic 'sometimes'" in out.stderr
    assert not os.path.exists(os.path.join(str(tmp_path), "statistics.csv"))      # written only when a BLAS call was served


# one trace line of the reference's TRACE_OUTPUT format (lib/obj_tracker.c:426-483)
TRACE_LINE = re.compile(r"^([TUC]) #(\d+) \[(0x[0-9a-f]+)\] fun=\[(\w+)\] reqsize=\[(\d+)\] tid=\[\d+\] time=\[\d+s\+\d+ns\] uid=\[(\d+)\]$")


def test_trace_line_format_matches_the_real_reference_tracer(tmp_path):
    """oracle/_ref/libobjtracker.so is the reference's own stand-alone tracer compiled from its sources where they lie
    (oracle/build_ref.sh).  Preloaded into the allocator-churn driver it prints real T/U lines: every one of them must parse
    with the pattern the GPU test applies to the lines libb200blas.so prints under BLAS2CUDA_OPTIONS=trace, and obey the same
    invariants (T/U pair up by uid, U repeats the T's address and size)."""
    tracer = os.path.join(ROOT, "oracle", "_ref", "libobjtracker.so")
    if not os.path.exists(tracer):
        pytest.skip("oracle/_ref/libobjtracker.so not built (needs /root/reference)")
    exe = build_driver("allocs")
    env = dict(os.environ, LD_PRELOAD=tracer, OBJTRACKER_OPTIONS="")
    out = subprocess.run([exe], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=120)
    assert "RESULT ok=1" in out.stdout
    lines = [l for l in out.stdout.splitlines() if l[:2] in ("T ", "U ", "C ")]
    assert len(lines) > 1000
    ev = [TRACE_LINE.match(l) for l in lines]
    assert all(ev), [l for l, m in zip(lines, ev) if not m][:3]
    live = {}
    for m in ev:
        kind, nth, addr, fun, size, uid = m.groups()
        if kind == "T":
            assert fun in ("malloc", "calloc", "realloc")
            live[uid] = (addr, size)
        elif kind == "U":
            assert live.pop(uid) == (addr, size)


REF_MICRO = [("copy", 20000, np.float32), ("dsdot", 3000, np.float64), ("rot", 20000, np.complex64), ("gbmv", 400, np.float32),
             ("trmv", 400, np.float32), ("trsm", 300, np.float64), ("gemm", 256, np.float32), ("hemm", 200, np.complex64)]


def run_ref_micro(tmp_path, preload):
    """the reference's tests/c micro-tests (restated in tests/drivers/ref_micro.c): result arrays and RESULT fields per test"""
    exe = build_driver("ref_micro")
    res = {}
    for name, n, dt in REF_MICRO:
        outp = os.path.join(str(tmp_path), "%s_%d.bin" % (name, int(preload)))
        out, _ = run(exe, [name, n, outp], preload=preload, cwd=str(tmp_path), timeout=120)
        res[name] = (np.fromfile(outp, dtype=dt), fields([l for l in out.splitlines() if l.startswith("RESULT")][0]))
    return res


def test_ref_micro_tests_on_cpu_blas(tmp_path):
    """The reference's own micro-tests (tests/c/copy.c, dsdot.c, rot.c, gbmv.c, trmv.c, trsm.c, hemm.c) against the CPU BLAS:
    closed forms where the fill has one, and a numpy evaluation of the same fills for the matrix routines."""
    res = run_ref_micro(tmp_path, preload=False)
    assert float(res["copy"][1]["closed_form_err"]) == 0.0 and float(res["rot"][1]["closed_form_err"]) == 0.0
    assert float(res["gemm"][1]["closed_form_err"]) == 0.0      # the reference's own gemm.c in f32: k*r*c is exact below 2^24 (n = 256)
    assert float(res["dsdot"][1]["closed_form_err"]) <= 1e-7 * 9.0e9     # OpenBLAS sums blocks of its float kernel in float (5e-9 relative here);
    #                                                                    # netlib DSDOT -- and the GPU kernel -- are exact on this input
    n = 400      # gbmv: A read as column-major band storage AB[ku+i-j, j] with lda = n, as cblas_sgbmv(ColMajor) reads the reference's array
    raw = np.zeros((n, n), np.float32)
    for row in range(n):
        for col in range(max(0, row - 2), min(n, row + 3)):
            raw.ravel()[(2 - row + col) + row * n] = 1
    AB = raw.reshape((n, n)).T                      # AB[r, j] = raw[r + j*n]
    x = np.arange(n, dtype=np.float64)
    y = x.copy()
    for j in range(n):
        for i in range(max(0, j - 2), min(n, j + 3)):
            y[i] += AB[2 + i - j, j] * x[j]
    assert np.allclose(res["gbmv"][0], y, rtol=1e-5, atol=1e-3)
    A = np.zeros((n, n)); mod = (n * n) // 10
    for r in range(n):
        A[r, r:] = (r * n + np.arange(r, n)) % mod
    assert np.allclose(res["trmv"][0], A @ np.arange(n, dtype=np.float64), rtol=2e-5)
    m, nrhs = 300, 150
    A = np.zeros((m, m))
    for r in range(m):
        A[r, r:] = (r * m + np.arange(r, m)) % 10
    A += 10 * np.eye(m)
    B = ((np.arange(m)[:, None] * nrhs + np.arange(nrhs)[None, :]) % 10).astype(np.float64)
    assert np.allclose(res["trsm"][0].reshape((nrhs, m)).T, np.linalg.solve(A, B), rtol=1e-9, atol=1e-9)
    m = 200
    r, c = np.indices((m, m))
    H = np.where(c > r, c + 1j * r, np.where(c < r, r - 1j * c, c)).astype(np.complex128)
    Bm = ((r * m + c) % m + 1j * (r % 10)).astype(np.complex128)
    assert np.allclose(res["hemm"][0].reshape((m, m)).T, H @ Bm, rtol=1e-4, atol=1e-2)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gemm_fixture_under_preload(tmp_path):
    """BASELINE config 1: DGEMM 1024^3 through the interposer == the closed form, via dgemm_ and cblas_dgemm;
    operands come from the interposed calloc, so every operand is a tracker hit (no staging copies)."""
    exe = build_driver("gemm_fixture")
    out, err = run(exe, [1024, 11], preload=True, cwd=str(tmp_path))
    res = [fields(l) for l in out.splitlines() if l.startswith("RESULT")]
    assert len(res) == 2 and all(float(r["max_abs_err"]) == 0.0 for r in res), out
    st = fields([l for l in out.splitlines() if l.startswith("STATS")][0])
    assert int(st["misses"]) == 0 and int(st["hits"]) == 2 * 11 * 3 and int(st["h2d"]) == 0
    assert int(st["managed_allocs"]) >= 2 * 11 * 3
    cpu_out, _ = run(exe, [1024, 3])
    cpu = [fields(l) for l in cpu_out.splitlines() if l.startswith("RESULT")]
    assert cpu[0]["checksum"] == res[0]["checksum"]
    # reference blas2cuda.c:266-273: statistics.csv with the hit/miss counters is written at exit
    stats = open(os.path.join(str(tmp_path), "statistics.csv")).read().splitlines()
    assert stats[0].startswith("Hits, Misses") and stats[1].split(",")[1].strip() == "0"


@pytest.mark.gpu
def test_cg_chain_under_preload(tmp_path):
    """BASELINE config 3 (scaled for the test): chained Level-1/2 calls on tracked managed buffers agree with
    the CPU BLAS run of the same binary and never stage a copy."""
    exe = build_driver("cg_chain")
    # 2048 doubles = 16 KiB per vector: below the default 64 KiB rule, so lower it (north_star (3): size-based selector)
    gout, _ = run(exe, [2048, 25], preload=True, env_extra={"BLAS2CUDA_OPTIONS": "threshold=8192"}, cwd=str(tmp_path))
    cout, _ = run(exe, [2048, 25])
    gr = fields([l for l in gout.splitlines() if l.startswith("RESULT")][0])
    cr = fields([l for l in cout.splitlines() if l.startswith("RESULT")][0])
    assert float(gr["rnorm"]) < 1e-10 and float(cr["rnorm"]) < 1e-10
    assert abs(float(gr["xsum"]) - float(cr["xsum"])) <= 1e-12 * max(1.0, abs(float(cr["xsum"])))
    assert gr["imax"] == cr["imax"]
    st = fields([l for l in gout.splitlines() if l.startswith("STATS")][0])
    assert int(st["h2d"]) == 0 and int(st["d2h"]) == 0 and int(st["hits"]) > 0


@pytest.mark.gpu
def test_l1_chain_under_preload(tmp_path):
    """BASELINE config 3, Level-1 half (scaled): CPU-initialised calloc'd vectors, chained ddot/daxpy/dnrm2/idamax under
    LD_PRELOAD agree with the CPU BLAS run, nothing is staged, and the vectors end up resident on the device."""
    exe = build_driver("l1_chain")
    gout, _ = run(exe, [1 << 22, 6], preload=True, cwd=str(tmp_path))
    cout, _ = run(exe, [1 << 22, 6])
    gr = fields([l for l in gout.splitlines() if l.startswith("RESULT")][0])
    cr = fields([l for l in cout.splitlines() if l.startswith("RESULT")][0])
    assert gr["imax"] == cr["imax"] == str((1 << 22) // 3 + 1)          # first of the planted tie, 1-based
    assert abs(float(gr["acc"]) - float(cr["acc"])) <= 1e-9 * abs(float(cr["acc"])) + 1e-18
    assert abs(float(gr["nrm"]) - float(cr["nrm"])) <= 1e-12 * float(cr["nrm"])
    st = fields([l for l in gout.splitlines() if l.startswith("STATS")][0])
    assert int(st["h2d"]) == 0 and int(st["d2h"]) == 0 and int(st["misses"]) == 0 and int(st["prefetch"]) > 0
    res = [l for l in gout.splitlines() if l.startswith("RESIDENCY")][0].split()
    assert res[1] == "x=0" and res[2] == "y=0"


@pytest.mark.gpu
def test_allocs_under_preload(tmp_path):
    exe = build_driver("allocs")
    out, _ = run(exe, preload=True, cwd=str(tmp_path), timeout=90)
    r = fields([l for l in out.splitlines() if l.startswith("RESULT")][0])
    assert r["ok"] == "1" and int(r["tracked"]) > 300          # blocks >= 64 KiB are managed
    out, _ = run(exe, preload=True, env_extra={"BLAS2CUDA_OPTIONS": "heuristic=false"}, cwd=str(tmp_path), timeout=90)
    assert "RESULT ok=1 tracked=0" in out
    out, _ = run(exe, preload=True, env_extra={"BLAS2CUDA_OPTIONS": "heuristic=true"}, cwd=str(tmp_path), timeout=90)
    assert "RESULT ok=1 tracked=512" in out


@pytest.mark.gpu
def test_preload_untracked_operands_are_staged(tmp_path):
    """heuristic=false: operands stay on the glibc heap -> every operand is a miss, staged through the
    workspace and copied back (the reference's gpuptr miss path, runtime-mem.hpp:84-135); result identical."""
    exe = build_driver("gemm_fixture")
    out, _ = run(exe, [512, 2], preload=True, env_extra={"BLAS2CUDA_OPTIONS": "heuristic=false"}, cwd=str(tmp_path))
    res = [fields(l) for l in out.splitlines() if l.startswith("RESULT")]
    assert all(float(r["max_abs_err"]) == 0.0 for r in res)
    st = fields([l for l in out.splitlines() if l.startswith("STATS")][0])
    assert int(st["hits"]) == 0 and int(st["misses"]) == 2 * 2 * 3 and int(st["h2d"]) > 0 and int(st["d2h"]) > 0


def test_allocation_heuristics_and_oracle_file(tmp_path):
    """north_star (3): the size-based selector and the reference's other heuristics (obj_tracker.c:226-243), including the
    per-allocation oracle file of lib/oracle.c:26-72 ("H #<nth> ..." host, "D #<nth> ..." device lines).  CPU-only: asks
    the tracker what it WOULD decide, touches no device."""
    import ctypes
    import libgpublas_b200 as g
    lib = g.load()
    lib.b200blas_tracker_decision.argtypes = [ctypes.c_ulonglong, ctypes.c_size_t]
    try:
        lib.b200blas_set_options(b"heuristic=size;threshold=65536")
        assert [lib.b200blas_tracker_decision(7, s) for s in (1, 65535, 65536, 1 << 30)] == [0, 0, 1, 1]
        lib.b200blas_set_options(b"threshold=4096")
        assert lib.b200blas_tracker_decision(0, 4096) == 1 and lib.b200blas_tracker_decision(0, 4095) == 0
        lib.b200blas_set_options(b"heuristic=true")
        assert lib.b200blas_tracker_decision(3, 1) == 1
        lib.b200blas_set_options(b"heuristic=false")
        assert lib.b200blas_tracker_decision(3, 1 << 30) == 0
        trace = tmp_path / "objtrace.txt"
        trace.write_text("H #0 [0x1] fun=[malloc]\nD #1 [0x2] fun=[calloc]\nD #5 [0x3] fun=[malloc]\nH #6 x\ngarbage line\nD #70 y\n")
        lib.b200blas_set_options(("heuristic=oracle:%s" % trace).encode())
        got = [lib.b200blas_tracker_decision(n, 8) for n in (0, 1, 2, 5, 6, 69, 70, 71, 10 ** 6)]
        assert got == [0, 1, 0, 1, 0, 0, 1, 0, 0]
    finally:
        lib.b200blas_set_options(b"heuristic=size;threshold=65536")


@pytest.mark.gpu
def test_object_trace_feeds_the_oracle_heuristic(tmp_path):
    """BLAS2CUDA_OPTIONS=trace prints the reference's TRACE_OUTPUT lines (obj_tracker.c:426-483); turned into an oracle file
    (what scripts/analyze_trace.py does) they reproduce the placement under heuristic=oracle:<file>."""
    exe = build_driver("cg_chain")
    # 1024 doubles = 8 KiB per vector: above stdio's own 4 KiB buffer malloc, which must stay on the heap
    out, _ = run(exe, [1024, 3], preload=True, env_extra={"BLAS2CUDA_OPTIONS": "trace;threshold=8192"}, cwd=str(tmp_path))
    pat = TRACE_LINE
    ev = [pat.match(l).groups() for l in out.splitlines() if l[:2] in ("T ", "U ", "C ")]
    tracked = [e for e in ev if e[0] == "T"]
    assert len(tracked) == 5 and {e[3] for e in tracked} == {"calloc"}            # A, x, r, p, q
    assert len([e for e in ev if e[0] == "U"]) == 5
    calls = [e for e in ev if e[0] == "C"]
    assert {"dgemv_", "ddot_", "daxpy_", "dscal_"} <= {e[3] for e in calls}
    assert {e[2] for e in calls} <= {e[2] for e in tracked}                        # every C line names a tracked block
    oracle = tmp_path / "oracle.txt"
    big = max(int(e[4]) for e in tracked)
    oracle.write_text("".join("%s #%s\n" % ("D" if int(e[4]) == big else "H", e[1]) for e in tracked))
    out2, _ = run(exe, [1024, 3], preload=True, env_extra={"BLAS2CUDA_OPTIONS": "trace;heuristic=oracle:%s" % oracle}, cwd=str(tmp_path))
    t2 = [l for l in out2.splitlines() if l.startswith("T ")]
    assert len(t2) == 1 and ("reqsize=[%d]" % big) in t2[0]                        # only the matrix was placed on the device
    r1 = [l for l in out.splitlines() if l.startswith("RESULT")][0].split()[:6]
    r2 = [l for l in out2.splitlines() if l.startswith("RESULT")][0].split()[:6]
    assert r1 == r2                                                                # same numbers either way
