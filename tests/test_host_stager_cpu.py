"""The pageable-memory bounce ring (libgpublas_b200/csrc/host_stager.cu) on the CPU: the file is pure host code, compiled here with
g++ against a deferred-execution stand-in for the CUDA runtime calls it makes (tests/drivers/stager_mock.cpp), under
AddressSanitizer.  Piece arithmetic, slot rotation and slot re-use ordering for awkward shapes; the real transfers are what
tests/test_level12_gpu.py::test_pageable_operands_through_the_bounce_ring checks on a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "tests", "drivers")
BUILD = os.path.join(DRV, "_build")


def test_bounce_ring_round_trips_with_a_deferred_stream(tmp_path):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "stager_mock")
    src = os.path.join(DRV, "stager_mock.cpp")
    deps = [src, os.path.join(ROOT, "libgpublas_b200", "csrc", "host_stager.cu"), os.path.join(ROOT, "libgpublas_b200", "csrc", "runtime.h")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-x", "c++", "-fsanitize=address", "-fno-omit-frame-pointer", "-Wall",
                               "-I/usr/local/cuda/include", "-o", exe, src, "-lpthread"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "RESULT cases=9 bad=0" in out.stdout, out.stdout


def test_staging_pipelines_under_the_stream_simulator(tmp_path):
    """staged_gemm.cuh (gemm_pipelined, gemm_first_touch), staged_level3.cuh (syrk_pipelined, trxm_pipelined) and host_stager.cu
    compiled unchanged against the CUDA stream / event simulator (tests/drivers/simcuda.inc; see mgsim.cpp): the three streams of a
    call are ordered only by events, and the simulator executes the queued operations in adversarial orders (random, kernels first,
    copies first) with OpenBLAS as the kernels -- a chunk multiplied before its copy landed, a panel returned before its multiply or
    a bounce slot re-packed in flight gives a wrong result here.  65 cases: pinned / pageable / fresh managed operands, all routines."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import find_openblas
    import pytest
    ob = find_openblas()
    if ob is None:
        pytest.skip("no CPU BLAS in this image")
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "stagesim")
    csrc = os.path.join(ROOT, "libgpublas_b200", "csrc")
    srcs = [os.path.join(DRV, "stagesim.cpp"), os.path.join(csrc, "host_stager.cu")]
    deps = srcs + [os.path.join(DRV, "simcuda.inc")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".h", ".cuh"))]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-Wall", "-I/usr/local/cuda/include", "-o", exe, srcs[0], "-x", "c++", srcs[1], "-ldl", "-lpthread"])
    env = dict(os.environ, MGSIM_OPENBLAS=ob, OPENBLAS_CORETYPE="SkylakeX", OPENBLAS_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    env["LD_LIBRARY_PATH"] = os.path.dirname(ob) + ":" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([exe], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert "RESULT cases=65 failed=0" in out.stdout, out.stdout[-2000:]
    assert "DEADLOCK" not in out.stderr
    for needle in ("first-touch dgemm", "pipelined dsyrk", "pipelined dtrsm", "pipelined dtrmm", "operands=pageable", "operands=managed", "policy=1", "policy=2"):
        assert needle in out.stdout, needle
