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
