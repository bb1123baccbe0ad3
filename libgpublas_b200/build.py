"""Build libb200blas.so in-tree with nvcc for sm_100a (no torch extension machinery: the product
is a plain C-ABI shared object).  Used by __graft_entry__.build() and `python -m libgpublas_b200.build`."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200blas.so")
OBJDIR = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


INTERPOSED = ["malloc", "calloc", "realloc", "free", "memalign", "posix_memalign", "aligned_alloc", "valloc", "malloc_usable_size",
              "pthread_create"]          # csrc/tracker.cpp


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def headers_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(verbose=False, force=False):
    os.makedirs(OBJDIR, exist_ok=True)
    hm = headers_mtime()
    objs, procs = [], []
    for src in sources():
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        objs.append(op)
        if not force and os.path.exists(op) and os.path.getmtime(op) > max(os.path.getmtime(sp), hm):
            continue
        if src.endswith(".cpp"):   # host-only units (the malloc interposer cannot be a .cu)
            cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-Wall", "-I/usr/local/cuda/include", "-c", sp, "-o", op]
        else:
            cmd = [NVCC] + ARCH + FLAGS + ["-c", sp, "-o", op]
        log = open(op + ".log", "w")
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=log, stderr=subprocess.STDOUT), log, op))
    failed = False
    for src, cmd, p, log, op in procs:
        rc = p.wait()
        log.close()
        text = open(op + ".log").read()
        if rc != 0:
            failed = True
            sys.stderr.write("FAILED: %s\n%s\n" % (" ".join(cmd), text))
        elif verbose:
            sys.stderr.write(text)
    if failed:
        raise RuntimeError("nvcc failed")
    if procs or not os.path.exists(OUT) or force:
        # exactly the C ABI of include/b200blas.h plus the allocator interposers leave the library: weak libstdc++ template
        # instantiations (std::thread, std::vector) must not be visible to -- or interposable by -- the program it is preloaded into
        hdr = open(os.path.join(os.path.dirname(HERE), "include", "b200blas.h")).read()
        names = sorted(set(re.findall(r"^B200_API [^;(]*?(\w+)\(", hdr, re.M)) | set(INTERPOSED))
        vs = os.path.join(OBJDIR, "exports.map")
        with open(vs, "w") as f:
            f.write("{\n  global:\n" + "".join("    %s;\n" % n for n in names) + "  local: *;\n};\n")
        # -e: running the .so prints the option help (reference entry.c / meson.build:25)
        link = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-Xlinker", "--no-undefined", "-Xlinker", "-e,b200blas_entry", "-Xlinker", "--version-script=" + vs,
                                                                "-ldl", "-lpthread", "-cudart", "static"]
        subprocess.check_call(link)
    return OUT


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
