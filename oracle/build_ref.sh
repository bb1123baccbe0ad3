#!/bin/sh
# Compiles the host-side pieces of the REAL reference that are on this path and build from their own sources, where they
# lie under $1 (default /root/reference), into oracle/_ref/libref_host.so -- nothing is copied into this repository:
#   lib/oracle.c      oracle_load_file / oracle_should_alloc_managed_ptr  (the "H #nth" / "D #nth" placement file)
#   runtime-blas.c    func_name_to_f77                                    ("dgemm_" -> "DGEMM " for XERBLA)
# and its stand-alone tracer into oracle/_ref/libobjtracker.so (see below).
# The arithmetic of the path is NOT in the reference (it forwards to cuBLAS / the CPU BLAS), and its
# blas_level3/*.cc need cuBLAS at run time, so there is nothing numerical to build.  tests/test_oracle.py uses this
# library, when present, to check the restatements in libb200blas.so (tracker.cpp) and oracle/refblas.c against the
# reference's own code.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/_ref"
CC=${CC:-gcc}
$CC -O1 -fPIC -shared -std=gnu11 -w -DUSE_CUDA=1 -I/usr/local/cuda/include -I"$REF" -I"$REF/lib" \
    -o "$HERE/_ref/libref_host.so" "$REF/lib/oracle.c" "$REF/runtime-blas.c" "$HERE/ref_shims.c" -ldl \
    -Wl,--allow-shlib-undefined -Wl,-z,lazy
echo "built $HERE/_ref/libref_host.so"
# The reference's stand-alone object tracer (lib/meson.build:22-35: blas_tracker.c + obj_tracker.c with -DSTANDALONE
# -DTRACE_OUTPUT): CPU-only, LD_PRELOADed into a program it prints the T/U/C trace lines the oracle heuristic and
# scripts/analyze_trace.py consume.  tests/test_preload.py runs it to pin the trace-line FORMAT libb200blas.so prints
# under BLAS2CUDA_OPTIONS=trace against the real thing.
$CC -O1 -fPIC -shared -std=gnu11 -w -DSTANDALONE -DTRACE_OUTPUT -I"$REF" -I"$REF/lib" \
    -o "$HERE/_ref/libobjtracker.so" "$REF/lib/blas_tracker.c" "$REF/lib/obj_tracker.c" -ldl -lpthread \
    -Wl,-init,obj_tracker_init,-fini,obj_tracker_fini
echo "built $HERE/_ref/libobjtracker.so"
