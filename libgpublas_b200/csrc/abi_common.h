// abi_common.h -- helpers shared by the Fortran and CBLAS entry-point files.
#pragma once
#include "common.cuh"
#include "runtime.h"
#include "kernels.h"
#include "tracker.h"
#include <cuComplex.h>

namespace b200 {

// netlib LSAME on the first character (the reference calls the CPU BLAS's lsame_ once per test,
// runtime-blas.c:55-57)
static inline bool lsame(const char* p, char c) {
    char a = *p;
    if (a >= 'a' && a <= 'z') a -= 32;
    return a == c;
}
static inline int imax(int a, int b) { return a > b ? a : b; }

// Report an illegal argument the way the reference does (runtime-blas.c:34-52): SRNAME is the
// routine name upper-cased and blank-padded to 6 characters ("DGEMM "), passed with INFO to the
// xerbla_ found by normal symbol lookup, so an application- or tester-supplied XERBLA wins.
void call_xerbla(const char* routine, int info);

template <typename T> struct is_zero_t;
static inline bool is0(float a) { return a == 0.f; }
static inline bool is0(double a) { return a == 0.0; }
static inline bool is0(cuFloatComplex a) { return a.x == 0.f && a.y == 0.f; }
static inline bool is0(cuDoubleComplex a) { return a.x == 0.0 && a.y == 0.0; }
static inline bool is1(float a) { return a == 1.f; }
static inline bool is1(double a) { return a == 1.0; }
static inline bool is1(cuFloatComplex a) { return a.x == 1.f && a.y == 0.f; }
static inline bool is1(cuDoubleComplex a) { return a.x == 1.0 && a.y == 0.0; }

// RAII bracket of one interposed call: tracker off for this thread (call_kernel's
// obj_tracker_internal_enter/leave), workspace reset on entry, synchronous return on exit.
struct CallScope {
    TrackerGuard guard;
    CallScope() { ensure_init(); ws_reset(); t_call_name = "blas"; }
    explicit CallScope(const char* routine) { ensure_init(); ws_reset(); t_call_name = routine; }
    ~CallScope() { finish_call(); }
};

}  // namespace b200
