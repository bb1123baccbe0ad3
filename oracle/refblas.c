/* oracle/refblas.c -- the CPU parity oracle.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The product library (libb200blas.so) never
 * links, dlopens or calls it and has no CPU fallback of any kind.
 *
 * What it restates.  The reference (Prince781/libgpublas) contains no arithmetic of its
 * own on the BLAS hot path: each interposed routine validates its arguments on the host
 * (blas_level3/gemm.cc:87-127, syrk.cc:78-146, trsm.cc:76-132, trmm.cc:81-137, ...) and
 * then forwards to a third-party BLAS that is NOT in /root/reference and NOT version
 * pinned by it (meson.build:100-114 `find_library('blas')`, `find_library('cublas')`):
 * cuBLAS on the GPU (gemm.cc:63-70 ...) or the CPU `libblas` via dlsym(RTLD_NEXT)
 * (runtime-blas.c:59-69).  Both implement the published netlib reference BLAS
 * semantics, so this oracle restates (a) the reference's host-side argument checks and
 * quick returns, routine by routine, and (b) the netlib definition of each routine as
 * plain loops.
 *
 * How it is pinned (tests/test_oracle.py, CPU-only):
 *   1. the reference's own closed-form fixture tests/c/gemm.c:29-35 (A[i,j]=i, B[i,j]=j,
 *      alpha=1, beta=0  =>  C[i,j] = k*i*j exactly in f64) -- a golden vector needing no
 *      library; committed as tests/golden/ files;
 *   2. the netlib ?blat3 suite the reference runs (tests/netlib/dblat3.f, zblat3.f with
 *      input.dblat3 / input.zblat3) restated in tests/blat3.py: DBEG/ZBEG generator,
 *      DMAKE rogue padding, DMMCH ratio test < 16, LDERES untouched-padding test and the
 *      DCHKE error-exit INFO table;
 *   3. the CPU BLAS the reference's interposer sits in front of in this image,
 *      OpenBLAS 0.3.15 (opencv_python_headless.libs/libopenblasp-r0-*.so), on seeded
 *      random inputs for every routine, including Level 1/2 where the reference ships
 *      no tester ("Level-1/Level-2 parity unpinned by the reference";
 *      pinned here by OpenBLAS only).
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>

static int imax(int a, int b) { return a > b ? a : b; }
/* netlib LSAME restated (reference runtime-blas.c:55-57 calls the CPU BLAS's lsame_) */
static int ref_lsame(char a, char b)
{
    if (a >= 'a' && a <= 'z') a -= 32;
    if (b >= 'a' && b <= 'z') b -= 32;
    return a == b;
}

#define T double
#define R(name) ref_d##name
#define RI(name) ref_id##name
#include "refblas_real.inc"
#include "refblas_l2x.inc"
#undef T
#undef R
#undef RI

#define T float
#define R(name) ref_s##name
#define RI(name) ref_is##name
#include "refblas_real.inc"
#include "refblas_l2x.inc"
#undef T
#undef R
#undef RI

#define T double complex
#define RT double
#define R(name) ref_z##name
#define RI(name) ref_iz##name
#include "refblas_cplx.inc"
#define L2X_COMPLEX 1
#include "refblas_l2x.inc"
#undef L2X_COMPLEX
#undef T
#undef RT
#undef R
#undef RI

#define T float complex
#define RT float
#define R(name) ref_c##name
#define RI(name) ref_ic##name
#include "refblas_cplx.inc"
#define L2X_COMPLEX 1
#include "refblas_l2x.inc"
#undef L2X_COMPLEX
#undef T
#undef RT
#undef R
#undef RI

/* netlib DSDOT / SDSDOT (reference cblas.h declares cblas_dsdot / cblas_sdsdot): float operands, double accumulation */
double ref_dsdot(int n, const float *x, int incx, const float *y, int incy)
{
    double s = 0;
    for (int i = 0; i < n; i++) {
        size_t ix = incx > 0 ? (size_t)i * incx : (size_t)(n - 1 - i) * (-incx), iy = incy > 0 ? (size_t)i * incy : (size_t)(n - 1 - i) * (-incy);
        s += (double)x[ix] * (double)y[iy];
    }
    return s;
}
float ref_sdsdot(int n, float sb, const float *x, int incx, const float *y, int incy) { return (float)((double)sb + ref_dsdot(n, x, incx, y, incy)); }

/* reference runtime-blas.c:38-52 func_name_to_f77: "dgemm_" -> "DGEMM " (upper-case,
 * '_' -> ' '); out must hold strlen(name)+1 bytes. */
void ref_func_name_to_f77(const char *name, char *out)
{
    for (; *name; name++, out++) {
        char ch = *name;
        if (ch == '_') ch = ' ';
        else if (ch >= 'a' && ch <= 'z') ch -= 32;
        *out = ch;
    }
    *out = 0;
}

/* Blocked right-looking lower Cholesky used by the C4 workload checker (netlib DPOTRF2-style
 * unblocked kernel); returns 0 or the 1-based index of the first non-positive pivot. */
int ref_dpotrf_lower(int n, double *a, int lda)
{
    for (int j = 0; j < n; j++) {
        double d = a[(size_t)j + (size_t)j * lda];
        for (int l = 0; l < j; l++) d -= a[(size_t)j + (size_t)l * lda] * a[(size_t)j + (size_t)l * lda];
        if (!(d > 0)) return j + 1;
        d = sqrt(d);
        a[(size_t)j + (size_t)j * lda] = d;
        for (int i = j + 1; i < n; i++) {
            double s = a[(size_t)i + (size_t)j * lda];
            for (int l = 0; l < j; l++) s -= a[(size_t)i + (size_t)l * lda] * a[(size_t)j + (size_t)l * lda];
            a[(size_t)i + (size_t)j * lda] = s / d;
        }
    }
    return 0;
}
