/* b200blas.h -- the C ABI of libb200blas.so, a B200-native drop-in for the BLAS hot path of
 * Prince781/libgpublas ("blas2cuda").  Plain C: pointers and sizes only.
 *
 * The library is meant to be LD_PRELOADed (reference scripts/blas2cuda.sh:27, tests/netlib/
 * test.py:28) or linked ahead of the CPU BLAS.  It exports exactly the symbols the reference's
 * interposer exports for this path, with the CPU BLAS's own (gfortran) calling convention:
 *
 *   1. Fortran BLAS symbols  (reference blas.h:202-314, blas_level3/[star].cc F77_xxx wrappers;
 *      Level 1/2 -- every routine blas_level1/[star].cc and blas_level2/[star].cc name -- follow the CPU BLAS
 *      ABI because blas.h:11-198 is unreliable)
 *   2. CBLAS symbols         (reference cblas.h:46-824, both layouts; complex scalars by pointer as in
 *      standard CBLAS, not by value as cblas.h:662-677 mis-declares)
 *   3. allocator symbols     (reference lib/obj_tracker.c:789,842,902,948, plus the aligned allocators)
 *   4. a small control API   (b200blas_[star]) for embedding, tests and benchmarks.
 * COVERAGE.md lists the 300 BLAS entry points by level, precision and symbol family.
 *
 * Operands may live in: tracked managed memory (from the interposed malloc/calloc), any other
 * CUDA managed or device memory (used in place), or ordinary host memory (staged).  Every entry
 * point returns only after results are visible to the CPU (unless b200blas_set_sync(0)).
 * Illegal arguments call xerbla_(SRNAME,INFO) with netlib numbering and return; device failures
 * print to fd 2 and abort() (reference runtime.c:208-213).  There is no CPU fallback.
 */
#ifndef B200BLAS_H
#define B200BLAS_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float re, im; } b200_c32;     /* == float _Complex, interleaved (re,im) */
typedef struct { double re, im; } b200_c64;    /* == double _Complex */

/* ------------------------------ 1. Fortran BLAS ABI ------------------------------ */
/* Level 3 -- reference blas.h:202-215 (gemm), :268-274 (syrk), :290-301 (trmm), :303-314 (trsm),
 * :254-266 (symm), :276-288 (syr2k), :217-227 (hemm), :230-239 (herk), :242-252 (her2k) */
B200_API void sgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const float* alpha, const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
B200_API void dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
B200_API void cgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* b, const int* ldb, const b200_c32* beta, b200_c32* c, const int* ldc);
B200_API void zgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* b, const int* ldb, const b200_c64* beta, b200_c64* c, const int* ldc);

B200_API void ssyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda, const float* beta, float* c, const int* ldc);
B200_API void dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* beta, double* c, const int* ldc);
B200_API void csyrk_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* beta, b200_c32* c, const int* ldc);
B200_API void zsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* beta, b200_c64* c, const int* ldc);
B200_API void strsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const float* alpha, const float* a, const int* lda, float* b, const int* ldb);
B200_API void strmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const float* alpha, const float* a, const int* lda, float* b, const int* ldb);
B200_API void dtrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const double* alpha, const double* a, const int* lda, double* b, const int* ldb);
B200_API void dtrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const double* alpha, const double* a, const int* lda, double* b, const int* ldb);
B200_API void ctrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const b200_c32* alpha, const b200_c32* a, const int* lda, b200_c32* b, const int* ldb);
B200_API void ctrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const b200_c32* alpha, const b200_c32* a, const int* lda, b200_c32* b, const int* ldb);
B200_API void ztrsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const b200_c64* alpha, const b200_c64* a, const int* lda, b200_c64* b, const int* ldb);
B200_API void ztrmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, const b200_c64* alpha, const b200_c64* a, const int* lda, b200_c64* b, const int* ldb);

/* the rest of the reference's Level-3 family -- blas.h:254-266 (symm), :217-227 (hemm), :276-288 (syr2k), :230-239 (herk: real
 * alpha, beta), :242-252 (her2k: complex alpha, real beta) */
B200_API void ssymm_(const char* side, const char* uplo, const int* m, const int* n, const float* alpha, const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
B200_API void dsymm_(const char* side, const char* uplo, const int* m, const int* n, const double* alpha, const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
B200_API void csymm_(const char* side, const char* uplo, const int* m, const int* n, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* b, const int* ldb, const b200_c32* beta, b200_c32* c, const int* ldc);
B200_API void zsymm_(const char* side, const char* uplo, const int* m, const int* n, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* b, const int* ldb, const b200_c64* beta, b200_c64* c, const int* ldc);
B200_API void chemm_(const char* side, const char* uplo, const int* m, const int* n, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* b, const int* ldb, const b200_c32* beta, b200_c32* c, const int* ldc);
B200_API void zhemm_(const char* side, const char* uplo, const int* m, const int* n, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* b, const int* ldb, const b200_c64* beta, b200_c64* c, const int* ldc);
B200_API void ssyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c, const int* ldc);
B200_API void dsyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc);
B200_API void csyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* b, const int* ldb, const b200_c32* beta, b200_c32* c, const int* ldc);
B200_API void zsyr2k_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* b, const int* ldb, const b200_c64* beta, b200_c64* c, const int* ldc);
B200_API void cherk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const b200_c32* a, const int* lda, const float* beta, b200_c32* c, const int* ldc);
B200_API void zherk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const b200_c64* a, const int* lda, const double* beta, b200_c64* c, const int* ldc);
B200_API void cher2k_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* b, const int* ldb, const float* beta, b200_c32* c, const int* ldc);
B200_API void zher2k_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* b, const int* ldb, const double* beta, b200_c64* c, const int* ldc);

/* Level 1 -- gfortran ABI of the CPU BLAS (functions return their result; complex results by value):
 * reference blas_level1/dot.cc:38-48, dotc.cc, dotu.cc, nrm2.cc:31-54, asum.cc, amax.cc:32-56, axpy.cc:44-59,
 * scal.cc, copy.cc, swap.cc (dead wrappers naming the routines).  i?amax_ is 1-based, 0 if n<1 or incx<=0.
 * Level 2 -- reference blas_level2/gemv.cc:11-118, trsv.cc:11-114. */
B200_API float sdot_(const int* n, const float* x, const int* incx, const float* y, const int* incy);
B200_API double ddot_(const int* n, const double* x, const int* incx, const double* y, const int* incy);
B200_API b200_c32 cdotu_(const int* n, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy);
B200_API b200_c32 cdotc_(const int* n, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy);
B200_API b200_c64 zdotu_(const int* n, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy);
B200_API b200_c64 zdotc_(const int* n, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy);
B200_API float snrm2_(const int* n, const float* x, const int* incx);
B200_API double dnrm2_(const int* n, const double* x, const int* incx);
B200_API float scnrm2_(const int* n, const b200_c32* x, const int* incx);
B200_API double dznrm2_(const int* n, const b200_c64* x, const int* incx);
B200_API float sasum_(const int* n, const float* x, const int* incx);
B200_API double dasum_(const int* n, const double* x, const int* incx);
B200_API float scasum_(const int* n, const b200_c32* x, const int* incx);
B200_API double dzasum_(const int* n, const b200_c64* x, const int* incx);
B200_API int isamax_(const int* n, const float* x, const int* incx);
B200_API int idamax_(const int* n, const double* x, const int* incx);
B200_API int icamax_(const int* n, const b200_c32* x, const int* incx);
B200_API int izamax_(const int* n, const b200_c64* x, const int* incx);
B200_API void saxpy_(const int* n, const float* alpha, const float* x, const int* incx, float* y, const int* incy);
B200_API void daxpy_(const int* n, const double* alpha, const double* x, const int* incx, double* y, const int* incy);
B200_API void caxpy_(const int* n, const b200_c32* alpha, const b200_c32* x, const int* incx, b200_c32* y, const int* incy);
B200_API void zaxpy_(const int* n, const b200_c64* alpha, const b200_c64* x, const int* incx, b200_c64* y, const int* incy);
B200_API void sscal_(const int* n, const float* alpha, float* x, const int* incx);
B200_API void dscal_(const int* n, const double* alpha, double* x, const int* incx);
B200_API void cscal_(const int* n, const b200_c32* alpha, b200_c32* x, const int* incx);
B200_API void zscal_(const int* n, const b200_c64* alpha, b200_c64* x, const int* incx);
B200_API void csscal_(const int* n, const float* alpha, b200_c32* x, const int* incx);
B200_API void zdscal_(const int* n, const double* alpha, b200_c64* x, const int* incx);
B200_API void scopy_(const int* n, const float* x, const int* incx, float* y, const int* incy);
B200_API void dcopy_(const int* n, const double* x, const int* incx, double* y, const int* incy);
B200_API void ccopy_(const int* n, const b200_c32* x, const int* incx, b200_c32* y, const int* incy);
B200_API void zcopy_(const int* n, const b200_c64* x, const int* incx, b200_c64* y, const int* incy);
B200_API void sswap_(const int* n, float* x, const int* incx, float* y, const int* incy);
B200_API void dswap_(const int* n, double* x, const int* incx, double* y, const int* incy);
B200_API void cswap_(const int* n, b200_c32* x, const int* incx, b200_c32* y, const int* incy);
B200_API void zswap_(const int* n, b200_c64* x, const int* incx, b200_c64* y, const int* incy);
B200_API void sgemv_(const char* trans, const int* m, const int* n, const float* alpha, const float* a, const int* lda, const float* x, const int* incx, const float* beta, float* y, const int* incy);
B200_API void dgemv_(const char* trans, const int* m, const int* n, const double* alpha, const double* a, const int* lda, const double* x, const int* incx, const double* beta, double* y, const int* incy);
B200_API void cgemv_(const char* trans, const int* m, const int* n, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* x, const int* incx, const b200_c32* beta, b200_c32* y, const int* incy);
B200_API void zgemv_(const char* trans, const int* m, const int* n, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* x, const int* incx, const b200_c64* beta, b200_c64* y, const int* incy);
B200_API void strsv_(const char* uplo, const char* trans, const char* diag, const int* n, const float* a, const int* lda, float* x, const int* incx);
B200_API void dtrsv_(const char* uplo, const char* trans, const char* diag, const int* n, const double* a, const int* lda, double* x, const int* incx);
B200_API void ctrsv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c32* a, const int* lda, b200_c32* x, const int* incx);
B200_API void ztrsv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c64* a, const int* lda, b200_c64* x, const int* incx);

/* more Level 1/2 (reference blas_level2/ger.cc, syr.cc, symv.cc, trmv.cc, blas_level1/rot.cc, rotg.cc): real types.
 * The Fortran declarations must precede the CBLAS enums they do not use; the cblas_ ones follow the enums below. */
B200_API void sger_(const int* m, const int* n, const float* alpha, const float* x, const int* incx, const float* y, const int* incy, float* a, const int* lda);
B200_API void ssyr_(const char* uplo, const int* n, const float* alpha, const float* x, const int* incx, float* a, const int* lda);
B200_API void ssymv_(const char* uplo, const int* n, const float* alpha, const float* a, const int* lda, const float* x, const int* incx, const float* beta, float* y, const int* incy);
B200_API void strmv_(const char* uplo, const char* trans, const char* diag, const int* n, const float* a, const int* lda, float* x, const int* incx);
B200_API void srot_(const int* n, float* x, const int* incx, float* y, const int* incy, const float* c, const float* s);
B200_API void srotg_(float* a, float* b, float* c, float* s);
B200_API void dger_(const int* m, const int* n, const double* alpha, const double* x, const int* incx, const double* y, const int* incy, double* a, const int* lda);
B200_API void dsyr_(const char* uplo, const int* n, const double* alpha, const double* x, const int* incx, double* a, const int* lda);
B200_API void dsymv_(const char* uplo, const int* n, const double* alpha, const double* a, const int* lda, const double* x, const int* incx, const double* beta, double* y, const int* incy);
B200_API void dtrmv_(const char* uplo, const char* trans, const char* diag, const int* n, const double* a, const int* lda, double* x, const int* incx);
B200_API void drot_(const int* n, double* x, const int* incx, double* y, const int* incy, const double* c, const double* s);
B200_API void drotg_(double* a, double* b, double* c, double* s);

/* banded, packed, Hermitian and complex Level 2 (reference blas_level2/gbmv.cc, bmv.cc (sbmv/hbmv), pmv.cc (spmv/hpmv),
 * hemv.cc, her.cc, her2.cc, hpr.cc, hpr2.cc, spr.cc, spr2.cc, syr2.cc, tbmv.cc, tbsv.cc, tpmv.cc, tpsv.cc, ger.cc (geru/gerc), trmv.cc):
 * netlib signatures and INFO numbering; band storage AB(ku+1+i-j, j), packed storage by columns. */
B200_API void sgbmv_(const char* trans, const int* m, const int* n, const int* kl, const int* ku, const float* alpha, const float* a, const int* lda, const float* x, const int* incx, const float* beta, float* y, const int* incy);
B200_API void stbmv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const float* a, const int* lda, float* x, const int* incx);
B200_API void stbsv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const float* a, const int* lda, float* x, const int* incx);
B200_API void stpmv_(const char* uplo, const char* trans, const char* diag, const int* n, const float* ap, float* x, const int* incx);
B200_API void stpsv_(const char* uplo, const char* trans, const char* diag, const int* n, const float* ap, float* x, const int* incx);
B200_API void ssbmv_(const char* uplo, const int* n, const int* k, const float* alpha, const float* a, const int* lda, const float* x, const int* incx, const float* beta, float* y, const int* incy);
B200_API void sspmv_(const char* uplo, const int* n, const float* alpha, const float* ap, const float* x, const int* incx, const float* beta, float* y, const int* incy);
B200_API void ssyr2_(const char* uplo, const int* n, const float* alpha, const float* x, const int* incx, const float* y, const int* incy, float* a, const int* lda);
B200_API void sspr_(const char* uplo, const int* n, const float* alpha, const float* x, const int* incx, float* ap);
B200_API void sspr2_(const char* uplo, const int* n, const float* alpha, const float* x, const int* incx, const float* y, const int* incy, float* ap);
B200_API void dgbmv_(const char* trans, const int* m, const int* n, const int* kl, const int* ku, const double* alpha, const double* a, const int* lda, const double* x, const int* incx, const double* beta, double* y, const int* incy);
B200_API void dtbmv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const double* a, const int* lda, double* x, const int* incx);
B200_API void dtbsv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const double* a, const int* lda, double* x, const int* incx);
B200_API void dtpmv_(const char* uplo, const char* trans, const char* diag, const int* n, const double* ap, double* x, const int* incx);
B200_API void dtpsv_(const char* uplo, const char* trans, const char* diag, const int* n, const double* ap, double* x, const int* incx);
B200_API void dsbmv_(const char* uplo, const int* n, const int* k, const double* alpha, const double* a, const int* lda, const double* x, const int* incx, const double* beta, double* y, const int* incy);
B200_API void dspmv_(const char* uplo, const int* n, const double* alpha, const double* ap, const double* x, const int* incx, const double* beta, double* y, const int* incy);
B200_API void dsyr2_(const char* uplo, const int* n, const double* alpha, const double* x, const int* incx, const double* y, const int* incy, double* a, const int* lda);
B200_API void dspr_(const char* uplo, const int* n, const double* alpha, const double* x, const int* incx, double* ap);
B200_API void dspr2_(const char* uplo, const int* n, const double* alpha, const double* x, const int* incx, const double* y, const int* incy, double* ap);
B200_API void cgbmv_(const char* trans, const int* m, const int* n, const int* kl, const int* ku, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* x, const int* incx, const b200_c32* beta, b200_c32* y, const int* incy);
B200_API void ctbmv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const b200_c32* a, const int* lda, b200_c32* x, const int* incx);
B200_API void ctbsv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const b200_c32* a, const int* lda, b200_c32* x, const int* incx);
B200_API void ctpmv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c32* ap, b200_c32* x, const int* incx);
B200_API void ctpsv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c32* ap, b200_c32* x, const int* incx);
B200_API void chbmv_(const char* uplo, const int* n, const int* k, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* x, const int* incx, const b200_c32* beta, b200_c32* y, const int* incy);
B200_API void chpmv_(const char* uplo, const int* n, const b200_c32* alpha, const b200_c32* ap, const b200_c32* x, const int* incx, const b200_c32* beta, b200_c32* y, const int* incy);
B200_API void chemv_(const char* uplo, const int* n, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* x, const int* incx, const b200_c32* beta, b200_c32* y, const int* incy);
B200_API void cgeru_(const int* m, const int* n, const b200_c32* alpha, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy, b200_c32* a, const int* lda);
B200_API void cgerc_(const int* m, const int* n, const b200_c32* alpha, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy, b200_c32* a, const int* lda);
B200_API void cher_(const char* uplo, const int* n, const float* alpha, const b200_c32* x, const int* incx, b200_c32* a, const int* lda);
B200_API void cher2_(const char* uplo, const int* n, const b200_c32* alpha, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy, b200_c32* a, const int* lda);
B200_API void chpr_(const char* uplo, const int* n, const float* alpha, const b200_c32* x, const int* incx, b200_c32* ap);
B200_API void chpr2_(const char* uplo, const int* n, const b200_c32* alpha, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy, b200_c32* ap);
B200_API void ctrmv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c32* a, const int* lda, b200_c32* x, const int* incx);
B200_API void zgbmv_(const char* trans, const int* m, const int* n, const int* kl, const int* ku, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* x, const int* incx, const b200_c64* beta, b200_c64* y, const int* incy);
B200_API void ztbmv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const b200_c64* a, const int* lda, b200_c64* x, const int* incx);
B200_API void ztbsv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const b200_c64* a, const int* lda, b200_c64* x, const int* incx);
B200_API void ztpmv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c64* ap, b200_c64* x, const int* incx);
B200_API void ztpsv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c64* ap, b200_c64* x, const int* incx);
B200_API void zhbmv_(const char* uplo, const int* n, const int* k, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* x, const int* incx, const b200_c64* beta, b200_c64* y, const int* incy);
B200_API void zhpmv_(const char* uplo, const int* n, const b200_c64* alpha, const b200_c64* ap, const b200_c64* x, const int* incx, const b200_c64* beta, b200_c64* y, const int* incy);
B200_API void zhemv_(const char* uplo, const int* n, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* x, const int* incx, const b200_c64* beta, b200_c64* y, const int* incy);
B200_API void zgeru_(const int* m, const int* n, const b200_c64* alpha, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy, b200_c64* a, const int* lda);
B200_API void zgerc_(const int* m, const int* n, const b200_c64* alpha, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy, b200_c64* a, const int* lda);
B200_API void zher_(const char* uplo, const int* n, const double* alpha, const b200_c64* x, const int* incx, b200_c64* a, const int* lda);
B200_API void zher2_(const char* uplo, const int* n, const b200_c64* alpha, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy, b200_c64* a, const int* lda);
B200_API void zhpr_(const char* uplo, const int* n, const double* alpha, const b200_c64* x, const int* incx, b200_c64* ap);
B200_API void zhpr2_(const char* uplo, const int* n, const b200_c64* alpha, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy, b200_c64* ap);
B200_API void ztrmv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c64* a, const int* lda, b200_c64* x, const int* incx);

/* more Level 1 (reference blas_level1/rotm.cc, rotmg.cc, amin.cc; cblas.h dsdot / sdsdot; netlib CSROT / ZDROT).  i?amin_ is not
 * netlib; the reference's amin.cc names it and the CPU BLAS (OpenBLAS) exports it: 1-based index of the first smallest |x|. */
B200_API void srotm_(const int* n, float* x, const int* incx, float* y, const int* incy, const float* param);
B200_API void drotm_(const int* n, double* x, const int* incx, double* y, const int* incy, const double* param);
B200_API void srotmg_(float* d1, float* d2, float* x1, const float* y1, float* param);
B200_API void drotmg_(double* d1, double* d2, double* x1, const double* y1, double* param);
B200_API void csrot_(const int* n, b200_c32* x, const int* incx, b200_c32* y, const int* incy, const float* c, const float* s);
B200_API void zdrot_(const int* n, b200_c64* x, const int* incx, b200_c64* y, const int* incy, const double* c, const double* s);
B200_API void crotg_(b200_c32* ca, const b200_c32* cb, float* c, b200_c32* s);
B200_API void zrotg_(b200_c64* ca, const b200_c64* cb, double* c, b200_c64* s);
B200_API int isamin_(const int* n, const float* x, const int* incx);
B200_API int idamin_(const int* n, const double* x, const int* incx);
B200_API int icamin_(const int* n, const b200_c32* x, const int* incx);
B200_API int izamin_(const int* n, const b200_c64* x, const int* incx);
B200_API double dsdot_(const int* n, const float* x, const int* incx, const float* y, const int* incy);
B200_API float sdsdot_(const int* n, const float* sb, const float* x, const int* incx, const float* y, const int* incy);

/* ------------------------------ 2. CBLAS ABI ------------------------------ */
/* enums: reference cblas.h:21-25 */
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
typedef enum CBLAS_ORDER CBLAS_LAYOUT;
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
typedef size_t CBLAS_INDEX;

B200_API void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, int k, float alpha, const float* a, int lda, const float* b, int ldb, float beta, float* c, int ldc);
B200_API void cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, int k, double alpha, const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc);
B200_API void cblas_cgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, int k, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);
B200_API void cblas_zgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, int k, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);

/* CBLAS Level 3 beyond gemm -- reference cblas.h:693-824 */
B200_API void cblas_ssyrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, float alpha, const float* a, int lda, float beta, float* c, int ldc);
B200_API void cblas_ssyr2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, float alpha, const float* a, int lda, const float* b, int ldb, float beta, float* c, int ldc);
B200_API void cblas_ssymm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, float alpha, const float* a, int lda, const float* b, int ldb, float beta, float* c, int ldc);
B200_API void cblas_strsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, float alpha, const float* a, int lda, float* b, int ldb);
B200_API void cblas_strmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, float alpha, const float* a, int lda, float* b, int ldb);
B200_API void cblas_dsyrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, double alpha, const double* a, int lda, double beta, double* c, int ldc);
B200_API void cblas_dsyr2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, double alpha, const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc);
B200_API void cblas_dsymm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, double alpha, const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc);
B200_API void cblas_dtrsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, double alpha, const double* a, int lda, double* b, int ldb);
B200_API void cblas_dtrmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, double alpha, const double* a, int lda, double* b, int ldb);
B200_API void cblas_csyrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, const void* a, int lda, const void* beta, void* c, int ldc);
B200_API void cblas_ctrsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, const void* alpha, const void* a, int lda, void* b, int ldb);
B200_API void cblas_ctrmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, const void* alpha, const void* a, int lda, void* b, int ldb);
B200_API void cblas_zsyrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, const void* a, int lda, const void* beta, void* c, int ldc);
B200_API void cblas_ztrsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, const void* alpha, const void* a, int lda, void* b, int ldb);
B200_API void cblas_ztrmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa, enum CBLAS_DIAG diag, int m, int n, const void* alpha, const void* a, int lda, void* b, int ldb);

/* CBLAS Level 1 / 2 -- reference cblas.h:46-656; cblas_i?amax is 0-based */
B200_API float cblas_sdot(int n, const float* x, int incx, const float* y, int incy);
B200_API double cblas_ddot(int n, const double* x, int incx, const double* y, int incy);
B200_API void cblas_cdotu_sub(int n, const void* x, int incx, const void* y, int incy, void* out);
B200_API void cblas_cdotc_sub(int n, const void* x, int incx, const void* y, int incy, void* out);
B200_API void cblas_zdotu_sub(int n, const void* x, int incx, const void* y, int incy, void* out);
B200_API void cblas_zdotc_sub(int n, const void* x, int incx, const void* y, int incy, void* out);
B200_API float cblas_snrm2(int n, const float* x, int incx);
B200_API double cblas_dnrm2(int n, const double* x, int incx);
B200_API float cblas_scnrm2(int n, const void* x, int incx);
B200_API double cblas_dznrm2(int n, const void* x, int incx);
B200_API float cblas_sasum(int n, const float* x, int incx);
B200_API double cblas_dasum(int n, const double* x, int incx);
B200_API CBLAS_INDEX cblas_isamax(int n, const float* x, int incx);
B200_API CBLAS_INDEX cblas_idamax(int n, const double* x, int incx);
B200_API CBLAS_INDEX cblas_icamax(int n, const void* x, int incx);
B200_API CBLAS_INDEX cblas_izamax(int n, const void* x, int incx);
B200_API void cblas_saxpy(int n, float alpha, const float* x, int incx, float* y, int incy);
B200_API void cblas_daxpy(int n, double alpha, const double* x, int incx, double* y, int incy);
B200_API void cblas_caxpy(int n, const void* alpha, const void* x, int incx, void* y, int incy);
B200_API void cblas_zaxpy(int n, const void* alpha, const void* x, int incx, void* y, int incy);
B200_API void cblas_sscal(int n, float alpha, float* x, int incx);
B200_API void cblas_dscal(int n, double alpha, double* x, int incx);
B200_API void cblas_scopy(int n, const float* x, int incx, float* y, int incy);
B200_API void cblas_dcopy(int n, const double* x, int incx, double* y, int incy);
B200_API void cblas_sswap(int n, float* x, int incx, float* y, int incy);
B200_API void cblas_dswap(int n, double* x, int incx, double* y, int incy);
B200_API void cblas_sgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, float alpha, const float* a, int lda, const float* x, int incx, float beta, float* y, int incy);
B200_API void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, double alpha, const double* a, int lda, const double* x, int incx, double beta, double* y, int incy);
B200_API void cblas_strsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const float* a, int lda, float* x, int incx);
B200_API void cblas_dtrsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const double* a, int lda, double* x, int incx);

B200_API void cblas_sger(enum CBLAS_ORDER order, int m, int n, float alpha, const float* x, int incx, const float* y, int incy, float* a, int lda);
B200_API void cblas_ssyr(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const float* x, int incx, float* a, int lda);
B200_API void cblas_ssymv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const float* a, int lda, const float* x, int incx, float beta, float* y, int incy);
B200_API void cblas_strmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const float* a, int lda, float* x, int incx);
B200_API void cblas_srot(int n, float* x, int incx, float* y, int incy, float c, float s);
B200_API void cblas_srotg(float* a, float* b, float* c, float* s);
B200_API void cblas_dger(enum CBLAS_ORDER order, int m, int n, double alpha, const double* x, int incx, const double* y, int incy, double* a, int lda);
B200_API void cblas_dsyr(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const double* x, int incx, double* a, int lda);
B200_API void cblas_dsymv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const double* a, int lda, const double* x, int incx, double beta, double* y, int incy);
B200_API void cblas_dtrmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const double* a, int lda, double* x, int incx);
B200_API void cblas_drot(int n, double* x, int incx, double* y, int incy, double c, double s);
B200_API void cblas_drotg(double* a, double* b, double* c, double* s);

/* CBLAS forms of the banded / packed / Hermitian / complex Level 2 (row-major handled by the transposed view: uplo flipped, kl/ku swapped,
 * ConjTrans as conjugate-no-transpose; reference cblas.h DECLARE_CBLAS__GBMV ... __TRMV) */
B200_API void cblas_sgbmv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, int kl, int ku, float alpha, const float* a, int lda, const float* x, int incx, float beta, float* y, int incy);
B200_API void cblas_stbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const float* a, int lda, float* x, int incx);
B200_API void cblas_stbsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const float* a, int lda, float* x, int incx);
B200_API void cblas_stpmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const float* ap, float* x, int incx);
B200_API void cblas_stpsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const float* ap, float* x, int incx);
B200_API void cblas_ssbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, int k, float alpha, const float* a, int lda, const float* x, int incx, float beta, float* y, int incy);
B200_API void cblas_sspmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const float* ap, const float* x, int incx, float beta, float* y, int incy);
B200_API void cblas_ssyr2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const float* x, int incx, const float* y, int incy, float* a, int lda);
B200_API void cblas_sspr(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const float* x, int incx, float* ap);
B200_API void cblas_sspr2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const float* x, int incx, const float* y, int incy, float* ap);
B200_API void cblas_dgbmv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, int kl, int ku, double alpha, const double* a, int lda, const double* x, int incx, double beta, double* y, int incy);
B200_API void cblas_dtbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const double* a, int lda, double* x, int incx);
B200_API void cblas_dtbsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const double* a, int lda, double* x, int incx);
B200_API void cblas_dtpmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const double* ap, double* x, int incx);
B200_API void cblas_dtpsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const double* ap, double* x, int incx);
B200_API void cblas_dsbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, int k, double alpha, const double* a, int lda, const double* x, int incx, double beta, double* y, int incy);
B200_API void cblas_dspmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const double* ap, const double* x, int incx, double beta, double* y, int incy);
B200_API void cblas_dsyr2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const double* x, int incx, const double* y, int incy, double* a, int lda);
B200_API void cblas_dspr(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const double* x, int incx, double* ap);
B200_API void cblas_dspr2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const double* x, int incx, const double* y, int incy, double* ap);
B200_API void cblas_cgbmv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, int kl, int ku, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_ctbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const void* a, int lda, void* x, int incx);
B200_API void cblas_ctbsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const void* a, int lda, void* x, int incx);
B200_API void cblas_ctpmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* ap, void* x, int incx);
B200_API void cblas_ctpsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* ap, void* x, int incx);
B200_API void cblas_chbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, int k, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_chpmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* ap, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_chemv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_cgeru(enum CBLAS_ORDER order, int m, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda);
B200_API void cblas_cgerc(enum CBLAS_ORDER order, int m, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda);
B200_API void cblas_cher(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const void* x, int incx, void* a, int lda);
B200_API void cblas_cher2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda);
B200_API void cblas_chpr(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, float alpha, const void* x, int incx, void* ap);
B200_API void cblas_chpr2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* ap);
B200_API void cblas_ctrmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* a, int lda, void* x, int incx);
B200_API void cblas_cgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_zgbmv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, int kl, int ku, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_ztbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const void* a, int lda, void* x, int incx);
B200_API void cblas_ztbsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, int k, const void* a, int lda, void* x, int incx);
B200_API void cblas_ztpmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* ap, void* x, int incx);
B200_API void cblas_ztpsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* ap, void* x, int incx);
B200_API void cblas_zhbmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, int k, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_zhpmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* ap, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_zhemv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);
B200_API void cblas_zgeru(enum CBLAS_ORDER order, int m, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda);
B200_API void cblas_zgerc(enum CBLAS_ORDER order, int m, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda);
B200_API void cblas_zher(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const void* x, int incx, void* a, int lda);
B200_API void cblas_zher2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda);
B200_API void cblas_zhpr(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, double alpha, const void* x, int incx, void* ap);
B200_API void cblas_zhpr2(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* ap);
B200_API void cblas_ztrmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* a, int lda, void* x, int incx);
B200_API void cblas_zgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, const void* alpha, const void* a, int lda, const void* x, int incx, const void* beta, void* y, int incy);

B200_API void cblas_srotm(int n, float* x, int incx, float* y, int incy, const float* param);
B200_API void cblas_drotm(int n, double* x, int incx, double* y, int incy, const double* param);
B200_API void cblas_srotmg(float* d1, float* d2, float* x1, float y1, float* param);
B200_API void cblas_drotmg(double* d1, double* d2, double* x1, double y1, double* param);
B200_API void cblas_csrot(int n, void* x, int incx, void* y, int incy, float c, float s);
B200_API void cblas_zdrot(int n, void* x, int incx, void* y, int incy, double c, double s);
B200_API void cblas_crotg(void* a, void* b, float* c, void* s);
B200_API void cblas_zrotg(void* a, void* b, double* c, void* s);
B200_API CBLAS_INDEX cblas_isamin(int n, const float* x, int incx);
B200_API CBLAS_INDEX cblas_idamin(int n, const double* x, int incx);
B200_API CBLAS_INDEX cblas_icamin(int n, const void* x, int incx);
B200_API CBLAS_INDEX cblas_izamin(int n, const void* x, int incx);
B200_API double cblas_dsdot(int n, const float* x, int incx, const float* y, int incy);
B200_API float cblas_sdsdot(int n, float sb, const float* x, int incx, const float* y, int incy);
B200_API void cblas_ccopy(int n, const void* x, int incx, void* y, int incy);
B200_API void cblas_zcopy(int n, const void* x, int incx, void* y, int incy);
B200_API void cblas_cswap(int n, void* x, int incx, void* y, int incy);
B200_API void cblas_zswap(int n, void* x, int incx, void* y, int incy);
B200_API void cblas_cscal(int n, const void* alpha, void* x, int incx);
B200_API void cblas_zscal(int n, const void* alpha, void* x, int incx);
B200_API void cblas_csscal(int n, float alpha, void* x, int incx);
B200_API void cblas_zdscal(int n, double alpha, void* x, int incx);
B200_API float cblas_scasum(int n, const void* x, int incx);
B200_API double cblas_dzasum(int n, const void* x, int incx);
B200_API void cblas_ctrsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* a, int lda, void* x, int incx);
B200_API void cblas_ztrsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const void* a, int lda, void* x, int incx);

/* complex SYMM / HEMM / SYR2K / HERK / HER2K (reference cblas.h DECLARE_CBLAS__SYMM, __HEMM, __SYR2K, __HERK, __HER2K; complex scalars by pointer) */
B200_API void cblas_csymm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);
B200_API void cblas_chemm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);
B200_API void cblas_csyr2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);
B200_API void cblas_cherk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, float alpha, const void* a, int lda, float beta, void* c, int ldc);
B200_API void cblas_cher2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, const void* a, int lda, const void* b, int ldb, float beta, void* c, int ldc);
B200_API void cblas_zsymm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);
B200_API void cblas_zhemm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);
B200_API void cblas_zsyr2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc);
B200_API void cblas_zherk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, double alpha, const void* a, int lda, double beta, void* c, int ldc);
B200_API void cblas_zher2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, const void* a, int lda, const void* b, int ldb, double beta, void* c, int ldc);

/* ------------------------------ 3. allocator symbols ------------------------------ */
/* malloc / calloc / realloc / free are exported with their libc prototypes (<stdlib.h>);
 * reference lib/obj_tracker.c:789 (malloc), :842 (calloc), :902 (realloc), :948 (free).
 * Also exported, which the reference leaves to glibc (so aligned operands are never tracked there): posix_memalign,
 * aligned_alloc, memalign, valloc -- a managed block is handed out when the placement heuristic says so and its base
 * satisfies the alignment -- and malloc_usable_size (answers for managed blocks from the registry). */

/* ------------------------------ 4. control API ------------------------------ */
struct b200blas_stats {
    uint64_t hits, misses;           /* operands used in place / staged (reference b2c_hits, b2c_misses) */
    uint64_t calls, h2d_bytes, d2h_bytes, prefetch_bytes;
    uint64_t managed_allocs, managed_frees, managed_bytes_live;
};
typedef void (*b200blas_xerbla_fn)(const char* srname, int* info, size_t srname_len);

B200_API int b200blas_version(void);
B200_API void b200blas_set_options(const char* opts);            /* same grammar as BLAS2CUDA_OPTIONS */
B200_API void b200blas_set_xerbla(b200blas_xerbla_fn fn);        /* NULL: look xerbla_ up by symbol (default) */
B200_API void b200blas_set_stream(void* cuda_stream);            /* run this thread's calls on a caller stream (0 = legacy default stream) */
B200_API void b200blas_reset_stream(void);                       /* back to the library's own per-thread stream */
B200_API void b200blas_set_sync(int on);                         /* 0: do not wait for completion before returning */
B200_API void b200blas_synchronize(void);
B200_API const char* b200blas_last_variant(void);                /* kernel variant the calling thread's last call used */
B200_API void b200blas_force_variant(const char* name);          /* NULL/"auto": size-based selection */
B200_API void b200blas_get_stats(struct b200blas_stats* out);
B200_API void* b200blas_malloc_managed(size_t bytes);            /* tracked managed block (what malloc() hands out) */
B200_API void b200blas_free_managed(void* p);
B200_API int b200blas_is_tracked(const void* p);
B200_API int b200blas_tracker_decision(unsigned long long nth, size_t request);   /* would the nth allocation of `request` bytes be managed (current heuristic)? */
B200_API int b200blas_device_count(void);
/* Partitioned Level-3 calls (option devices=<n>, csrc/multi_gemm.cu): the hop list a call of this shape issues -- 7 ints per hop
 * (kind 0 A row-group / 1 B column band, grid row or column, piece, offset, length, forwarding slot or -1 = origin, receiving
 * slot); returns the hop count.  Pure host logic.  b200blas_mg_geometry: [r0, r1, c0, c1) of a slot's C tile.
 * b200blas_mg_stats: calls, devices, bytes read from the origin, bytes forwarded between devices, hops -- since load. */
B200_API int b200blas_mg_plan(int ndev, long long m, long long n, int host_source, int* out, int cap);
B200_API void b200blas_mg_geometry(int ndev, long long m, long long n, int slot, long long* out4);
B200_API void b200blas_mg_stats(unsigned long long* out5);
/* ?syrk_, ?trsm_, ?trmm_ with devices=<n> (csrc/multi_level3.cu; reference blas_level3/syrk.cc:43-76, trsm.cc:40-73, trmm.cc:42-79).
 * b200blas_ml3_strips: the ndev + 1 boundaries of the equal-area strips a partitioned ?syrk_ of order n cuts its triangle into.
 * b200blas_ml3_plan: hop list as above; which = 0: ?syrk_ row pieces of op(A) (grid index = strip the piece lies in, consumed by
 * slots 0..strip; offset = global row), which = 1: ?trsm_/?trmm_ column groups of the order-n triangle (consumed by every slot). */
B200_API void b200blas_ml3_strips(long long n, int ndev, long long* out);
B200_API int b200blas_ml3_plan(int which, int ndev, long long n, int host_source, int* out, int cap);
/* The blocked Cholesky workload of BASELINE.json configs[3] (diagonal-block factorisation + DTRSM panel + DSYRK/DGEMM trailing
 * update, look-ahead 1) as one call over the GPUs selected with devices=<n>: lower factor in place, LAPACK info returned. */
B200_API int b200blas_cholesky_lower(int n, double* a, long long lda, int nb);
/* FP64 tensor-pipe (DMMA) ceiling measured in this process: sustained TFLOP/s over ~`seconds` of register-resident mma.sync f64
 * loops (no memory traffic); *burst (may be NULL) = best single launch.  The denominator of bench.py's tensor rooflines. */
B200_API double b200blas_probe_fp64_tflops(double seconds, double* burst);
B200_API int b200blas_residency(const void* p, size_t bytes);     /* device ordinal a managed range was last prefetched to; -1 host; -2 unknown */
/* D := alpha*op(A)*op(B) + beta*C on device pointers with a separate output (may be peer-mapped) */
B200_API void b200blas_dgemm_out(char transa, char transb, int m, int n, int k, double alpha, const double* a, long long lda, const double* b, long long ldb, double beta, const double* c, long long ldc, double* d, long long ldd);
/* the same for a rank whose operand panels are still arriving over NVLink: a tile reads rows [g*a_group, ...) of op(A) only
 * after aflags[g] >= epoch and columns [h*b_group, ...) of op(B) only after bflags[h] >= epoch (flags: device memory of this
 * GPU, written remotely; group sizes multiples of 128) */
B200_API void b200blas_dgemm_out_flagged(char transa, char transb, int m, int n, int k, double alpha, const double* a, long long lda, const double* b, long long ldb, double beta, const double* c, long long ldc, double* d, long long ldd, const unsigned* aflags, int a_group, const unsigned* bflags, int b_group, unsigned epoch);
/* copy-engine building blocks of the panel push: strided device->device (peer) copy, stream-ordered flag write, memset */
B200_API void b200blas_copy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height, void* cuda_stream);
B200_API void b200blas_write_flag_async(void* dst_flag, unsigned value, void* cuda_stream);
B200_API void b200blas_memset_async(void* dst, int byte, size_t bytes, void* cuda_stream);
B200_API void b200blas_wait_flag_async(const void* flag, unsigned value, void* cuda_stream);   /* stream waits until *flag >= value */
/* Lower Cholesky factor in place on the library's own Level-3 kernels (the diagonal-block step of the blocked Cholesky
 * workload, BASELINE.json configs[3]); a may be host, managed or device memory.  Returns LAPACK-style info. */
B200_API int b200blas_dpotrf_lower(int n, double* a, long long lda);
/* raw device memory + CUDA IPC: lets another process's GEMM epilogue store C tiles into this allocation */
B200_API void* b200blas_device_malloc(size_t bytes);
B200_API void b200blas_device_free(void* p);
B200_API int b200blas_ipc_get_handle(void* dev_ptr, void* handle64);   /* returns handle size (64) or -1 */
B200_API void* b200blas_ipc_open(const void* handle64);                /* peer-mapped pointer or NULL */
B200_API void b200blas_ipc_close(void* p);
B200_API void b200blas_print_help(void);
B200_API void b200blas_entry(void);                              /* ELF entry: prints option help (reference entry.c) */

#ifdef __cplusplus
}
#endif
#endif /* B200BLAS_H */
