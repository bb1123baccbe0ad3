// fortran_l3.cu -- Fortran-ABI Level-3 entry points: netlib argument checks and quick returns on
// the host (reference blas_level3/*.cc `xxx_check`), operand residency, device kernel, synchronous
// return.  No size cutoff to a CPU BLAS (reference gemm.cc:129-141 gemm_perf_check): small
// problems pick a small-tile GPU variant instead.
#include "abi_common.h"
#include "../../include/b200blas.h"

using namespace b200;

namespace {

template <typename T> struct GemmDev;
template <> struct GemmDev<float> { static constexpr auto fn = sgemm_dev; };
template <> struct GemmDev<double> { static constexpr auto fn = dgemm_dev; };
template <> struct GemmDev<cuFloatComplex> { static constexpr auto fn = cgemm_dev; };
template <> struct GemmDev<cuDoubleComplex> { static constexpr auto fn = zgemm_dev; };

// reference gemm.cc:87-127 (gemm_check) + :46-83 (_b2c_gemm)
template <typename T>
void gemm_entry(const char* name, const char* transa, const char* transb, const int* m, const int* n, const int* k,
                const T* alpha, const T* a, const int* lda, const T* b, const int* ldb, const T* beta, T* c,
                const int* ldc) {
    const bool nota = lsame(transa, 'N'), notb = lsame(transb, 'N');
    const int nrowa = nota ? *m : *k, nrowb = notb ? *k : *n;
    int info = 0;
    if (!nota && !lsame(transa, 'C') && !lsame(transa, 'T')) info = 1;
    else if (!notb && !lsame(transb, 'C') && !lsame(transb, 'T')) info = 2;
    else if (*m < 0) info = 3;
    else if (*n < 0) info = 4;
    else if (*k < 0) info = 5;
    else if (*lda < imax(1, nrowa)) info = 8;
    else if (*ldb < imax(1, nrowb)) info = 10;
    else if (*ldc < imax(1, *m)) info = 13;
    if (info) { call_xerbla(name, info); return; }
    if (*m == 0 || *n == 0 || ((is0(*alpha) || *k == 0) && is1(*beta))) return;

    CallScope scope;
    const bool scale_only = is0(*alpha) || *k == 0;
    const char ta = nota ? 'N' : (lsame(transa, 'T') ? 'T' : 'C');
    const char tb = notb ? 'N' : (lsame(transb, 'T') ? 'T' : 'C');
    // A is lda x (nota ? k : m), B is ldb x (notb ? n : k)  (reference compute_size, gemm.cc:36-44,
    // done in 64-bit here: the reference's int byte counts overflow at n=16384)
    Operand oa(scale_only ? nullptr : a, nrowa, nota ? *k : *m, *lda, sizeof(T), ACC_IN);
    Operand ob(scale_only ? nullptr : b, nrowb, notb ? *n : *k, *ldb, sizeof(T), ACC_IN);
    Operand oc(c, *m, *n, *ldc, sizeof(T), is0(*beta) ? ACC_OUT : ACC_INOUT);
    GemmDev<T>::fn(current_stream(), ta, tb, *m, *n, *k, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ob.dev(), ob.ld(),
                   *beta, (T*)oc.dev(), oc.ld(), MASK_FULL);
    oc.release();
    log_exec(name, "%c%c m=%d n=%d k=%d lda=%d ldb=%d ldc=%d", ta, tb, *m, *n, *k, *lda, *ldb, *ldc);
}

}  // namespace

extern "C" {

void sgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const float* alpha,
            const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c, const int* ldc) {
    gemm_entry<float>("sgemm_", transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha,
            const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc) {
    gemm_entry<double>("dgemm_", transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const b200_c32* alpha,
            const b200_c32* a, const int* lda, const b200_c32* b, const int* ldb, const b200_c32* beta, b200_c32* c, const int* ldc) {
    gemm_entry<cuFloatComplex>("cgemm_", transa, transb, m, n, k, (const cuFloatComplex*)alpha, (const cuFloatComplex*)a, lda,
                               (const cuFloatComplex*)b, ldb, (const cuFloatComplex*)beta, (cuFloatComplex*)c, ldc);
}
void zgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const b200_c64* alpha,
            const b200_c64* a, const int* lda, const b200_c64* b, const int* ldb, const b200_c64* beta, b200_c64* c, const int* ldc) {
    gemm_entry<cuDoubleComplex>("zgemm_", transa, transb, m, n, k, (const cuDoubleComplex*)alpha, (const cuDoubleComplex*)a, lda,
                                (const cuDoubleComplex*)b, ldb, (const cuDoubleComplex*)beta, (cuDoubleComplex*)c, ldc);
}

}  // extern "C"
