// cblas_l3.cu -- CBLAS Level-3 entry points (reference cblas.h:657-824).  Row-major is mapped onto
// the column-major core by swapping operands: C^T = op(B)^T op(A)^T.
#include "abi_common.h"
#include "../../include/b200blas.h"

using namespace b200;

static inline char tr(enum CBLAS_TRANSPOSE t) { return t == CblasNoTrans ? 'N' : (t == CblasTrans ? 'T' : (t == CblasConjTrans ? 'C' : '?')); }

#define B200_CBLAS_GEMM(NAME, F77, T, AT, DEREF)                                                                  \
    void NAME(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n,    \
              int k, AT alpha, const T* a, int lda, const T* b, int ldb, AT beta, T* c, int ldc) {                \
        char ta = tr(transa), tb = tr(transb);                                                                    \
        if (order == CblasColMajor)                                                                               \
            F77(&ta, &tb, &m, &n, &k, DEREF(alpha), a, &lda, b, &ldb, DEREF(beta), c, &ldc);                     \
        else                                                                                                      \
            F77(&tb, &ta, &n, &m, &k, DEREF(alpha), b, &ldb, a, &lda, DEREF(beta), c, &ldc);                     \
    }
#define B200_ADDR(x) (&x)
#define B200_ASIS(x) (x)

extern "C" {
B200_CBLAS_GEMM(cblas_sgemm, sgemm_, float, float, B200_ADDR)
B200_CBLAS_GEMM(cblas_dgemm, dgemm_, double, double, B200_ADDR)
void cblas_cgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, int k,
                 const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc) {
    char ta = tr(transa), tb = tr(transb);
    if (order == CblasColMajor)
        cgemm_(&ta, &tb, &m, &n, &k, (const b200_c32*)alpha, (const b200_c32*)a, &lda, (const b200_c32*)b, &ldb, (const b200_c32*)beta, (b200_c32*)c, &ldc);
    else
        cgemm_(&tb, &ta, &n, &m, &k, (const b200_c32*)alpha, (const b200_c32*)b, &ldb, (const b200_c32*)a, &lda, (const b200_c32*)beta, (b200_c32*)c, &ldc);
}
void cblas_zgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE transa, enum CBLAS_TRANSPOSE transb, int m, int n, int k,
                 const void* alpha, const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc) {
    char ta = tr(transa), tb = tr(transb);
    if (order == CblasColMajor)
        zgemm_(&ta, &tb, &m, &n, &k, (const b200_c64*)alpha, (const b200_c64*)a, &lda, (const b200_c64*)b, &ldb, (const b200_c64*)beta, (b200_c64*)c, &ldc);
    else
        zgemm_(&tb, &ta, &n, &m, &k, (const b200_c64*)alpha, (const b200_c64*)b, &ldb, (const b200_c64*)a, &lda, (const b200_c64*)beta, (b200_c64*)c, &ldc);
}
}
