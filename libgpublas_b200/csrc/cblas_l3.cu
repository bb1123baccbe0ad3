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
// ---- CBLAS SYRK / TRSM / TRMM / SYMM / SYR2K (reference cblas.h:693-824).  Row-major C is the column-major C^T:
//   syrk/syr2k: the other triangle of C^T and the other transpose of A;  trsm/trmm/symm: the other side and triangle.
static inline char up(enum CBLAS_UPLO u) { return u == CblasUpper ? 'U' : (u == CblasLower ? 'L' : '?'); }
static inline char sd(enum CBLAS_SIDE x) { return x == CblasLeft ? 'L' : (x == CblasRight ? 'R' : '?'); }
static inline char dg(enum CBLAS_DIAG d) { return d == CblasUnit ? 'U' : (d == CblasNonUnit ? 'N' : '?'); }
static inline char flip_up(char u) { return u == 'U' ? 'L' : (u == 'L' ? 'U' : u); }
static inline char flip_sd(char x) { return x == 'L' ? 'R' : (x == 'R' ? 'L' : x); }
static inline char flip_tr(char t) { return t == 'N' ? 'T' : (t == 'T' || t == 'C' ? 'N' : t); }

#define B200_CBLAS_REAL(P, T)                                                                                                     \
    void cblas_##P##syrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, T alpha,         \
                         const T* a, int lda, T beta, T* c, int ldc) {                                                            \
        char u = up(uplo), t = tr(trans);                                                                                         \
        if (order == CblasRowMajor) { u = flip_up(u); t = flip_tr(t); }                                                           \
        P##syrk_(&u, &t, &n, &k, &alpha, a, &lda, &beta, c, &ldc);                                                                \
    }                                                                                                                             \
    void cblas_##P##syr2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, T alpha,        \
                          const T* a, int lda, const T* b, int ldb, T beta, T* c, int ldc) {                                      \
        char u = up(uplo), t = tr(trans);                                                                                         \
        if (order == CblasRowMajor) { u = flip_up(u); t = flip_tr(t); }                                                           \
        P##syr2k_(&u, &t, &n, &k, &alpha, a, &lda, b, &ldb, &beta, c, &ldc);                                                      \
    }                                                                                                                             \
    void cblas_##P##symm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, T alpha, const T* a,   \
                         int lda, const T* b, int ldb, T beta, T* c, int ldc) {                                                   \
        char s = sd(side), u = up(uplo);                                                                                          \
        if (order == CblasRowMajor) { s = flip_sd(s); u = flip_up(u); P##symm_(&s, &u, &n, &m, &alpha, a, &lda, b, &ldb, &beta, c, &ldc); } \
        else P##symm_(&s, &u, &m, &n, &alpha, a, &lda, b, &ldb, &beta, c, &ldc);                                                  \
    }                                                                                                                             \
    void cblas_##P##trsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa,         \
                         enum CBLAS_DIAG diag, int m, int n, T alpha, const T* a, int lda, T* b, int ldb) {                       \
        char s = sd(side), u = up(uplo), t = tr(transa), d = dg(diag);                                                            \
        if (order == CblasRowMajor) { s = flip_sd(s); u = flip_up(u); P##trsm_(&s, &u, &t, &d, &n, &m, &alpha, a, &lda, b, &ldb); } \
        else P##trsm_(&s, &u, &t, &d, &m, &n, &alpha, a, &lda, b, &ldb);                                                          \
    }                                                                                                                             \
    void cblas_##P##trmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa,         \
                         enum CBLAS_DIAG diag, int m, int n, T alpha, const T* a, int lda, T* b, int ldb) {                       \
        char s = sd(side), u = up(uplo), t = tr(transa), d = dg(diag);                                                            \
        if (order == CblasRowMajor) { s = flip_sd(s); u = flip_up(u); P##trmm_(&s, &u, &t, &d, &n, &m, &alpha, a, &lda, b, &ldb); } \
        else P##trmm_(&s, &u, &t, &d, &m, &n, &alpha, a, &lda, b, &ldb);                                                          \
    }
B200_CBLAS_REAL(s, float)
B200_CBLAS_REAL(d, double)

// complex: scalars by pointer (standard CBLAS; the reference header declares them by value, cblas.h:662-677)
#define B200_CBLAS_CPLX(P, CT)                                                                                                    \
    void cblas_##P##syrk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, \
                         const void* a, int lda, const void* beta, void* c, int ldc) {                                            \
        char u = up(uplo), t = tr(trans);                                                                                         \
        if (order == CblasRowMajor) { u = flip_up(u); t = flip_tr(t); }                                                           \
        P##syrk_(&u, &t, &n, &k, (const CT*)alpha, (const CT*)a, &lda, (const CT*)beta, (CT*)c, &ldc);                            \
    }                                                                                                                             \
    void cblas_##P##trsm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa,         \
                         enum CBLAS_DIAG diag, int m, int n, const void* alpha, const void* a, int lda, void* b, int ldb) {       \
        char s = sd(side), u = up(uplo), t = tr(transa), d = dg(diag);                                                            \
        if (order == CblasRowMajor) { s = flip_sd(s); u = flip_up(u); P##trsm_(&s, &u, &t, &d, &n, &m, (const CT*)alpha, (const CT*)a, &lda, (CT*)b, &ldb); } \
        else P##trsm_(&s, &u, &t, &d, &m, &n, (const CT*)alpha, (const CT*)a, &lda, (CT*)b, &ldb);                                \
    }                                                                                                                             \
    void cblas_##P##trmm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE transa,         \
                         enum CBLAS_DIAG diag, int m, int n, const void* alpha, const void* a, int lda, void* b, int ldb) {       \
        char s = sd(side), u = up(uplo), t = tr(transa), d = dg(diag);                                                            \
        if (order == CblasRowMajor) { s = flip_sd(s); u = flip_up(u); P##trmm_(&s, &u, &t, &d, &n, &m, (const CT*)alpha, (const CT*)a, &lda, (CT*)b, &ldb); } \
        else P##trmm_(&s, &u, &t, &d, &m, &n, (const CT*)alpha, (const CT*)a, &lda, (CT*)b, &ldb);                                \
    }
B200_CBLAS_CPLX(c, b200_c32)
B200_CBLAS_CPLX(z, b200_c64)

// complex SYMM / HEMM / SYR2K / HERK / HER2K (reference cblas.h DECLARE_CBLAS__SYMM, __HEMM, __SYR2K, __HERK, __HER2K).
// Row-major: C^T is the column-major matrix.  symm/hemm: other side and triangle, m <-> n (the row-major array of a Hermitian
// A is the column-major conj(A) = A^T, which is what the transposed product needs).  syr2k: other triangle, N <-> T.
// herk / her2k: conj(C) = C^T, so NoTrans <-> ConjTrans on the other triangle, and her2k's alpha is conjugated.
static inline char flip_herm_tr(char t) { return t == 'N' ? 'C' : (t == 'C' ? 'N' : '?'); }
#define B200_CBLAS_CPLX_MORE(P, CT, RT)                                                                                           \
    void cblas_##P##symm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, const void* alpha,     \
                         const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc) {                    \
        char s = sd(side), u = up(uplo);                                                                                          \
        if (order == CblasRowMajor) { s = flip_sd(s); u = flip_up(u); P##symm_(&s, &u, &n, &m, (const CT*)alpha, (const CT*)a, &lda, (const CT*)b, &ldb, (const CT*)beta, (CT*)c, &ldc); } \
        else P##symm_(&s, &u, &m, &n, (const CT*)alpha, (const CT*)a, &lda, (const CT*)b, &ldb, (const CT*)beta, (CT*)c, &ldc);   \
    }                                                                                                                             \
    void cblas_##P##hemm(enum CBLAS_ORDER order, enum CBLAS_SIDE side, enum CBLAS_UPLO uplo, int m, int n, const void* alpha,     \
                         const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc) {                    \
        char s = sd(side), u = up(uplo);                                                                                          \
        if (order == CblasRowMajor) { s = flip_sd(s); u = flip_up(u); P##hemm_(&s, &u, &n, &m, (const CT*)alpha, (const CT*)a, &lda, (const CT*)b, &ldb, (const CT*)beta, (CT*)c, &ldc); } \
        else P##hemm_(&s, &u, &m, &n, (const CT*)alpha, (const CT*)a, &lda, (const CT*)b, &ldb, (const CT*)beta, (CT*)c, &ldc);   \
    }                                                                                                                             \
    void cblas_##P##syr2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, \
                          const void* a, int lda, const void* b, int ldb, const void* beta, void* c, int ldc) {                   \
        char u = up(uplo), t = tr(trans);                                                                                         \
        if (order == CblasRowMajor) { u = flip_up(u); t = t == 'N' ? 'T' : (t == 'T' ? 'N' : '?'); }                              \
        P##syr2k_(&u, &t, &n, &k, (const CT*)alpha, (const CT*)a, &lda, (const CT*)b, &ldb, (const CT*)beta, (CT*)c, &ldc);       \
    }                                                                                                                             \
    void cblas_##P##herk(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, RT alpha,        \
                         const void* a, int lda, RT beta, void* c, int ldc) {                                                     \
        char u = up(uplo), t = tr(trans);                                                                                         \
        if (order == CblasRowMajor) { u = flip_up(u); t = flip_herm_tr(t); }                                                      \
        P##herk_(&u, &t, &n, &k, &alpha, (const CT*)a, &lda, &beta, (CT*)c, &ldc);                                                \
    }                                                                                                                             \
    void cblas_##P##her2k(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, int n, int k, const void* alpha, \
                          const void* a, int lda, const void* b, int ldb, RT beta, void* c, int ldc) {                            \
        char u = up(uplo), t = tr(trans);                                                                                         \
        CT al = *(const CT*)alpha;                                                                                                \
        if (order == CblasRowMajor) { u = flip_up(u); t = flip_herm_tr(t); al.im = -al.im; }                                      \
        P##her2k_(&u, &t, &n, &k, &al, (const CT*)a, &lda, (const CT*)b, &ldb, &beta, (CT*)c, &ldc);                              \
    }
B200_CBLAS_CPLX_MORE(c, b200_c32, float)
B200_CBLAS_CPLX_MORE(z, b200_c64, double)
}
