// kernels.h -- device-level routines (operands are device-accessible pointers; all work is
// enqueued on `s`, nothing synchronises).  The ABI layers (fortran_l3.cu, fortran_l12.cu, cblas_l3.cu, and the
// entry points at the bottom of level2_struct.cu / level1_more.cu) sit above this and own argument
// checking, residency and the synchronous return.
#pragma once
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <stdint.h>

namespace b200 {

enum TriMask { MASK_FULL = 0, MASK_LOWER = 1, MASK_UPPER = 2 };

// Which kernel a call was routed to -- recorded for tests / debug_exec tracing.
enum Variant {
    VAR_NONE = 0,
    VAR_SCALE_ONLY,        // alpha==0 or k==0: C := beta*C on the device
    VAR_GENERIC_TILE,      // type-generic register-tiled kernel (any alignment, any type)
    VAR_DMMA_TMA,          // FP64 tensor pipe, TMA-staged warp-specialised pipeline
    VAR_DMMA_LDG,          // same compute core, producer warps stage tiles with LDG (unaligned operands)
    VAR_TF32X3_TCGEN05,    // SGEMM: 3xTF32 split on tcgen05 with TMEM accumulators
    VAR_COUNT
};
const char* variant_name(int v);
extern thread_local int last_variant;      // set by every *_dev launcher
extern int force_variant;                  // BLAS2CUDA_OPTIONS variant=...; VAR_NONE = size-based selection

// ---- Level 3 ----  C := alpha*op(A)*op(B) + beta*C on the part of C selected by `mask`
void dgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
               const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int mask = MASK_FULL);
void dgemm_out_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
                   const double* B, int64_t ldb, double beta, const double* Cin, int64_t ldc, double* D, int64_t ldd,
                   int mask = MASK_FULL);
void dgemm_set_panel_flags(const uint32_t* aflags, int a_group, const uint32_t* bflags, int b_group, uint32_t epoch);   // applies to this thread's next dgemm_out_dev
void sgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, float alpha, const float* A, int64_t lda,
               const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mask = MASK_FULL);
void zgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuDoubleComplex alpha,
               const cuDoubleComplex* A, int64_t lda, const cuDoubleComplex* B, int64_t ldb, cuDoubleComplex beta,
               cuDoubleComplex* C, int64_t ldc, int mask = MASK_FULL);
void cgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuFloatComplex alpha, const cuFloatComplex* A,
               int64_t lda, const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C,
               int64_t ldc, int mask = MASK_FULL);

// typed front door to the four GEMMs
template <typename T> void gemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, T alpha, const T* A, int64_t lda, const T* B,
                                    int64_t ldb, T beta, T* C, int64_t ldc, int mask = MASK_FULL);
// C := alpha*op(A)*op(A)^T + beta*C on one triangle (trans 'N': A n x k; 'T': A k x n).  herm: C/A^H variant.
template <typename T> void syrk_dev(cudaStream_t s, char uplo, char trans, int n, int k, T alpha, const T* A, int64_t lda, T beta, T* C,
                                    int64_t ldc);
// C := alpha*A*B + beta*C ('L') or alpha*B*A + beta*C ('R'), A symmetric (herm: Hermitian) with one triangle stored
template <typename T> void symm_dev(cudaStream_t s, bool herm, char side, char uplo, int m, int n, T alpha, const T* A, int64_t lda, const T* B,
                                    int64_t ldb, T beta, T* C, int64_t ldc);
template <typename T> void syr2k_dev(cudaStream_t s, char uplo, char trans, int n, int k, T alpha, const T* A, int64_t lda, const T* B, int64_t ldb,
                                     T beta, T* C, int64_t ldc);
template <typename T, typename R> void herk_dev(cudaStream_t s, char uplo, char trans, int n, int k, R alpha, const T* A, int64_t lda, R beta, T* C,
                                                int64_t ldc);
template <typename T, typename R> void her2k_dev(cudaStream_t s, char uplo, char trans, int n, int k, T alpha, const T* A, int64_t lda, const T* B,
                                                 int64_t ldb, R beta, T* C, int64_t ldc);
// B := alpha*op(A)*B (side 'L') or alpha*B*op(A) ('R'), A triangular; in place, blocked on the GEMM tiles
template <typename T> void trmm_dev(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* A,
                                    int64_t lda, T* B, int64_t ldb);
// solve op(A)*X = alpha*B ('L') or X*op(A) = alpha*B ('R'), X overwrites B
template <typename T> void trsm_dev(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* A,
                                    int64_t lda, T* B, int64_t ldb);

// lower Cholesky factor of an n x n block in place (potrf.cu); *info_dev (device int, zero on entry) receives base + the
// 1-based index of the first non-positive pivot
void potrf_lower_dev(cudaStream_t s, int n, double* A, int64_t lda, int* info_dev, int base = 0);

// ---- Level 1 ----  x, y device-accessible; incx/incy are the BLAS increments (may be negative where
// netlib allows it).  Reductions write their result to `out` (device memory) deterministically:
// fixed grid, fixed combination order, independent of scheduling.
template <typename T> void dot_dev(cudaStream_t s, int64_t n, const T* x, int64_t incx, const T* y, int64_t incy, T* out, bool conj_x);
template <typename T, typename R> void nrm2_dev(cudaStream_t s, int64_t n, const T* x, int64_t incx, R* out);
template <typename T, typename R> void asum_dev(cudaStream_t s, int64_t n, const T* x, int64_t incx, R* out);
template <typename T> void iamax_dev(cudaStream_t s, int64_t n, const T* x, int64_t incx, long long* out);   // 0-based, -1 if none
template <typename T> void iamin_dev(cudaStream_t s, int64_t n, const T* x, int64_t incx, long long* out);   // 0-based, -1 if none
void dsdot_dev(cudaStream_t s, int64_t n, const float* x, int64_t incx, const float* y, int64_t incy, double sb, void* out, bool out_double);
template <typename T> void axpy_dev(cudaStream_t s, int64_t n, T alpha, const T* x, int64_t incx, T* y, int64_t incy);
template <typename T, typename S> void scal_dev(cudaStream_t s, int64_t n, S alpha, T* x, int64_t incx);
template <typename T> void copy_dev(cudaStream_t s, int64_t n, const T* x, int64_t incx, T* y, int64_t incy);
template <typename T> void swap_dev(cudaStream_t s, int64_t n, T* x, int64_t incx, T* y, int64_t incy);

// ---- Level 2 ----
template <typename T> void gemv_dev(cudaStream_t s, char trans, int m, int n, T alpha, const T* A, int64_t lda, const T* x,
                                    int64_t incx, T beta, T* y, int64_t incy);
template <typename T> void trsv_dev(cudaStream_t s, char uplo, char trans, char diag, int n, const T* A, int64_t lda, T* x,
                                    int64_t incx);
// the same solve through the panel solver of level2_struct.cu (32-wide diagonal blocks staged ahead of the dependency chain)
template <typename T> void trsv_struct_dev(cudaStream_t s, char uplo, char trans, char diag, int n, const T* A, int64_t lda, T* x,
                                           int64_t incx);

}  // namespace b200
