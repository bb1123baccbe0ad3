// kernels.h -- device-level routines (operands are device-accessible pointers; all work is
// enqueued on `s`, nothing synchronises).  The ABI layers (fortran_abi.cu, cblas_abi.cu) sit
// above this and own argument checking, residency and the synchronous return.
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
void sgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, float alpha, const float* A, int64_t lda,
               const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mask = MASK_FULL);
void zgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuDoubleComplex alpha,
               const cuDoubleComplex* A, int64_t lda, const cuDoubleComplex* B, int64_t ldb, cuDoubleComplex beta,
               cuDoubleComplex* C, int64_t ldc, int mask = MASK_FULL);
void cgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuFloatComplex alpha, const cuFloatComplex* A,
               int64_t lda, const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C,
               int64_t ldc, int mask = MASK_FULL);

template <typename T> struct Scalar { using type = T; };

}  // namespace b200
