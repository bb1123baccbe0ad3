// gemm_other.cu -- CGEMM launcher (reference gemm.cc:181-198); SGEMM lives in gemm_f32.cu, ZGEMM in gemm_z.cu.
#include "gemm_generic.cuh"
#include "runtime.h"

namespace b200 {

void cgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuFloatComplex alpha, const cuFloatComplex* A,
               int64_t lda, const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C, int64_t ldc,
               int mask) {
    if (m <= 0 || n <= 0) return;
    if (num<cuFloatComplex>::is_zero(alpha) || k <= 0) { scale_matrix<cuFloatComplex>(s, m, n, beta, C, ldc, mask); last_variant = VAR_SCALE_ONLY; return; }
    gemm_generic_launch<cuFloatComplex>(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
}
}  // namespace b200
