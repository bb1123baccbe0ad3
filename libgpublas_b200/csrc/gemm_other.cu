// gemm_other.cu -- SGEMM / CGEMM launchers (reference gemm.cc:143-160, :181-198); ZGEMM lives in gemm_z.cu.
#include "gemm_generic.cuh"
#include "runtime.h"

namespace b200 {

void sgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, float alpha, const float* A, int64_t lda,
               const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mask) {
    if (m <= 0 || n <= 0) return;
    if (alpha == 0.f || k <= 0) { scale_matrix<float>(s, m, n, beta, C, ldc, mask); last_variant = VAR_SCALE_ONLY; return; }
    gemm_generic_launch<float>(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
}
void cgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuFloatComplex alpha, const cuFloatComplex* A,
               int64_t lda, const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C, int64_t ldc,
               int mask) {
    if (m <= 0 || n <= 0) return;
    if (num<cuFloatComplex>::is_zero(alpha) || k <= 0) { scale_matrix<cuFloatComplex>(s, m, n, beta, C, ldc, mask); last_variant = VAR_SCALE_ONLY; return; }
    gemm_generic_launch<cuFloatComplex>(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
}
}  // namespace b200
