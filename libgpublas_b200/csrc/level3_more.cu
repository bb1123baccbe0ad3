// level3_more.cu -- the rest of the reference's exported Level-3 family on the same GEMM tiles
// (SURVEY.md section 8(f) rank 1): SYMM / HEMM (reference blas_level3/symm.cc:43-68, hemm.cc:42-67),
// SYR2K (syr2k.cc:40-65), HERK (herk.cc:39-62), HER2K (her2k.cc:44-70) -- all forwards to cuBLAS there.
//
//  SYR2K  two masked GEMM launches on the referenced triangle: C := alpha*A*B^T + beta*C, then
//         C += alpha*B*A^T (the tiles outside the triangle exit, diagonal tiles store their half).
//  HERK / HER2K  the same with conjugate transposes and a real beta; a one-thread-per-row kernel then clears
//         the imaginary part of the diagonal, which netlib defines as exactly zero (the input's imaginary
//         diagonal is never used: beta is real, so it cannot leak into the real part).
//  SYMM / HEMM  the stored triangle is expanded once into a full matrix in the workspace (reflected, and
//         conjugated for HEMM, diagonal made real) and the product is a plain GEMM: n^2 extra bytes of HBM
//         traffic against 2*m*n^2 flops on the tensor pipe.
#include "common.cuh"
#include "kernels.h"
#include "gemm_generic.cuh"
#include "runtime.h"

namespace b200 {

template <typename T> __device__ __forceinline__ T clear_imag(T v);
template <> __device__ __forceinline__ cuFloatComplex clear_imag(cuFloatComplex v) { v.y = 0.f; return v; }
template <> __device__ __forceinline__ cuDoubleComplex clear_imag(cuDoubleComplex v) { v.y = 0.0; return v; }
template <> __device__ __forceinline__ float clear_imag(float v) { return v; }
template <> __device__ __forceinline__ double clear_imag(double v) { return v; }

template <typename T> __global__ void real_diag_kernel(int n, T* C, int64_t ldc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) C[i + (int64_t)i * ldc] = clear_imag<T>(C[i + (int64_t)i * ldc]);
}
template <typename T> static void real_diag(cudaStream_t s, int n, T* C, int64_t ldc) {
    if (n > 0) real_diag_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, C, ldc);
}

// full (n x n, ld = ldf) := the symmetric / Hermitian matrix whose `upper` or lower triangle is stored in A
template <typename T, bool HERM>
__global__ void expand_sym_kernel(int n, const T* __restrict__ A, int64_t lda, bool upper, T* __restrict__ full, int64_t ldf) {
    const int i = blockIdx.x * 32 + threadIdx.x, j0 = blockIdx.y * 32;
    if (i >= n) return;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        const int j = j0 + jj;
        if (j >= n) break;
        const bool stored = upper ? i <= j : i >= j;
        T v = stored ? A[i + (int64_t)j * lda] : A[j + (int64_t)i * lda];
        if (HERM) { if (!stored) v = num<T>::conj(v); if (i == j) v = clear_imag<T>(v); }
        full[i + (int64_t)j * ldf] = v;
    }
}

template <typename T>
void symm_dev(cudaStream_t s, bool herm, char side, char uplo, int m, int n, T alpha, const T* A, int64_t lda, const T* B, int64_t ldb,
              T beta, T* C, int64_t ldc) {
    if (m <= 0 || n <= 0) return;
    if (num<T>::is_zero(alpha)) { scale_matrix<T>(s, m, n, beta, C, ldc, MASK_FULL); last_variant = VAR_SCALE_ONLY; return; }
    const bool left = (side == 'L' || side == 'l'), upper = (uplo == 'U' || uplo == 'u');
    const int na = left ? m : n;
    int64_t per16 = 16 / (int64_t)sizeof(T); if (per16 < 1) per16 = 1;
    const int64_t ldf = ((int64_t)na + per16 - 1) / per16 * per16;
    T* full = (T*)ws_alloc((size_t)ldf * na * sizeof(T));
    dim3 blk(32, 8), grd((na + 31) / 32, (na + 31) / 32);
    if (herm) expand_sym_kernel<T, true><<<grd, blk, 0, s>>>(na, A, lda, upper, full, ldf);
    else      expand_sym_kernel<T, false><<<grd, blk, 0, s>>>(na, A, lda, upper, full, ldf);
    if (left) gemm_dev<T>(s, 'N', 'N', m, n, m, alpha, full, ldf, B, ldb, beta, C, ldc, MASK_FULL);
    else      gemm_dev<T>(s, 'N', 'N', m, n, n, alpha, B, ldb, full, ldf, beta, C, ldc, MASK_FULL);
}

// C := alpha*op(A)*op(B)^T + alpha*op(B)*op(A)^T + beta*C on one triangle (trans 'N': A, B are n x k)
template <typename T>
void syr2k_dev(cudaStream_t s, char uplo, char trans, int n, int k, T alpha, const T* A, int64_t lda, const T* B, int64_t ldb, T beta,
               T* C, int64_t ldc) {
    if (n <= 0) return;
    const int mask = (uplo == 'U' || uplo == 'u') ? MASK_UPPER : MASK_LOWER;
    if (num<T>::is_zero(alpha) || k <= 0) { scale_matrix<T>(s, n, n, beta, C, ldc, mask); last_variant = VAR_SCALE_ONLY; return; }
    const bool nota = op_code(trans) == 0;
    const char t1 = nota ? 'N' : 'T', t2 = nota ? 'T' : 'N';
    gemm_dev<T>(s, t1, t2, n, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
    gemm_dev<T>(s, t1, t2, n, n, k, alpha, B, ldb, A, lda, num<T>::real(1.0), C, ldc, mask);
}

// C := alpha*op(A)*op(A)^H + beta*C, alpha and beta real (trans 'N' or 'C')
template <typename T, typename R>
void herk_dev(cudaStream_t s, char uplo, char trans, int n, int k, R alpha, const T* A, int64_t lda, R beta, T* C, int64_t ldc) {
    if (n <= 0) return;
    const int mask = (uplo == 'U' || uplo == 'u') ? MASK_UPPER : MASK_LOWER;
    const T a = num<T>::real((double)alpha), b = num<T>::real((double)beta);
    if (alpha == R(0) || k <= 0) {
        scale_matrix<T>(s, n, n, b, C, ldc, mask);
        real_diag<T>(s, n, C, ldc);
        last_variant = VAR_SCALE_ONLY;
        return;
    }
    const bool nota = op_code(trans) == 0;
    gemm_dev<T>(s, nota ? 'N' : 'C', nota ? 'C' : 'N', n, n, k, a, A, lda, A, lda, b, C, ldc, mask);
    real_diag<T>(s, n, C, ldc);
}

// C := alpha*op(A)*op(B)^H + conj(alpha)*op(B)*op(A)^H + beta*C, beta real
template <typename T, typename R>
void her2k_dev(cudaStream_t s, char uplo, char trans, int n, int k, T alpha, const T* A, int64_t lda, const T* B, int64_t ldb, R beta,
               T* C, int64_t ldc) {
    if (n <= 0) return;
    const int mask = (uplo == 'U' || uplo == 'u') ? MASK_UPPER : MASK_LOWER;
    const T b = num<T>::real((double)beta);
    if (num<T>::is_zero(alpha) || k <= 0) {
        scale_matrix<T>(s, n, n, b, C, ldc, mask);
        real_diag<T>(s, n, C, ldc);
        last_variant = VAR_SCALE_ONLY;
        return;
    }
    const bool nota = op_code(trans) == 0;
    const char t1 = nota ? 'N' : 'C', t2 = nota ? 'C' : 'N';
    T ca = alpha; ca.y = -ca.y;
    if (nota) {
        gemm_dev<T>(s, t1, t2, n, n, k, alpha, A, lda, B, ldb, b, C, ldc, mask);
        gemm_dev<T>(s, t1, t2, n, n, k, ca, B, ldb, A, lda, num<T>::real(1.0), C, ldc, mask);
    } else {   // C := alpha*A^H*B + conj(alpha)*B^H*A + beta*C
        gemm_dev<T>(s, t1, t2, n, n, k, alpha, A, lda, B, ldb, b, C, ldc, mask);
        gemm_dev<T>(s, t1, t2, n, n, k, ca, B, ldb, A, lda, num<T>::real(1.0), C, ldc, mask);
    }
    real_diag<T>(s, n, C, ldc);
}

#define B200_INST_SYM(T)                                                                                                              \
    template void symm_dev<T>(cudaStream_t, bool, char, char, int, int, T, const T*, int64_t, const T*, int64_t, T, T*, int64_t);    \
    template void syr2k_dev<T>(cudaStream_t, char, char, int, int, T, const T*, int64_t, const T*, int64_t, T, T*, int64_t);
B200_INST_SYM(float)
B200_INST_SYM(double)
B200_INST_SYM(cuFloatComplex)
B200_INST_SYM(cuDoubleComplex)
template void herk_dev<cuFloatComplex, float>(cudaStream_t, char, char, int, int, float, const cuFloatComplex*, int64_t, float, cuFloatComplex*, int64_t);
template void herk_dev<cuDoubleComplex, double>(cudaStream_t, char, char, int, int, double, const cuDoubleComplex*, int64_t, double, cuDoubleComplex*, int64_t);
template void her2k_dev<cuFloatComplex, float>(cudaStream_t, char, char, int, int, cuFloatComplex, const cuFloatComplex*, int64_t, const cuFloatComplex*, int64_t, float, cuFloatComplex*, int64_t);
template void her2k_dev<cuDoubleComplex, double>(cudaStream_t, char, char, int, int, cuDoubleComplex, const cuDoubleComplex*, int64_t, const cuDoubleComplex*, int64_t, double, cuDoubleComplex*, int64_t);

}  // namespace b200
