// gemm_generic.cuh -- type-generic register-tiled GEMM for every (type, transpose, alignment)
// combination: the small-problem / odd-type variant of the selector and the kernel the
// complex-float path and small float problems run on (large SGEMM: gemm_f32.cu).  64x64 CTA tile, 16x16 threads, 4x4 per
// thread, BK=16, operands staged through shared memory with the transpose/conjugate applied
// while staging so the inner loop is layout-free.  HBM-coalesced along whichever dimension is
// contiguous in memory.
#pragma once
#include "common.cuh"
#include "kernels.h"
#include <cuComplex.h>
#include <type_traits>

namespace b200 {

template <typename T> struct num;
template <> struct num<float> {
    static __host__ __device__ float zero() { return 0.f; }
    static __host__ __device__ float real(double r) { return (float)r; }
    static __device__ float conj(float a) { return a; }
    static __device__ float mul(float a, float b) { return a * b; }
    static __device__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ float add(float a, float b) { return a + b; }
    static __device__ float sub(float a, float b) { return a - b; }
    static __host__ __device__ bool is_zero(float a) { return a == 0.f; }
    static __host__ __device__ bool is_one(float a) { return a == 1.f; }
};
template <> struct num<double> {
    static __host__ __device__ double zero() { return 0.0; }
    static __host__ __device__ double real(double r) { return r; }
    static __device__ double conj(double a) { return a; }
    static __device__ double mul(double a, double b) { return a * b; }
    static __device__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ double add(double a, double b) { return a + b; }
    static __device__ double sub(double a, double b) { return a - b; }
    static __host__ __device__ bool is_zero(double a) { return a == 0.0; }
    static __host__ __device__ bool is_one(double a) { return a == 1.0; }
};
template <> struct num<cuFloatComplex> {
    static __host__ __device__ cuFloatComplex zero() { return make_cuFloatComplex(0.f, 0.f); }
    static __host__ __device__ cuFloatComplex real(double r) { return make_cuFloatComplex((float)r, 0.f); }
    static __device__ cuFloatComplex conj(cuFloatComplex a) { return make_cuFloatComplex(a.x, -a.y); }
    static __device__ cuFloatComplex mul(cuFloatComplex a, cuFloatComplex b) {
        return make_cuFloatComplex(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    static __device__ cuFloatComplex fma(cuFloatComplex a, cuFloatComplex b, cuFloatComplex c) {
        c.x = fmaf(a.x, b.x, c.x); c.x = fmaf(-a.y, b.y, c.x);
        c.y = fmaf(a.x, b.y, c.y); c.y = fmaf(a.y, b.x, c.y);
        return c;
    }
    static __device__ cuFloatComplex add(cuFloatComplex a, cuFloatComplex b) { return make_cuFloatComplex(a.x + b.x, a.y + b.y); }
    static __device__ cuFloatComplex sub(cuFloatComplex a, cuFloatComplex b) { return make_cuFloatComplex(a.x - b.x, a.y - b.y); }
    static __host__ __device__ bool is_zero(cuFloatComplex a) { return a.x == 0.f && a.y == 0.f; }
    static __host__ __device__ bool is_one(cuFloatComplex a) { return a.x == 1.f && a.y == 0.f; }
};
template <> struct num<cuDoubleComplex> {
    static __host__ __device__ cuDoubleComplex zero() { return make_cuDoubleComplex(0.0, 0.0); }
    static __host__ __device__ cuDoubleComplex real(double r) { return make_cuDoubleComplex(r, 0.0); }
    static __device__ cuDoubleComplex conj(cuDoubleComplex a) { return make_cuDoubleComplex(a.x, -a.y); }
    static __device__ cuDoubleComplex mul(cuDoubleComplex a, cuDoubleComplex b) {
        return make_cuDoubleComplex(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    static __device__ cuDoubleComplex fma(cuDoubleComplex a, cuDoubleComplex b, cuDoubleComplex c) {
        c.x = ::fma(a.x, b.x, c.x); c.x = ::fma(-a.y, b.y, c.x);
        c.y = ::fma(a.x, b.y, c.y); c.y = ::fma(a.y, b.x, c.y);
        return c;
    }
    static __device__ cuDoubleComplex add(cuDoubleComplex a, cuDoubleComplex b) { return make_cuDoubleComplex(a.x + b.x, a.y + b.y); }
    static __device__ cuDoubleComplex sub(cuDoubleComplex a, cuDoubleComplex b) { return make_cuDoubleComplex(a.x - b.x, a.y - b.y); }
    static __host__ __device__ bool is_zero(cuDoubleComplex a) { return a.x == 0.0 && a.y == 0.0; }
    static __host__ __device__ bool is_one(cuDoubleComplex a) { return a.x == 1.0 && a.y == 0.0; }
};

// inside/outside test for triangular outputs (SYRK-style): element (i,j) of C is written iff
// mask==FULL, or it lies in the named triangle including the diagonal.
__host__ __device__ __forceinline__ bool tri_keep(int mask, int64_t i, int64_t j) {
    return mask == MASK_FULL || (mask == MASK_LOWER ? i >= j : i <= j);
}

static inline int op_code(char t) { return (t == 'N' || t == 'n') ? 0 : ((t == 'T' || t == 't') ? 1 : 2); }

#ifdef __CUDACC__      // kernels and their launchers: nvcc only (the multi-device drivers that include this header for num<T> are
                       // also compiled by g++ against a stream simulator, tests/drivers/mgsim.cpp)
// C := beta*C over the masked region (the alpha==0 / k==0 quick path of netlib xGEMM/xSYRK, which the
// reference runs as a HOST loop over managed memory, gemm.cc:118-125 -- here it stays on the GPU).
template <typename T>
__global__ void scale_matrix_kernel(int m, int n, T beta, T* __restrict__ C, int64_t ldc, int mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int64_t j = (int64_t)blockIdx.y * blockDim.y + threadIdx.y; j < n; j += (int64_t)gridDim.y * blockDim.y) {
        if (!tri_keep(mask, i, j)) continue;
        T* p = C + i + j * ldc;
        *p = num<T>::is_zero(beta) ? num<T>::zero() : num<T>::mul(beta, *p);
    }
}
template <typename T>
void scale_matrix(cudaStream_t s, int m, int n, T beta, T* C, int64_t ldc, int mask) {
    if (m <= 0 || n <= 0 || num<T>::is_one(beta)) return;
    dim3 blk(64, 4), grd((m + 63) / 64, (n + 3) / 4 > 65535 ? 65535 : (n + 3) / 4);
    scale_matrix_kernel<T><<<grd, blk, 0, s>>>(m, n, beta, C, ldc, mask);
}

constexpr int GT_BM = 64, GT_BN = 64, GT_BK = 16;

// OPA/OPB: 0 = 'N', 1 = 'T', 2 = 'C'
template <typename T, int OPA, int OPB>
__global__ void __launch_bounds__(256) gemm_generic_kernel(int m, int n, int k, T alpha, const T* __restrict__ A,
                                                           int64_t lda, const T* __restrict__ B, int64_t ldb, T beta,
                                                           T* __restrict__ C, int64_t ldc, int mask, const int* __restrict__ gate) {
    if (gate && *gate == 0) return;                              // conditional launch (SGEMM's Inf/NaN stand-in, gemm_f32.cu)
    const int tm0 = blockIdx.x * GT_BM, tn0 = blockIdx.y * GT_BN;
    if (mask == MASK_LOWER && tm0 + GT_BM - 1 < tn0) return;   // tile entirely above the diagonal
    if (mask == MASK_UPPER && tn0 + GT_BN - 1 < tm0) return;
    __shared__ T As[GT_BK][GT_BM + 1];   // As[kk][i] = op(A)(tm0+i, k0+kk)
    __shared__ T Bs[GT_BK][GT_BN + 1];   // Bs[kk][j] = op(B)(k0+kk, tn0+j)
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;   // thread computes rows tx+16*r, cols ty+16*c
    T acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = num<T>::zero();

    for (int k0 = 0; k0 < k; k0 += GT_BK) {
        // stage A: 64x16 elements, 4 per thread, coalesced along the memory-contiguous index
#pragma unroll
        for (int e = 0; e < 4; e++) {
            int idx = tid + e * 256;
            int i, kk;
            if (OPA == 0) { i = idx % GT_BM; kk = idx / GT_BM; } else { kk = idx % GT_BK; i = idx / GT_BK; }
            T v = num<T>::zero();
            if (tm0 + i < m && k0 + kk < k) {
                v = (OPA == 0) ? A[(int64_t)(tm0 + i) + (int64_t)(k0 + kk) * lda] : A[(int64_t)(k0 + kk) + (int64_t)(tm0 + i) * lda];
                if (OPA == 2) v = num<T>::conj(v);
            }
            As[kk][i] = v;
        }
#pragma unroll
        for (int e = 0; e < 4; e++) {
            int idx = tid + e * 256;
            int j, kk;
            if (OPB == 0) { kk = idx % GT_BK; j = idx / GT_BK; } else { j = idx % GT_BN; kk = idx / GT_BN; }
            T v = num<T>::zero();
            if (tn0 + j < n && k0 + kk < k) {
                v = (OPB == 0) ? B[(int64_t)(k0 + kk) + (int64_t)(tn0 + j) * ldb] : B[(int64_t)(tn0 + j) + (int64_t)(k0 + kk) * ldb];
                if (OPB == 2) v = num<T>::conj(v);
            }
            Bs[kk][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GT_BK; kk++) {
            T a[4], b[4];
#pragma unroll
            for (int r = 0; r < 4; r++) a[r] = As[kk][tx + 16 * r];
#pragma unroll
            for (int c = 0; c < 4; c++) b[c] = Bs[kk][ty + 16 * c];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) acc[r][c] = num<T>::fma(a[r], b[c], acc[r][c]);
        }
        __syncthreads();
    }
    const bool beta0 = num<T>::is_zero(beta);
#pragma unroll
    for (int c = 0; c < 4; c++) {
        int64_t j = tn0 + ty + 16 * c;
        if (j >= n) continue;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            int64_t i = tm0 + tx + 16 * r;
            if (i >= m || !tri_keep(mask, i, j)) continue;
            T* p = C + i + j * ldc;
            T v = num<T>::mul(alpha, acc[r][c]);
            if (!beta0) v = num<T>::fma(beta, *p, v);
            *p = v;
        }
    }
}

template <typename T>
void gemm_generic_launch(cudaStream_t s, char ta, char tb, int m, int n, int k, T alpha, const T* A, int64_t lda,
                         const T* B, int64_t ldb, T beta, T* C, int64_t ldc, int mask, const int* gate = nullptr) {
    dim3 grd((m + GT_BM - 1) / GT_BM, (n + GT_BN - 1) / GT_BN), blk(256);
    int oa = op_code(ta), ob = op_code(tb);
#define B200_GG(OA, OB) \
    gemm_generic_kernel<T, OA, OB><<<grd, blk, 0, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask, gate)
    switch (oa * 3 + ob) {
        case 0: B200_GG(0, 0); break;
        case 1: B200_GG(0, 1); break;
        case 2: B200_GG(0, 2); break;
        case 3: B200_GG(1, 0); break;
        case 4: B200_GG(1, 1); break;
        case 5: B200_GG(1, 2); break;
        case 6: B200_GG(2, 0); break;
        case 7: B200_GG(2, 1); break;
        default: B200_GG(2, 2); break;
    }
#undef B200_GG
    last_variant = VAR_GENERIC_TILE;
}
#endif  // __CUDACC__

}  // namespace b200
