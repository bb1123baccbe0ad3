// level2.cu -- GEMV and TRSV (reference blas_level2/gemv.cc:11-118, trsv.cc:11-114 are dead wrappers
// forwarding to cublas<t>gemv / cublas<t>trsv).
//
// DGEMV is HBM-bound: 8 bytes of A per 2 flops.  Both kernels stream A exactly once with 128-bit
// coalesced loads along the memory-contiguous (row) index, keep 8+ independent loads in flight per
// thread, never stage A through shared memory (no reuse), and produce deterministic results:
//   'N': CTA = 256 rows x a chunk of columns; per-chunk partial vectors are combined in chunk order
//        by a finishing kernel that also applies alpha/beta.
//   'T': CTA = 4 columns x all rows; x is read once per 4 columns (L2-resident), block tree reduce.
#include "common.cuh"
#include <cstring>
#include "kernels.h"
#include "gemm_generic.cuh"
#include "runtime.h"

namespace b200 {

// the fast DGEMV path is taken only when T is double; this keeps the other instantiations free of type punning
template <typename T> static inline double scalar_as_f64(const T& v) {
    double d = 0;
    if (sizeof(T) == sizeof(double)) memcpy(&d, &v, sizeof d);
    return d;
}


__device__ __forceinline__ double2 ldg_stream2(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int64_t vix2(int64_t i, int64_t n, int64_t inc) { return inc >= 0 ? i * inc : (n - 1 - i) * (-inc); }

// ------------------------------------------------------------------ DGEMV 'N', vector path
constexpr int GN_THREADS = 128;            // x 2 rows (double2) = 256 rows per CTA
__global__ void __launch_bounds__(GN_THREADS) dgemv_n_vec_kernel(int m, int n, const double* __restrict__ A, int64_t lda,
                                                                const double* __restrict__ x, int64_t incx, int cols_per_chunk,
                                                                double* __restrict__ part /* [chunks][mpad] */, int64_t mpad) {
    const int row = (blockIdx.x * GN_THREADS + threadIdx.x) * 2;
    const int c0 = blockIdx.y * cols_per_chunk, c1 = min(n, c0 + cols_per_chunk);
    if (row >= m) return;
    const bool pair = row + 1 < m;
    double2 acc0 = {0, 0}, acc1 = {0, 0}, acc2 = {0, 0}, acc3 = {0, 0};
    const double* a = A + row + (int64_t)c0 * lda;
    int j = c0;
    if (pair) {
        for (; j + 7 < c1; j += 8) {
            double2 v0 = ldg_stream2((const double2*)(a)), v1 = ldg_stream2((const double2*)(a + lda));
            double2 v2 = ldg_stream2((const double2*)(a + 2 * lda)), v3 = ldg_stream2((const double2*)(a + 3 * lda));
            double2 v4 = ldg_stream2((const double2*)(a + 4 * lda)), v5 = ldg_stream2((const double2*)(a + 5 * lda));
            double2 v6 = ldg_stream2((const double2*)(a + 6 * lda)), v7 = ldg_stream2((const double2*)(a + 7 * lda));
            const double x0 = __ldg(x + vix2(j, n, incx)), x1 = __ldg(x + vix2(j + 1, n, incx)), x2 = __ldg(x + vix2(j + 2, n, incx)),
                         x3 = __ldg(x + vix2(j + 3, n, incx)), x4 = __ldg(x + vix2(j + 4, n, incx)), x5 = __ldg(x + vix2(j + 5, n, incx)),
                         x6 = __ldg(x + vix2(j + 6, n, incx)), x7 = __ldg(x + vix2(j + 7, n, incx));
            acc0.x = fma(v0.x, x0, acc0.x); acc0.y = fma(v0.y, x0, acc0.y); acc1.x = fma(v1.x, x1, acc1.x); acc1.y = fma(v1.y, x1, acc1.y);
            acc2.x = fma(v2.x, x2, acc2.x); acc2.y = fma(v2.y, x2, acc2.y); acc3.x = fma(v3.x, x3, acc3.x); acc3.y = fma(v3.y, x3, acc3.y);
            acc0.x = fma(v4.x, x4, acc0.x); acc0.y = fma(v4.y, x4, acc0.y); acc1.x = fma(v5.x, x5, acc1.x); acc1.y = fma(v5.y, x5, acc1.y);
            acc2.x = fma(v6.x, x6, acc2.x); acc2.y = fma(v6.y, x6, acc2.y); acc3.x = fma(v7.x, x7, acc3.x); acc3.y = fma(v7.y, x7, acc3.y);
            a += 8 * lda;
        }
        for (; j < c1; j++) {
            double2 v = *(const double2*)a;
            const double xv = __ldg(x + vix2(j, n, incx));
            acc0.x = fma(v.x, xv, acc0.x); acc0.y = fma(v.y, xv, acc0.y);
            a += lda;
        }
    } else {
        for (; j < c1; j++) { acc0.x = fma(*a, __ldg(x + vix2(j, n, incx)), acc0.x); a += lda; }
    }
    double* p = part + (int64_t)blockIdx.y * mpad + row;
    p[0] = (acc0.x + acc1.x) + (acc2.x + acc3.x);
    if (pair) p[1] = (acc0.y + acc1.y) + (acc2.y + acc3.y);
}
// y := alpha * sum_chunks part + beta*y
template <typename T>
__global__ void gemv_finish_kernel(int len, int chunks, const T* __restrict__ part, int64_t mpad, T alpha, T beta, T* __restrict__ y,
                                   int64_t incy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    T s = num<T>::zero();
    for (int c = 0; c < chunks; c++) s = num<T>::add(s, part[(int64_t)c * mpad + i]);
    T* p = y + vix2(i, len, incy);
    T v = num<T>::mul(alpha, s);
    if (!num<T>::is_zero(beta)) v = num<T>::fma(beta, *p, v);
    *p = v;
}

// ------------------------------------------------------------------ DGEMV 'T', vector path
constexpr int GT_THREADS = 256, GT_COLS = 4;
__global__ void __launch_bounds__(GT_THREADS) dgemv_t_vec_kernel(int m, int n, double alpha, const double* __restrict__ A, int64_t lda,
                                                                const double* __restrict__ x, double beta, double* __restrict__ y,
                                                                int64_t incy) {
    __shared__ double sm[GT_COLS][GT_THREADS / 32];
    const int j0 = blockIdx.x * GT_COLS;
    const int nc = min(GT_COLS, n - j0);
    double acc[GT_COLS] = {0, 0, 0, 0};
    const double* a[GT_COLS];
#pragma unroll
    for (int c = 0; c < GT_COLS; c++) a[c] = A + (int64_t)(j0 + (c < nc ? c : 0)) * lda;   // clamp: duplicates are discarded
    const int m2 = m >> 1;
    int i = threadIdx.x;
    for (; i + GT_THREADS < m2; i += 2 * GT_THREADS) {
        double2 xa = __ldg((const double2*)x + i), xb = __ldg((const double2*)x + i + GT_THREADS);
        double2 va[GT_COLS], vb[GT_COLS];
#pragma unroll
        for (int c = 0; c < GT_COLS; c++) { va[c] = ldg_stream2((const double2*)a[c] + i); vb[c] = ldg_stream2((const double2*)a[c] + i + GT_THREADS); }
#pragma unroll
        for (int c = 0; c < GT_COLS; c++) {
            acc[c] = fma(va[c].x, xa.x, acc[c]); acc[c] = fma(va[c].y, xa.y, acc[c]);
            acc[c] = fma(vb[c].x, xb.x, acc[c]); acc[c] = fma(vb[c].y, xb.y, acc[c]);
        }
    }
    for (; i < m2; i += GT_THREADS) {
        double2 xa = __ldg((const double2*)x + i);
#pragma unroll
        for (int c = 0; c < GT_COLS; c++) { double2 v = ldg_stream2((const double2*)a[c] + i); acc[c] = fma(v.x, xa.x, acc[c]); acc[c] = fma(v.y, xa.y, acc[c]); }
    }
    if ((m & 1) && threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < GT_COLS; c++) acc[c] = fma(a[c][m - 1], x[m - 1], acc[c]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < GT_COLS; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], o);
        if (lane == 0) sm[c][warp] = acc[c];
    }
    __syncthreads();
    if (threadIdx.x < nc) {
        double s = 0;
#pragma unroll
        for (int w = 0; w < GT_THREADS / 32; w++) s += sm[threadIdx.x][w];
        double* p = y + vix2(j0 + threadIdx.x, n, incy);
        *p = (beta == 0.0) ? alpha * s : fma(beta, *p, alpha * s);
    }
}

// ------------------------------------------------------------------ generic paths (all types, any stride/alignment)
// OPC: conjugate A (for 'C').  'N': one thread per row over a chunk of columns.
template <typename T, bool CONJ>
__global__ void __launch_bounds__(128) gemv_n_generic_kernel(int m, int n, const T* __restrict__ A, int64_t lda, const T* __restrict__ x,
                                                            int64_t incx, int cols_per_chunk, T* __restrict__ part, int64_t mpad) {
    const int row = blockIdx.x * 128 + threadIdx.x;
    const int c0 = blockIdx.y * cols_per_chunk, c1 = min(n, c0 + cols_per_chunk);
    if (row >= m) return;
    T acc = num<T>::zero();
    for (int j = c0; j < c1; j++) {
        T av = A[row + (int64_t)j * lda];
        if (CONJ) av = num<T>::conj(av);
        acc = num<T>::fma(av, x[vix2(j, n, incx)], acc);
    }
    part[(int64_t)blockIdx.y * mpad + row] = acc;
}
// 'T'/'C': one CTA (128 threads) per column
template <typename T, bool CONJ>
__global__ void __launch_bounds__(128) gemv_t_generic_kernel(int m, int n, T alpha, const T* __restrict__ A, int64_t lda,
                                                            const T* __restrict__ x, int64_t incx, T beta, T* __restrict__ y, int64_t incy) {
    __shared__ T sm[128];
    const int j = blockIdx.x;
    T acc = num<T>::zero();
    for (int i = threadIdx.x; i < m; i += 128) {
        T av = A[i + (int64_t)j * lda];
        if (CONJ) av = num<T>::conj(av);
        acc = num<T>::fma(av, x[vix2(i, m, incx)], acc);
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] = num<T>::add(sm[threadIdx.x], sm[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        T* p = y + vix2(j, n, incy);
        T v = num<T>::mul(alpha, sm[0]);
        if (!num<T>::is_zero(beta)) v = num<T>::fma(beta, *p, v);
        *p = v;
    }
}

template <typename T> __global__ void scale_vec_kernel(int len, T beta, T* y, int64_t incy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    T* p = y + vix2(i, len, incy);
    *p = num<T>::is_zero(beta) ? num<T>::zero() : num<T>::mul(beta, *p);
}

static int pick_chunks(int row_blocks, int n) {
    const int target = (sm_count() > 0 ? sm_count() : 148) * 12;
    int chunks = (target + row_blocks - 1) / row_blocks;
    if (chunks < 1) chunks = 1;
    int maxc = (n + 63) / 64;            // at least 64 columns per chunk
    if (chunks > maxc) chunks = maxc;
    if (chunks < 1) chunks = 1;
    return chunks;
}

template <typename T>
void gemv_dev(cudaStream_t s, char trans, int m, int n, T alpha, const T* A, int64_t lda, const T* x, int64_t incx, T beta, T* y,
              int64_t incy) {
    const int op = op_code(trans);
    const int leny = op == 0 ? m : n;
    if (m <= 0 || n <= 0) return;
    if (num<T>::is_zero(alpha)) {
        if (!num<T>::is_one(beta)) scale_vec_kernel<T><<<(leny + 255) / 256, 256, 0, s>>>(leny, beta, y, incy);
        last_variant = VAR_SCALE_ONLY;
        return;
    }
    constexpr bool is_f64 = sizeof(T) == 8 && !std::is_same<T, cuFloatComplex>::value;
    const bool vec = is_f64 && ((uintptr_t)A % 16 == 0) && (lda % 2 == 0);
    if (op == 0) {
        const int rows_per_cta = vec ? GN_THREADS * 2 : 128;
        const int row_blocks = (m + rows_per_cta - 1) / rows_per_cta;
        const int chunks = pick_chunks(row_blocks, n);
        const int cpc = ((n + chunks - 1) / chunks + 7) / 8 * 8;
        const int nchunks = (n + cpc - 1) / cpc;
        const int64_t mpad = ((int64_t)m + 31) / 32 * 32;
        T* part = (T*)ws_alloc((size_t)nchunks * mpad * sizeof(T));
        dim3 grd(row_blocks, nchunks);
        if (vec) dgemv_n_vec_kernel<<<grd, GN_THREADS, 0, s>>>(m, n, (const double*)A, lda, (const double*)x, incx, cpc, (double*)part, mpad);
        else gemv_n_generic_kernel<T, false><<<grd, 128, 0, s>>>(m, n, A, lda, x, incx, cpc, part, mpad);
        gemv_finish_kernel<T><<<(m + 255) / 256, 256, 0, s>>>(m, nchunks, part, mpad, alpha, beta, y, incy);
    } else {
        const bool vecx = vec && incx == 1 && ((uintptr_t)x % 16 == 0) && (op == 1 || std::is_same<T, double>::value);
        if (vecx) dgemv_t_vec_kernel<<<(n + GT_COLS - 1) / GT_COLS, GT_THREADS, 0, s>>>(m, n, scalar_as_f64(alpha), (const double*)A, lda,
                                                                                       (const double*)x, scalar_as_f64(beta), (double*)y, incy);
        else if (op == 2) gemv_t_generic_kernel<T, true><<<n, 128, 0, s>>>(m, n, alpha, A, lda, x, incx, beta, y, incy);
        else gemv_t_generic_kernel<T, false><<<n, 128, 0, s>>>(m, n, alpha, A, lda, x, incx, beta, y, incy);
    }
    last_variant = VAR_GENERIC_TILE;
}
template void gemv_dev<float>(cudaStream_t, char, int, int, float, const float*, int64_t, const float*, int64_t, float, float*, int64_t);
template void gemv_dev<double>(cudaStream_t, char, int, int, double, const double*, int64_t, const double*, int64_t, double, double*, int64_t);
template void gemv_dev<cuFloatComplex>(cudaStream_t, char, int, int, cuFloatComplex, const cuFloatComplex*, int64_t, const cuFloatComplex*, int64_t, cuFloatComplex, cuFloatComplex*, int64_t);
template void gemv_dev<cuDoubleComplex>(cudaStream_t, char, int, int, cuDoubleComplex, const cuDoubleComplex*, int64_t, const cuDoubleComplex*, int64_t, cuDoubleComplex, cuDoubleComplex*, int64_t);

// ------------------------------------------------------------------ TRSV
// Blocked: 64x64 diagonal blocks are solved by one CTA out of shared memory (exact substitution, same
// operation order as netlib), the rest of x is updated with the GEMV kernels above.
template <typename T> struct TvNb { static constexpr int value = sizeof(T) > 8 ? 32 : 64; };   // 48 KiB static smem cap
template <typename T> __device__ __forceinline__ T num_div(T a, T b);
template <> __device__ __forceinline__ float num_div(float a, float b) { return a / b; }
template <> __device__ __forceinline__ double num_div(double a, double b) { return a / b; }
template <> __device__ __forceinline__ cuFloatComplex num_div(cuFloatComplex a, cuFloatComplex b) { return cuCdivf(a, b); }
template <> __device__ __forceinline__ cuDoubleComplex num_div(cuDoubleComplex a, cuDoubleComplex b) { return cuCdiv(a, b); }

template <typename T> __device__ __forceinline__ T plus_one_dev() { return num<T>::real(1.0); }
// solves op(Ablk) x = x for one nb x nb block; opupper: op(A) is upper triangular (back substitution)
template <typename T>
__global__ void __launch_bounds__(TvNb<T>::value) trsv_diag_kernel(int nb, const T* __restrict__ A, int64_t lda, int op, bool opupper, bool unit, T* x) {
    constexpr int TV_NB = TvNb<T>::value;
    __shared__ T sA[TV_NB][TV_NB + 1];   // sA[i][l] = op(A)(i,l)
    __shared__ T sx[TV_NB];
    const int t = threadIdx.x;
    for (int l = 0; l < nb; l++) {
        if (t < nb) {
            // element op(A)(t? ...): load column l of the stored block coalesced, scatter by op
            T v = A[t + (int64_t)l * lda];               // stored A(t,l)
            if (op == 2) v = num<T>::conj(v);
            if (op == 0) sA[t][l] = v; else sA[l][t] = v;
        }
    }
    if (t < nb) sx[t] = x[t];
    __syncthreads();
    // pivot reciprocals, all at once: a division by one thread inside the loop below put an FP64 division sequence (hundreds of
    // cycles) on every step of the dependency chain (x * (1/d) instead of x / d: within the solves' stated tolerance)
    __shared__ T sr[TV_NB];
    if (t < nb && !unit) sr[t] = num_div<T>(plus_one_dev<T>(), sA[t][t]);
    __syncthreads();
    for (int step = 0; step < nb; step++) {
        const int j = opupper ? nb - 1 - step : step;
        if (t == j && !unit) sx[j] = num<T>::mul(sx[j], sr[j]);
        __syncthreads();
        const bool mine = opupper ? (t < j) : (t > j && t < nb);
        if (mine) sx[t] = num<T>::sub(sx[t], num<T>::mul(sA[t][j], sx[j]));
        __syncthreads();
    }
    if (t < nb) x[t] = sx[t];
}

template <typename T> static T minus_one();
template <> float minus_one<float>() { return -1.f; }
template <> double minus_one<double>() { return -1.0; }
template <> cuFloatComplex minus_one<cuFloatComplex>() { return make_cuFloatComplex(-1.f, 0.f); }
template <> cuDoubleComplex minus_one<cuDoubleComplex>() { return make_cuDoubleComplex(-1.0, 0.0); }
template <typename T> static T plus_one();
template <> float plus_one<float>() { return 1.f; }
template <> double plus_one<double>() { return 1.0; }
template <> cuFloatComplex plus_one<cuFloatComplex>() { return make_cuFloatComplex(1.f, 0.f); }
template <> cuDoubleComplex plus_one<cuDoubleComplex>() { return make_cuDoubleComplex(1.0, 0.0); }

template <typename T>
void trsv_dev(cudaStream_t s, char uplo, char trans, char diag, int n, const T* A, int64_t lda, T* x, int64_t incx) {
    if (n <= 0) return;
    const int op = op_code(trans);
    const bool upper = (uplo == 'U' || uplo == 'u'), unit = (diag == 'U' || diag == 'u');
    const bool opupper = op == 0 ? upper : !upper;
    T* xc = x;
    if (incx != 1) {   // work on a contiguous copy
        xc = (T*)ws_alloc((size_t)n * sizeof(T));
        copy_dev<T>(s, n, x, incx, xc, 1);
    }
    constexpr int TV_NB = TvNb<T>::value;
    const int nblk = (n + TV_NB - 1) / TV_NB;
    const T m1 = minus_one<T>(), p1 = plus_one<T>();
    for (int step = 0; step < nblk; step++) {
        const int b = opupper ? nblk - 1 - step : step;
        const int b0 = b * TV_NB, nb = min(TV_NB, n - b0), b1 = b0 + nb;
        trsv_diag_kernel<T><<<1, TV_NB, 0, s>>>(nb, A + b0 + (int64_t)b0 * lda, lda, op, opupper, unit, xc + b0);
        // update the not-yet-solved part of x with the solved block
        if (op == 0) {
            if (!opupper) { if (b1 < n) gemv_dev<T>(s, 'N', n - b1, nb, m1, A + b1 + (int64_t)b0 * lda, lda, xc + b0, 1, p1, xc + b1, 1); }
            else          { if (b0 > 0) gemv_dev<T>(s, 'N', b0, nb, m1, A + (int64_t)b0 * lda, lda, xc + b0, 1, p1, xc, 1); }
        } else {
            // op(A) = A^T/A^H: rows of the stored A are columns of op(A)
            if (!opupper) { if (b1 < n) gemv_dev<T>(s, trans, nb, n - b1, m1, A + b0 + (int64_t)b1 * lda, lda, xc + b0, 1, p1, xc + b1, 1); }
            else          { if (b0 > 0) gemv_dev<T>(s, trans, nb, b0, m1, A + b0, lda, xc + b0, 1, p1, xc, 1); }
        }
    }
    if (incx != 1) copy_dev<T>(s, n, xc, 1, x, incx);
    last_variant = VAR_GENERIC_TILE;
}
template void trsv_dev<float>(cudaStream_t, char, char, char, int, const float*, int64_t, float*, int64_t);
template void trsv_dev<double>(cudaStream_t, char, char, char, int, const double*, int64_t, double*, int64_t);
template void trsv_dev<cuFloatComplex>(cudaStream_t, char, char, char, int, const cuFloatComplex*, int64_t, cuFloatComplex*, int64_t);
template void trsv_dev<cuDoubleComplex>(cudaStream_t, char, char, char, int, const cuDoubleComplex*, int64_t, cuDoubleComplex*, int64_t);

}  // namespace b200
