// potrf.cu -- lower Cholesky factorisation on the device, built from this library's own Level-3 kernels.
// Not part of the reference's exported surface: it is the diagonal-block step of the blocked Cholesky
// WORKLOAD of BASELINE.json configs[3] (DSYRK + DTRSM + DGEMM panels, SURVEY.md section 8d "C4"), which a
// LAPACK-style caller would otherwise run on the CPU between interposed BLAS calls, bouncing the block
// across PCIe every step.
//
// Recursive (cache-oblivious) formulation: A = [A11 ; A21 A22]
//     potrf(A11);  A21 := A21 * L11^-T (TRSM R,L,T,N);  A22 -= A21 * A21^T (SYRK L,N);  potrf(A22)
// so all but O(n * 64^2) of the n^3/3 flops run in the DMMA GEMM tiles; 64x64 leaves are factored by
// one CTA out of shared memory.
#include "common.cuh"
#include "kernels.h"
#include "runtime.h"
#include "abi_common.h"
#include "../../include/b200blas.h"

namespace b200 {

constexpr int PO_LEAF = 64;

// One CTA, 256 threads: right-looking unblocked Cholesky of an nb x nb (nb <= 64) block held in shared memory.
// *info (device int, 0 on entry or an earlier failure) receives the 1-based global index of the first
// non-positive pivot.
__global__ void __launch_bounds__(256) potrf_leaf_kernel(int nb, double* __restrict__ A, int64_t lda, int base, int* info) {
    __shared__ double s[PO_LEAF][PO_LEAF + 1];     // s[i][j] = A(i,j), lower triangle used
    __shared__ int bad;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;      // 16 x 16: thread owns rows tx+16a, columns ty+16b
    if (t == 0) bad = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = tx + 16 * a, j = ty + 16 * b;
            if (i < nb && j < nb && i >= j) s[i][j] = A[i + (int64_t)j * lda];
        }
    __syncthreads();
    for (int j = 0; j < nb; j++) {
        const double d = s[j][j];
        if (!(d > 0.0)) { if (t == 0) bad = j + 1; break; }      // uniform: every thread reads the same s[j][j]
        const double r = sqrt(d);
        __syncthreads();
        if (t == 0) s[j][j] = r;
        if (t > j && t < nb) s[t][j] /= r;
        __syncthreads();
        // trailing update of the lower triangle: (i, c) with j < c <= i < nb
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int c = ty + 16 * b;
            if (c <= j || c >= nb) continue;
            const double lc = s[c][j];
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const int i = tx + 16 * a;
                if (i >= c && i < nb) s[i][c] -= s[i][j] * lc;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (bad) { if (t == 0 && *info == 0) *info = base + bad; return; }
#pragma unroll
    for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = tx + 16 * a, j = ty + 16 * b;
            if (i < nb && j < nb && i >= j) A[i + (int64_t)j * lda] = s[i][j];
        }
}

constexpr int PO_IB = 128;       // inner block of the left-looking variant
constexpr int PO_BASE = 2048;    // at or below this order the left-looking variant replaces the recursion

static void potrf_rec(cudaStream_t s, int n, double* A, int64_t lda, int base, int* info);

// Left-looking blocked factorisation for a diagonal block of moderate order: per 128-wide block column ONE masked
// GEMM applies all previous columns (A[jj:, jj:jj+ib] -= A[jj:, :jj] * A[jj:jj+ib, :jj]^T, lower part only -- the
// strictly upper triangle is never written), the 128x128 diagonal block is factored by two 64-leaves, and one
// TRSM finishes the rows below.  ~13 launches per block column instead of a 2-way recursion down to 64
// (potrf(2048): 8.0 ms -> see profiles/), because at these sizes everything is launch-latency bound.
static void potrf_left(cudaStream_t s, int n, double* A, int64_t lda, int base, int* info) {
    for (int jj = 0; jj < n; jj += PO_IB) {
        const int ib = n - jj < PO_IB ? n - jj : PO_IB;
        double* Ajj = A + jj + (int64_t)jj * lda;
        if (jj > 0) dgemm_dev(s, 'N', 'T', n - jj, ib, jj, -1.0, A + jj, lda, A + jj, lda, 1.0, Ajj, lda, MASK_LOWER);
        if (ib <= PO_LEAF) {
            potrf_leaf_kernel<<<1, 256, 0, s>>>(ib, Ajj, lda, base + jj, info);
        } else {
            const int i1 = PO_LEAF, i2 = ib - i1;
            potrf_leaf_kernel<<<1, 256, 0, s>>>(i1, Ajj, lda, base + jj, info);
            trsm_dev<double>(s, 'R', 'L', 'T', 'N', i2, i1, 1.0, Ajj, lda, Ajj + i1, lda);
            syrk_dev<double>(s, 'L', 'N', i2, i1, -1.0, Ajj + i1, lda, 1.0, Ajj + i1 + (int64_t)i1 * lda, lda);
            potrf_leaf_kernel<<<1, 256, 0, s>>>(i2, Ajj + i1 + (int64_t)i1 * lda, lda, base + jj + i1, info);
        }
        const int below = n - jj - ib;
        if (below > 0) trsm_dev<double>(s, 'R', 'L', 'T', 'N', below, ib, 1.0, Ajj, lda, Ajj + ib, lda);
    }
}

static void potrf_rec(cudaStream_t s, int n, double* A, int64_t lda, int base, int* info) {
    if (n <= PO_LEAF) {
        potrf_leaf_kernel<<<1, 256, 0, s>>>(n, A, lda, base, info);
        return;
    }
    if (n <= PO_BASE) { potrf_left(s, n, A, lda, base, info); return; }
    const int n1 = ((n / 2 + PO_IB - 1) / PO_IB) * PO_IB, n2 = n - n1;
    double* A21 = A + n1;
    double* A22 = A + n1 + (int64_t)n1 * lda;
    potrf_rec(s, n1, A, lda, base, info);
    trsm_dev<double>(s, 'R', 'L', 'T', 'N', n2, n1, 1.0, A, lda, A21, lda);
    syrk_dev<double>(s, 'L', 'N', n2, n1, -1.0, A21, lda, 1.0, A22, lda);
    potrf_rec(s, n2, A22, lda, base + n1, info);
}

// info_dev: device int, must be zero on entry; a failing pivot is reported as base + its 1-based index within the block
void potrf_lower_dev(cudaStream_t s, int n, double* A, int64_t lda, int* info_dev, int base) {
    if (n <= 0) return;
    potrf_rec(s, n, A, lda, base, info_dev);
}

}  // namespace b200

using namespace b200;

extern "C" {

// Lower Cholesky factor in place (LAPACK DPOTRF 'L' semantics: the strictly upper triangle is not referenced);
// a may be host, managed or device memory.  Returns 0, or i > 0 if the leading minor of order i is not positive
// definite, or -1/-2/-4 for an illegal n / a / lda.
int b200blas_dpotrf_lower(int n, double* a, long long lda) {
    if (n < 0) return -1;
    if (lda < (n > 1 ? n : 1)) return -4;
    if (n == 0) return 0;
    if (!a) return -2;
    int result = 0;
    {
        CallScope scope;
        Operand oa(a, n, n, lda, sizeof(double), ACC_INOUT);
        int* info = (int*)((char*)device_scalar() + 64);
        cudaStream_t s = current_stream();
        B200_CUDA(cudaMemsetAsync(info, 0, sizeof(int), s));
        potrf_lower_dev(s, n, (double*)oa.dev(), oa.ld(), info);
        oa.release();
        int* pin = (int*)pinned_scalar();
        B200_CUDA(cudaMemcpyAsync(pin, info, sizeof(int), cudaMemcpyDeviceToHost, s));
        B200_CUDA(cudaStreamSynchronize(s));
        result = *pin;
        log_exec("dpotrf_lower", "n=%d lda=%lld info=%d", n, lda, result);
    }
    return result;
}

}  // extern "C"
