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
#include <cstdlib>
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


// ---------------------------------------------------------------------------------------------------------------------
// Diagonal blocks of moderate order (64 < n <= 1024) in ONE launch.  The left-looking variant below needs ~8 dependent
// kernels per 128 columns (potrf(512): 0.59 ms for 45 Mflop, pure launch-to-launch latency), and on the multi-GPU Cholesky
// this factorisation is the fixed part of every step of the critical path.  Here a small grid of co-operating CTAs runs
// the whole right-looking factorisation with 64-wide panels -- CTA 0 factors the 64x64 diagonal leaf in shared memory,
// all CTAs solve 64-row chunks of the panel below it by substitution (one row per thread in registers, like
// trsm_leaf_kernel), all CTAs apply the 64x64 tiles of the symmetric trailing update -- separated by a grid barrier on a
// global counter (three per panel).  The matrix stays in L2 (2 MB for n = 512); data written by other CTAs is read
// with ld.global.cg.  The CTAs need not be co-resident from the start: a waiting CTA only spins, it holds nothing another
// kernel needs, so CTAs of other streams finish and free their SMs.
constexpr int PC_THREADS = 256;
constexpr int PC_MAX_CTAS = 32;
constexpr int PC_LD = PO_LEAF + 1;

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ void pc_grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// one row x of X * L^T = A by substitution, in registers: x[c] = (a[c] - sum_{l<c} x[l] L(c,l)) / L(c,c), right-looking so the
// updates after each pivot are independent; sL holds L with its diagonal inverted.  (A separate function: inside the kernel's loop
// nest the compiler does not fully unroll the 64 x 64 triangle and the row falls into local memory.)
__device__ __noinline__ void pc_solve_row(const double (*sL)[PC_LD], double* row) {
    double y[PO_LEAF];
#pragma unroll
    for (int c = 0; c < PO_LEAF; c++) y[c] = row[c];
#pragma unroll
    for (int l = 0; l < PO_LEAF; l++) {
        y[l] *= sL[l][l];
#pragma unroll
        for (int c = l + 1; c < PO_LEAF; c++) y[c] = fma(-sL[c][l], y[l], y[c]);
    }
#pragma unroll
    for (int c = 0; c < PO_LEAF; c++) row[c] = y[c];
}

__global__ void __launch_bounds__(PC_THREADS) potrf_coop_kernel(int n, double* __restrict__ A, int64_t lda, int base, int* info, unsigned* counter) {
    extern __shared__ double pc_smem[];
    double (*sL)[PC_LD] = (double (*)[PC_LD])pc_smem;                       // diagonal leaf / L11:  sL[i][j]
    double (*sA)[PC_LD] = (double (*)[PC_LD])(pc_smem + PO_LEAF * PC_LD);   // a 64 x 64 block of the panel: sA[r][c]
    double (*sB)[PC_LD] = (double (*)[PC_LD])(pc_smem + 2 * PO_LEAF * PC_LD);
    __shared__ int bad;
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    unsigned bar = 0;
#pragma unroll 1
    for (int jb = 0; jb < n; jb += PO_LEAF) {
        const int ib = min(PO_LEAF, n - jb);
        double* Ajj = A + jb + (int64_t)jb * lda;
        // ---------------- phase 1: the 64 x 64 diagonal leaf, CTA 0 ----------------
        if (blockIdx.x == 0) {
            if (t == 0) bad = 0;
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const int i = tx + 16 * a, j = ty + 16 * b;
                    if (i < ib && j < ib && i >= j) sL[i][j] = ldcg(Ajj + i + (int64_t)j * lda);
                }
            __syncthreads();
            for (int j = 0; j < ib; j++) {
                const double d = sL[j][j];
                if (!(d > 0.0)) { if (t == 0) bad = j + 1; break; }      // uniform: every thread reads the same value
                const double r = sqrt(d);
                __syncthreads();
                if (t == 0) sL[j][j] = r;
                if (t > j && t < ib) sL[t][j] /= r;
                __syncthreads();
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int c = ty + 16 * b;
                    if (c <= j || c >= ib) continue;
                    const double lc = sL[c][j];
#pragma unroll
                    for (int a = 0; a < 4; a++) {
                        const int i = tx + 16 * a;
                        if (i >= c && i < ib) sL[i][c] -= sL[i][j] * lc;
                    }
                }
                __syncthreads();
            }
            __syncthreads();
            if (bad) { if (t == 0 && *info == 0) *info = base + jb + bad; }
            else {
#pragma unroll
                for (int b = 0; b < 4; b++)
#pragma unroll
                    for (int a = 0; a < 4; a++) {
                        const int i = tx + 16 * a, j = ty + 16 * b;
                        if (i < ib && j < ib && i >= j) Ajj[i + (int64_t)j * lda] = sL[i][j];
                    }
            }
        }
        pc_grid_barrier(counter, gridDim.x * ++bar);
        const int m2 = n - jb - ib;
        if (m2 <= 0) break;                                                  // uniform over the grid
        // ---------------- phase 2: panel below the leaf, X * L11^T = A21, 64-row chunks ----------------
        // sL[c][l] = L11(c, l) with the diagonal inverted; identity in the padding of a ragged last leaf
        for (int idx = t; idx < PO_LEAF * PO_LEAF; idx += PC_THREADS) {
            const int c = idx % PO_LEAF, l = idx / PO_LEAF;
            double v = 0.0;
            if (c < ib && l < ib) { if (c >= l) v = ldcg(Ajj + c + (int64_t)l * lda); }
            if (c == l) v = (c < ib) ? 1.0 / v : 1.0;
            sL[c][l] = v;
        }
        double* A21 = Ajj + ib;                                              // m2 x ib, rows contiguous in memory
#pragma unroll 1
        for (int r0 = blockIdx.x * PO_LEAF; r0 < m2; r0 += gridDim.x * PO_LEAF) {
            const int nr = min(PO_LEAF, m2 - r0);
            __syncthreads();
            for (int idx = t; idx < PO_LEAF * PO_LEAF; idx += PC_THREADS) {
                const int r = idx % PO_LEAF, c = idx / PO_LEAF;
                sA[r][c] = (r < nr && c < ib) ? ldcg(A21 + r0 + r + (int64_t)c * lda) : 0.0;
            }
            __syncthreads();
            if (t < PO_LEAF) pc_solve_row(sL, sA[t]);
            __syncthreads();
            for (int idx = t; idx < PO_LEAF * PO_LEAF; idx += PC_THREADS) {
                const int r = idx % PO_LEAF, c = idx / PO_LEAF;
                if (r < nr && c < ib) A21[r0 + r + (int64_t)c * lda] = sA[r][c];
            }
        }
        pc_grid_barrier(counter, gridDim.x * ++bar);
        // ---------------- phase 3: trailing update A22 -= A21 * A21^T on the lower triangle, 64 x 64 tiles ----------------
        const int nt = (m2 + PO_LEAF - 1) / PO_LEAF, ntask = nt * (nt + 1) / 2;
        double* A22 = Ajj + ib + (int64_t)ib * lda;
#pragma unroll 1
        for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
            // task -> (ti >= tj): row-major enumeration of the lower triangle of tiles
            int ti = (int)((sqrtf(8.0f * task + 1.0f) - 1.0f) * 0.5f);
            while (ti * (ti + 1) / 2 > task) ti--;
            while ((ti + 1) * (ti + 2) / 2 <= task) ti++;
            const int tj = task - ti * (ti + 1) / 2;
            const int ri0 = ti * PO_LEAF, rj0 = tj * PO_LEAF;
            __syncthreads();
            for (int idx = t; idx < PO_LEAF * PO_LEAF; idx += PC_THREADS) {
                const int r = idx % PO_LEAF, c = idx / PO_LEAF;
                sA[r][c] = (ri0 + r < m2 && c < ib) ? ldcg(A21 + ri0 + r + (int64_t)c * lda) : 0.0;
                sB[r][c] = (rj0 + r < m2 && c < ib) ? ldcg(A21 + rj0 + r + (int64_t)c * lda) : 0.0;
            }
            __syncthreads();
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 8
            for (int kk = 0; kk < PO_LEAF; kk++) {
                double av[4], bv[4];
#pragma unroll
                for (int a = 0; a < 4; a++) av[a] = sA[tx + 16 * a][kk];
#pragma unroll
                for (int b = 0; b < 4; b++) bv[b] = sB[ty + 16 * b][kk];
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
            }
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const int i = ri0 + tx + 16 * a, j = rj0 + ty + 16 * b;
                    if (i < m2 && j < m2 && i >= j) {
                        double* cp = A22 + i + (int64_t)j * lda;
                        *cp = ldcg(cp) - acc[a][b];
                    }
                }
        }
        pc_grid_barrier(counter, gridDim.x * ++bar);
    }
}

static void potrf_coop(cudaStream_t s, int n, double* A, int64_t lda, int base, int* info) {
    const int nt0 = (n - PO_LEAF + PO_LEAF - 1) / PO_LEAF;
    int grid = nt0 * (nt0 + 1) / 2;
    if (grid < 1) grid = 1;
    if (grid > PC_MAX_CTAS) grid = PC_MAX_CTAS;
    const int smem = 3 * PO_LEAF * PC_LD * (int)sizeof(double);
    unsigned* counter = (unsigned*)ws_alloc(256);
    B200_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), s));
    set_max_dynamic_smem((const void*)potrf_coop_kernel, smem);
    potrf_coop_kernel<<<grid, PC_THREADS, smem, s>>>(n, A, lda, base, info, counter);
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
    static const int coop_max = getenv("B200BLAS_POTRF_COOP") ? atoi(getenv("B200BLAS_POTRF_COOP")) : 1024;
    if (n <= coop_max) { potrf_coop(s, n, A, lda, base, info); return; }
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
