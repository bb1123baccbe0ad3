// level3_blocked.cu -- SYRK, TRMM and TRSM as blocked recursions onto the GEMM tile engine
// (reference blas_level3/syrk.cc:43-76, trmm.cc:42-79, trsm.cc:40-73 forward to cublas<t>syrk /
// trmm / trsm; north_star (1): "SYRK and TRSM are built as blocked recursions onto the same GEMM
// tiles").
//
//  SYRK  one masked GEMM launch: tiles entirely outside the referenced triangle exit at once,
//        diagonal tiles store only the referenced half; nothing outside the triangle is written.
//  TRMM  recursive 2x2 splitting in place; off-diagonal blocks are GEMMs; a diagonal block of
//        <= 256 is copied to scratch with the unreferenced triangle zeroed (and the unit diagonal
//        made explicit), multiplied out of place by the GEMM kernel and copied back.
//  TRSM  same recursion; diagonal blocks of <= 64 (32 complex) are solved in place by substitution, one
//        right-hand side per thread in registers (trsm_leaf_kernel); every update is a GEMM.
#include "common.cuh"
#include "kernels.h"
#include "gemm_generic.cuh"
#include "runtime.h"
#include <cstdlib>

namespace b200 {

template <> void gemm_dev<float>(cudaStream_t s, char ta, char tb, int m, int n, int k, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mask) { sgemm_dev(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask); }
template <> void gemm_dev<double>(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int mask) { dgemm_dev(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask); }
template <> void gemm_dev<cuFloatComplex>(cudaStream_t s, char ta, char tb, int m, int n, int k, cuFloatComplex alpha, const cuFloatComplex* A, int64_t lda, const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C, int64_t ldc, int mask) { cgemm_dev(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask); }
template <> void gemm_dev<cuDoubleComplex>(cudaStream_t s, char ta, char tb, int m, int n, int k, cuDoubleComplex alpha, const cuDoubleComplex* A, int64_t lda, const cuDoubleComplex* B, int64_t ldb, cuDoubleComplex beta, cuDoubleComplex* C, int64_t ldc, int mask) { zgemm_dev(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask); }

// ------------------------------------------------------------------ SYRK
template <typename T>
void syrk_dev(cudaStream_t s, char uplo, char trans, int n, int k, T alpha, const T* A, int64_t lda, T beta, T* C, int64_t ldc) {
    const int mask = (uplo == 'U' || uplo == 'u') ? MASK_UPPER : MASK_LOWER;
    const bool nota = op_code(trans) == 0;
    // netlib xSYRK: trans 'T' and (real types) 'C' both mean A^T A -- never a conjugate
    gemm_dev<T>(s, nota ? 'N' : 'T', nota ? 'T' : 'N', n, n, k, alpha, A, lda, A, lda, beta, C, ldc, mask);
}
template void syrk_dev<float>(cudaStream_t, char, char, int, int, float, const float*, int64_t, float, float*, int64_t);
template void syrk_dev<double>(cudaStream_t, char, char, int, int, double, const double*, int64_t, double, double*, int64_t);
template void syrk_dev<cuFloatComplex>(cudaStream_t, char, char, int, int, cuFloatComplex, const cuFloatComplex*, int64_t, cuFloatComplex, cuFloatComplex*, int64_t);
template void syrk_dev<cuDoubleComplex>(cudaStream_t, char, char, int, int, cuDoubleComplex, const cuDoubleComplex*, int64_t, cuDoubleComplex, cuDoubleComplex*, int64_t);

// ------------------------------------------------------------------ helpers
// dst (nb x nb, ld = ldd) := the referenced triangle of src with zeros elsewhere; unit => diagonal = 1
template <typename T>
__global__ void tri_copy_kernel(int nb, const T* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd, bool upper, bool unit) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= nb || j >= nb) return;
    T v = num<T>::zero();
    if (i == j) v = unit ? num<T>::real(1.0) : src[i + (int64_t)j * lds];
    else if (upper ? i < j : i > j) v = src[i + (int64_t)j * lds];
    dst[i + (int64_t)j * ldd] = v;
}
template <typename T>
__global__ void copy_matrix_kernel(int m, int n, const T* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    for (int64_t j = blockIdx.y; j < n; j += gridDim.y) dst[i + j * ldd] = src[i + j * lds];
}
template <typename T> static void copy_matrix(cudaStream_t s, int m, int n, const T* src, int64_t lds, T* dst, int64_t ldd) {
    if (m <= 0 || n <= 0) return;
    dim3 grd((m + 255) / 256, n > 65535 ? 65535 : n);
    copy_matrix_kernel<T><<<grd, 256, 0, s>>>(m, n, src, lds, dst, ldd);
}
static inline int split_point(int n, int align) {
    int h = ((n / 2 + align - 1) / align) * align;
    if (h >= n) h = ((n - 1) / align) * align;
    return h;
}
static inline int64_t even_ld(int64_t rows, size_t elem) { int64_t per = 16 / (int64_t)elem; if (per < 1) per = 1; return (rows + per - 1) / per * per; }

// ------------------------------------------------------------------ TRMM
constexpr int TRMM_BASE = 256;

template <typename T> struct TrCtx {
    cudaStream_t s; bool left, upper /*stored triangle*/, unit; char trans; const T* A; int64_t lda; T* B; int64_t ldb;
    T* triw; T* outw;   // scratch: TRMM_BASE^2 triangle copy, and the out-of-place product
    int64_t ldo;
};

// sub-block accessors: A(i0.., j0..) of the stored triangular matrix
template <typename T> static inline const T* Aat(const TrCtx<T>& c, int i, int j) { return c.A + i + (int64_t)j * c.lda; }

// B[r0:r0+mr, c0:c0+nc] := alpha * op(T[d0:d0+nd]) * B_blk   (left: rows d0.., nd = mr)   or   B_blk * op(T) (right: cols d0.., nd = nc)
template <typename T>
static void trmm_rec(const TrCtx<T>& c, T alpha, int d0, int nd, int other /* n for left, m for right */) {
    const bool opupper = (op_code(c.trans) == 0) ? c.upper : !c.upper;
    if (nd <= TRMM_BASE) {
        dim3 blk(32, 8), grd((nd + 31) / 32, (nd + 7) / 8);
        const int64_t ldt = even_ld(nd, sizeof(T));
        tri_copy_kernel<T><<<grd, blk, 0, c.s>>>(nd, Aat(c, d0, d0), c.lda, c.triw, ldt, c.upper, c.unit);
        if (c.left) {
            T* Bb = c.B + d0;
            gemm_dev<T>(c.s, c.trans, 'N', nd, other, nd, alpha, c.triw, ldt, Bb, c.ldb, num<T>::zero(), c.outw, c.ldo);
            copy_matrix<T>(c.s, nd, other, c.outw, c.ldo, Bb, c.ldb);
        } else {
            T* Bb = c.B + (int64_t)d0 * c.ldb;
            gemm_dev<T>(c.s, 'N', c.trans, other, nd, nd, alpha, Bb, c.ldb, c.triw, ldt, num<T>::zero(), c.outw, c.ldo);
            copy_matrix<T>(c.s, other, nd, c.outw, c.ldo, Bb, c.ldb);
        }
        return;
    }
    const int n1 = split_point(nd, 128), n2 = nd - n1;
    const int a = d0, b = d0 + n1;          // block 1 = [a, b), block 2 = [b, d0+nd)
    const T one = num<T>::real(1.0);
    // off-diagonal block of the STORED matrix: lower -> A21 = A(b.., a..) (n2 x n1); upper -> A12 = A(a.., b..) (n1 x n2)
    const T* Aoff = c.upper ? Aat(c, a, b) : Aat(c, b, a);
    const bool notr = op_code(c.trans) == 0;
    if (c.left) {
        T* B1 = c.B + a; T* B2 = c.B + b;
        if (!opupper) {   // op(A) lower: B2 = T22 B2 + T21 B1 ; B1 = T11 B1   (B2 first: it needs the old B1)
            trmm_rec(c, alpha, b, n2, other);
            // T21 = op(A)(b.., a..): stored lower & 'N' -> A21 ; stored upper & 'T' -> A12^T
            gemm_dev<T>(c.s, notr ? 'N' : c.trans, 'N', n2, other, n1, alpha, Aoff, c.lda, B1, c.ldb, one, B2, c.ldb);
            trmm_rec(c, alpha, a, n1, other);
        } else {          // op(A) upper: B1 = T11 B1 + T12 B2 ; B2 = T22 B2
            trmm_rec(c, alpha, a, n1, other);
            gemm_dev<T>(c.s, notr ? 'N' : c.trans, 'N', n1, other, n2, alpha, Aoff, c.lda, B2, c.ldb, one, B1, c.ldb);
            trmm_rec(c, alpha, b, n2, other);
        }
    } else {
        T* B1 = c.B + (int64_t)a * c.ldb; T* B2 = c.B + (int64_t)b * c.ldb;
        if (!opupper) {   // op(A) lower: B1 = B1 T11 + B2 T21 ; B2 = B2 T22   (B1 first)
            trmm_rec(c, alpha, a, n1, other);
            gemm_dev<T>(c.s, 'N', notr ? 'N' : c.trans, other, n1, n2, alpha, B2, c.ldb, Aoff, c.lda, one, B1, c.ldb);
            trmm_rec(c, alpha, b, n2, other);
        } else {          // op(A) upper: B2 = B1 T12 + B2 T22 ; B1 = B1 T11   (B2 first)
            trmm_rec(c, alpha, b, n2, other);
            gemm_dev<T>(c.s, 'N', notr ? 'N' : c.trans, other, n2, n1, alpha, B1, c.ldb, Aoff, c.lda, one, B2, c.ldb);
            trmm_rec(c, alpha, a, n1, other);
        }
    }
}

template <typename T>
void trmm_dev(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* A, int64_t lda, T* B,
              int64_t ldb) {
    if (m <= 0 || n <= 0) return;
    if (num<T>::is_zero(alpha)) { scale_matrix<T>(s, m, n, num<T>::zero(), B, ldb, MASK_FULL); last_variant = VAR_SCALE_ONLY; return; }
    TrCtx<T> c;
    c.s = s; c.left = (side == 'L' || side == 'l'); c.upper = (uplo == 'U' || uplo == 'u'); c.unit = (diag == 'U' || diag == 'u');
    c.trans = op_code(trans) == 0 ? 'N' : (op_code(trans) == 1 ? 'T' : 'C');
    c.A = A; c.lda = lda; c.B = B; c.ldb = ldb;
    const int nd = c.left ? m : n, other = c.left ? n : m;
    const int base = nd < TRMM_BASE ? nd : TRMM_BASE;
    c.triw = (T*)ws_alloc((size_t)even_ld(base, sizeof(T)) * base * sizeof(T));
    c.ldo = even_ld(c.left ? base : other, sizeof(T));
    c.outw = (T*)ws_alloc((size_t)c.ldo * (c.left ? other : base) * sizeof(T));
    trmm_rec(c, alpha, 0, nd, other);
}
template void trmm_dev<float>(cudaStream_t, char, char, char, char, int, int, float, const float*, int64_t, float*, int64_t);
template void trmm_dev<double>(cudaStream_t, char, char, char, char, int, int, double, const double*, int64_t, double*, int64_t);
template void trmm_dev<cuFloatComplex>(cudaStream_t, char, char, char, char, int, int, cuFloatComplex, const cuFloatComplex*, int64_t, cuFloatComplex*, int64_t);
template void trmm_dev<cuDoubleComplex>(cudaStream_t, char, char, char, char, int, int, cuDoubleComplex, const cuDoubleComplex*, int64_t, cuDoubleComplex*, int64_t);

// ------------------------------------------------------------------ TRSM
// Leaf of the recursion: a diagonal block of at most NBL (64 real / 32 complex) is solved by SUBSTITUTION, in place, like
// netlib's xTRSM and the cublas<t>trsm the reference forwards to (trsm.cc:54-61) -- backward stable for any diagonal block.
// (Round 1 inverted the 64x64 diagonal blocks explicitly and applied the inverses by GEMM, which loses accuracy with
// cond(T_ii) and needed an out-of-place product plus a copy back per leaf.)
//   left : op(T) X = alpha B_blk, every column of B_blk an independent right-hand side
//   right: X op(T) = alpha B_blk  <=>  op(T)^T x^T = alpha b^T, every row of B_blk an independent right-hand side
// so both are "Meff y = alpha b" with Meff = op(T) or op(T)^T, lower (forward) or upper (backward).  One thread owns one
// right-hand side in registers (fully unrolled, compile-time indices); Meff sits in shared memory column by column with
// its diagonal already inverted, and is read by warp-wide broadcast loads; the update after each pivot is right-looking
// (b[i] -= Meff[i][l] * y[l] for all remaining i: independent FMAs).  The 128 right-hand sides of a CTA go through a
// shared-memory tile so that global loads and stores are coalesced for both sides.
template <typename T> struct leaf_width { static constexpr int value = 64; };
template <> struct leaf_width<cuFloatComplex> { static constexpr int value = 32; };
template <> struct leaf_width<cuDoubleComplex> { static constexpr int value = 32; };
constexpr int LEAF_RHS = 128;

template <typename T> __device__ __forceinline__ T tdiv(T a, T b);
template <> __device__ __forceinline__ float tdiv(float a, float b) { return a / b; }
template <> __device__ __forceinline__ double tdiv(double a, double b) { return a / b; }
template <> __device__ __forceinline__ cuFloatComplex tdiv(cuFloatComplex a, cuFloatComplex b) { return cuCdivf(a, b); }
template <> __device__ __forceinline__ cuDoubleComplex tdiv(cuDoubleComplex a, cuDoubleComplex b) { return cuCdiv(a, b); }

template <typename T, bool FWD>
__global__ void __launch_bounds__(LEAF_RHS) trsm_leaf_kernel(int nd, int64_t nrhs, T alpha, const T* __restrict__ S, int64_t lda, bool stored_upper,
                                                            int op /*0 N, 1 T, 2 C*/, bool unit, bool left, T* __restrict__ B, int64_t ldb) {
    constexpr int NBL = leaf_width<T>::value;
    extern __shared__ __align__(16) unsigned char smem_leaf[];
    T* sM = (T*)smem_leaf;                   // sM[l * NBL + i] = Meff[i][l]; diagonal holds 1 / Meff[l][l]
    T* sB = sM + NBL * NBL;                  // sB[c * (NBL + 1) + i]: element i of right-hand side c
    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)blockIdx.x * LEAF_RHS;
    const T zero = num<T>::zero(), one = num<T>::real(1.0);
    // ---- Meff (referenced triangle only) and the right-hand sides: ALL global loads of a phase are issued before the first
    // dependent shared-memory store -- a load -> store chain per element exposes one full memory latency per element
    // (measured: 42 us per leaf of 30720 right-hand sides against ~5 us of traffic), and Meff's loads travel together with the
    // first half of the right-hand sides so the two latencies overlap. ----
    constexpr int M_PER = NBL * NBL / LEAF_RHS;      // Meff elements per thread (32 real / 8 complex)
    constexpr int B_HALF = NBL / 2;                  // right-hand-side elements per thread and half
    T vm[M_PER], vb[B_HALF];
#pragma unroll
    for (int u = 0; u < M_PER; u++) {
        const int idx = tid + u * LEAF_RHS;
        const int i = idx % NBL, l = idx / NBL;                  // Meff[i][l]
        int a = left ? i : l, b = left ? l : i;                  // = op(S)[a][b]
        if (op != 0) { const int t = a; a = b; b = t; }          // = S[a][b] (conjugated if op == 2)
        vm[u] = zero;
        if (i < nd && l < nd) {
            if (a == b) { if (!unit) vm[u] = S[a + (int64_t)b * lda]; }
            else if (stored_upper ? a < b : a > b) vm[u] = S[a + (int64_t)b * lda];
        }
    }
    auto rhs_index = [&](int idx, int& i, int& c) { if (left) { i = idx % NBL; c = idx / NBL; } else { c = idx % LEAF_RHS; i = idx / LEAF_RHS; } };
#pragma unroll
    for (int h = 0; h < 2; h++) {
#pragma unroll
        for (int u = 0; u < B_HALF; u++) {
            int i, c; rhs_index(tid + (h * B_HALF + u) * LEAF_RHS, i, c);
            vb[u] = zero;
            if (i < nd && c0 + c < nrhs) vb[u] = left ? B[i + (c0 + c) * ldb] : B[(c0 + c) + (int64_t)i * ldb];
        }
        if (h == 0) {
#pragma unroll
            for (int u = 0; u < M_PER; u++) {
                const int idx = tid + u * LEAF_RHS;
                const int i = idx % NBL, l = idx / NBL;
                T w = vm[u];
                if (op == 2) w = num<T>::conj(w);
                if (i == l) w = (i < nd && !unit) ? tdiv<T>(one, w) : one;      // reciprocal pivot; unit diagonal and identity padding
                sM[l * NBL + i] = w;
            }
        }
#pragma unroll
        for (int u = 0; u < B_HALF; u++) {
            int i, c; rhs_index(tid + (h * B_HALF + u) * LEAF_RHS, i, c);
            sB[c * (NBL + 1) + i] = num<T>::mul(alpha, vb[u]);
        }
    }
    __syncthreads();
    T y[NBL];
#pragma unroll
    for (int i = 0; i < NBL; i++) y[i] = sB[tid * (NBL + 1) + i];
    if (FWD) {
#pragma unroll
        for (int l = 0; l < NBL; l++) {
            y[l] = num<T>::mul(y[l], sM[l * NBL + l]);
#pragma unroll
            for (int i = l + 1; i < NBL; i++) y[i] = num<T>::sub(y[i], num<T>::mul(sM[l * NBL + i], y[l]));
        }
    } else {
#pragma unroll
        for (int l = NBL - 1; l >= 0; l--) {
            y[l] = num<T>::mul(y[l], sM[l * NBL + l]);
#pragma unroll
            for (int i = 0; i < l; i++) y[i] = num<T>::sub(y[i], num<T>::mul(sM[l * NBL + i], y[l]));
        }
    }
#pragma unroll
    for (int i = 0; i < NBL; i++) sB[tid * (NBL + 1) + i] = y[i];
    __syncthreads();
    for (int idx = tid; idx < NBL * LEAF_RHS; idx += LEAF_RHS) {
        int i, c;
        if (left) { i = idx % NBL; c = idx / NBL; } else { c = idx % LEAF_RHS; i = idx / LEAF_RHS; }
        if (i < nd && c0 + c < nrhs) {
            const T v = sB[c * (NBL + 1) + i];
            if (left) B[i + (c0 + c) * ldb] = v; else B[(c0 + c) + (int64_t)i * ldb] = v;
        }
    }
}

template <typename T> struct TsCtx {
    cudaStream_t s; bool left, upper, unit; char trans; const T* A; int64_t lda; T* B; int64_t ldb;
};

template <typename T>
static void trsm_leaf(const TsCtx<T>& c, T alpha, int d0, int nd, int other) {
    constexpr int NBL = leaf_width<T>::value;
    const int op = op_code(c.trans);
    const bool eff_lower = (!c.upper) ^ (op != 0) ^ (!c.left);
    const int smem = (NBL * NBL + LEAF_RHS * (NBL + 1)) * (int)sizeof(T);
    const T* S = c.A + d0 + (int64_t)d0 * c.lda;
    T* Bb = c.left ? c.B + d0 : c.B + (int64_t)d0 * c.ldb;
    const unsigned grid = (unsigned)((other + LEAF_RHS - 1) / LEAF_RHS);
    if (eff_lower) {
        set_max_dynamic_smem((const void*)trsm_leaf_kernel<T, true>, smem);
        trsm_leaf_kernel<T, true><<<grid, LEAF_RHS, smem, c.s>>>(nd, other, alpha, S, c.lda, c.upper, op, c.unit, c.left, Bb, c.ldb);
    } else {
        set_max_dynamic_smem((const void*)trsm_leaf_kernel<T, false>, smem);
        trsm_leaf_kernel<T, false><<<grid, LEAF_RHS, smem, c.s>>>(nd, other, alpha, S, c.lda, c.upper, op, c.unit, c.left, Bb, c.ldb);
    }
}

template <typename T>
static void trsm_rec(const TsCtx<T>& c, T alpha, int d0, int nd, int other) {
    constexpr int NBL = leaf_width<T>::value;
    const bool notr = op_code(c.trans) == 0;
    const bool opupper = notr ? c.upper : !c.upper;
    if (nd <= NBL) { trsm_leaf<T>(c, alpha, d0, nd, other); return; }
    const int n1 = split_point(nd, NBL), n2 = nd - n1;
    const int a = d0, b = d0 + n1;
    const T one = num<T>::real(1.0), mone = num<T>::real(-1.0);
    const T* Aoff = c.upper ? (c.A + a + (int64_t)b * c.lda) : (c.A + b + (int64_t)a * c.lda);
    const char to = notr ? 'N' : c.trans;
    if (c.left) {
        T* B1 = c.B + a; T* B2 = c.B + b;
        if (!opupper) {   // [T11 0; T21 T22][X1;X2] = alpha [B1;B2]
            trsm_rec(c, alpha, a, n1, other);
            gemm_dev<T>(c.s, to, 'N', n2, other, n1, mone, Aoff, c.lda, B1, c.ldb, alpha, B2, c.ldb);
            trsm_rec(c, one, b, n2, other);
        } else {          // [T11 T12; 0 T22]
            trsm_rec(c, alpha, b, n2, other);
            gemm_dev<T>(c.s, to, 'N', n1, other, n2, mone, Aoff, c.lda, B2, c.ldb, alpha, B1, c.ldb);
            trsm_rec(c, one, a, n1, other);
        }
    } else {
        T* B1 = c.B + (int64_t)a * c.ldb; T* B2 = c.B + (int64_t)b * c.ldb;
        if (!opupper) {   // [X1 X2][T11 0; T21 T22] = alpha [B1 B2]:  X2 T22 = aB2 ; X1 T11 = aB1 - X2 T21
            trsm_rec(c, alpha, b, n2, other);
            gemm_dev<T>(c.s, 'N', to, other, n1, n2, mone, B2, c.ldb, Aoff, c.lda, alpha, B1, c.ldb);
            trsm_rec(c, one, a, n1, other);
        } else {          // X1 T11 = aB1 ; X2 T22 = aB2 - X1 T12
            trsm_rec(c, alpha, a, n1, other);
            gemm_dev<T>(c.s, 'N', to, other, n2, n1, mone, B1, c.ldb, Aoff, c.lda, alpha, B2, c.ldb);
            trsm_rec(c, one, b, n2, other);
        }
    }
}

template <typename T>
void trsm_dev(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* A, int64_t lda, T* B,
              int64_t ldb) {
    if (m <= 0 || n <= 0) return;
    if (num<T>::is_zero(alpha)) { scale_matrix<T>(s, m, n, num<T>::zero(), B, ldb, MASK_FULL); last_variant = VAR_SCALE_ONLY; return; }
    TsCtx<T> c;
    c.s = s; c.left = (side == 'L' || side == 'l'); c.upper = (uplo == 'U' || uplo == 'u'); c.unit = (diag == 'U' || diag == 'u');
    c.trans = op_code(trans) == 0 ? 'N' : (op_code(trans) == 1 ? 'T' : 'C');
    c.A = A; c.lda = lda; c.B = B; c.ldb = ldb;
    const int nd = c.left ? m : n, other = c.left ? n : m;
    trsm_rec(c, alpha, 0, nd, other);
}
template void trsm_dev<float>(cudaStream_t, char, char, char, char, int, int, float, const float*, int64_t, float*, int64_t);
template void trsm_dev<double>(cudaStream_t, char, char, char, char, int, int, double, const double*, int64_t, double*, int64_t);
template void trsm_dev<cuFloatComplex>(cudaStream_t, char, char, char, char, int, int, cuFloatComplex, const cuFloatComplex*, int64_t, cuFloatComplex*, int64_t);
template void trsm_dev<cuDoubleComplex>(cudaStream_t, char, char, char, char, int, int, cuDoubleComplex, const cuDoubleComplex*, int64_t, cuDoubleComplex*, int64_t);

}  // namespace b200
