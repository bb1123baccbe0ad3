// level2_struct.cu -- the banded, packed, Hermitian and complex Level-2 routines the reference's dead wrappers name
// (blas_level2/ger.cc, syr.cc, symv.cc, trmv.cc (real) and blas_level2/gbmv.cc, bmv.cc, pmv.cc, hemv.cc, her.cc, her2.cc, hpr.cc, hpr2.cc, spr.cc,
// spr2.cc, syr2.cc, tbmv.cc, tbsv.cc, tpmv.cc, tpsv.cc, ger.cc (geru/gerc), trmv.cc (c/z) forward to cublas<t>...).
// Fortran + CBLAS entry points with the netlib argument checks; s/d/c/z as netlib defines them.
//
// All are one pass over the stored part of the matrix (HBM-bound).  The index logic of every storage scheme lives in
// structured.cuh; the kernels here are its three bodies behind a launch shape:
//   matrix-vector   y = alpha*(N part + T part [+ x]) + beta*y    "N part": thread per row, column chunks, partial rows summed
//                                                                   in chunk order; "T part": warp per column, butterfly sum
//       GBMV 'N' = N part, 'T'/'C' = T part; SBMV/HBMV/SPMV/HPMV/HEMV = N part over the stored triangle + T part over
//       the strict triangle (conjugated for Hermitian); TBMV/TPMV/TRMV pick one part by trans.
//   rank updates    thread per row over the stored columns (GERU/GERC/HER/HER2/SYR2/SPR/SPR2/HPR/HPR2)
//   solves          TBSV/TPSV: a panel of columns per CTA (32-wide diagonal blocks solved by one warp -- coefficients
//                   preloaded, x passed by shuffle -- the CTA updating the rest of the panel in between), then the same
//                   N/T bodies subtract the panel's contribution from the rows it reaches outside; a narrow band is one launch.
// Row-major CBLAS calls map onto the same kernels: the row-major array is the column-major storage of the transpose with
// uplo flipped (kl/ku swapped); ConjTrans becomes "conjugate, no transpose" (a flag of the bodies), and Hermitian
// row-major operands are the conjugate of the column-major view (flags toggled / vectors conjugated while gathered).
// Deterministic: fixed grids, fixed summation order.
#include "abi_common.h"
#include "structured.cuh"
#include "../../include/b200blas.h"
#include <cstdlib>
#include <type_traits>

namespace b200 {
using namespace st;

typedef cuFloatComplex c32;
typedef cuDoubleComplex c64;

// ------------------------------------------------ kernels ------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ c32 warp_sum(c32 v) { return make_cuFloatComplex(warp_sum(v.x), warp_sum(v.y)); }
__device__ __forceinline__ c64 warp_sum(c64 v) { return make_cuDoubleComplex(warp_sum(v.x), warp_sum(v.y)); }
__device__ __forceinline__ float warp_bcast(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double warp_bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ c32 warp_bcast(c32 v, int src) { return make_cuFloatComplex(warp_bcast(v.x, src), warp_bcast(v.y, src)); }
__device__ __forceinline__ c64 warp_bcast(c64 v, int src) { return make_cuDoubleComplex(warp_bcast(v.x, src), warp_bcast(v.y, src)); }

// dst(i) = [conj] src(element i of the strided BLAS vector)
template <typename T> __global__ void gather_kernel(int n, const T* __restrict__ src, int64_t inc, T* __restrict__ dst, bool conj) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T v = src[vpos(i, n, inc)];
    dst[i] = conj ? el<T>::conj(v) : v;
}
template <typename T> __global__ void scatter_kernel(int n, const T* __restrict__ src, T* __restrict__ dst, int64_t inc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[vpos(i, n, inc)] = src[i];
}
// part[chunk][i] = N part of row i over the chunk's columns; rows [row0,row1), columns [c_lo,c_hi) cut into chunks of cpc
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS) npart_kernel(Desc D, const T* __restrict__ A, const T* __restrict__ v, int row0, int row1, int c_lo, int c_hi, int cpc,
                                                    int flags, T* __restrict__ part, int64_t npad) {
    const int i = row0 + blockIdx.x * ROW_THREADS + threadIdx.x;
    if (i >= row1) return;
    const int c0 = c_lo + blockIdx.y * cpc, c1 = st_min(c_hi, c0 + cpc);
    part[(int64_t)blockIdx.y * npad + i] = npart_row<T>(D, A, v, i, c0, c1, flags, i - (threadIdx.x & 31));
}
// tpart[j] = T part of column j over rows [r0,r1); one warp per column in [col0,col1)
template <typename T>
__global__ void __launch_bounds__(COL_WARPS * 32) tpart_kernel(Desc D, const T* __restrict__ A, const T* __restrict__ v, int col0, int col1, int r0, int r1, int flags,
                                                    T* __restrict__ tpart) {
    const int j = col0 + blockIdx.x * COL_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= col1) return;   // whole warp
    const T acc = warp_sum(tpart_lane<T>(D, A, v, j, lane, 32, r0, r1, flags));
    if (lane == 0) tpart[j] = acc;
}
// ---- one-pass symmetric product (structured.cuh: sym_row) ----
template <typename T> __device__ __forceinline__ T shfl_xor_any(T x, int m) {
    union { T t; unsigned u[sizeof(T) / 4]; } a;
    a.t = x;
#pragma unroll
    for (int q = 0; q < (int)(sizeof(T) / 4); q++) a.u[q] = __shfl_xor_sync(0xffffffffu, a.u[q], m);
    return a.t;
}
// Per-warp sink of sym_row: t[u] is this lane's (row's) product for column j + u.  Transposing butterfly: at mask 16 the lanes
// with that bit clear keep the low half of the columns and hand the high half to their partner (and vice versa), at 8 and 4 the
// same on what is left -- after log2(NU) exchanges a lane holds ONE column's sum over 32 / NU... lanes, the remaining masks finish
// it.  NU + log2(32) - 1 shuffles per step instead of 5 NU.  The lane with the low bits clear stores the column's sum.
template <typename T> struct WarpStripSink {
    enum { NU = unroll_of<T>::N };
    T* strip; int cw0, cw1, lane;
    __device__ __forceinline__ void step(int j, T (&t)[NU]) {
        int col = 0, m = 16;
#pragma unroll
        for (int cnt = NU / 2; cnt >= 1; cnt >>= 1, m >>= 1) {
            const bool up = (lane & m) != 0;
#pragma unroll
            for (int u = 0; u < cnt; u++) {
                const T keep = up ? t[u + cnt] : t[u], send = up ? t[u] : t[u + cnt];
                t[u] = el<T>::add(keep, shfl_xor_any<T>(send, m));
            }
            col += up ? cnt : 0;
        }
        T sum = t[0];
#pragma unroll
        for (; m >= 1; m >>= 1) sum = el<T>::add(sum, shfl_xor_any<T>(sum, m));
        const int jj = j + col;
        if ((lane & (32 / NU - 1)) == 0 && jj >= cw0 && jj < cw1) strip[jj - cw0] = sum;
    }
};
// grid (row blocks, column chunks); dynamic shared memory: one strip of `wstride` sums per warp + `wstride` elements of v
template <typename T, int MINB>
__global__ void __launch_bounds__(ROW_THREADS, MINB) sympart_kernel(Desc D, const T* __restrict__ A, const T* __restrict__ v, int cpc, int nflags, int tflags, T* __restrict__ part,
                                                               int64_t npad, T* __restrict__ tp2, int64_t npadw, int wstride) {
    extern __shared__ __align__(16) unsigned char sym_smem[];
    T* strips = reinterpret_cast<T*>(sym_smem);
    const int r0 = blockIdx.x * ROW_THREADS, i = r0 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = blockIdx.y * cpc, c1 = st_min(D.n, c0 + cpc);
    int cw0, cw1;
    sym_window(D, r0, c0, c1, cw0, cw1);
    T* vwin = strips + (ROW_THREADS / 32) * wstride;           // v over the CTA's window of columns
    for (int t = threadIdx.x; t < (ROW_THREADS / 32) * wstride; t += ROW_THREADS) strips[t] = el<T>::zero();
    for (int t = threadIdx.x; t < cw1 - cw0; t += ROW_THREADS) vwin[t] = v[cw0 + t];
    __syncthreads();
    WarpStripSink<T> sink = {strips + warp * wstride, cw0, cw1, lane};
    const T r = sym_row<T>(D, A, v, vwin - cw0, i, i - lane, c0, c1, nflags, tflags, sink);
    if (i < D.n) part[(int64_t)blockIdx.y * npad + i] = r;
    __syncthreads();
    T* row = tp2 + (int64_t)blockIdx.x * npadw - sym_jw0(D, r0);
    for (int t = threadIdx.x; t < cw1 - cw0; t += ROW_THREADS) {
        T s = strips[t];
#pragma unroll
        for (int w = 1; w < ROW_THREADS / 32; w++) s = el<T>::add(s, strips[w * wstride + t]);
        row[cw0 + t] = s;
    }
}
// 64 columns per CTA, four threads per column: each sums a quarter of the column's row blocks, the quarters are combined in order
// through shared memory (one thread per column walked up to n / 128 values strided by a whole tp2 row: 61 us for n = 32768, 7 % of
// the product, latency-bound -- profiles/r02fin_sympart_v3_ncu_raw.csv)
constexpr int FIN_COLS = 64, FIN_SPLIT = 4;
template <typename T>
__global__ void __launch_bounds__(FIN_COLS * FIN_SPLIT) sym_finish_kernel(Desc D, int nparts, const T* __restrict__ part, int64_t npad, const T* __restrict__ tp2, int64_t npadw,
                                                                            T alpha, T beta, T* __restrict__ out, int64_t inco) {
    __shared__ T quarter[FIN_SPLIT][FIN_COLS];
    const int c = threadIdx.x % FIN_COLS, g = threadIdx.x / FIN_COLS, split = blockDim.x / FIN_COLS;   // bands launch with split = 1
    const int j = blockIdx.x * FIN_COLS + c;
    T q = el<T>::zero();
    if (j < D.n) {
        int rb_lo, rb_hi;
        sym_col_blocks(D, j, rb_lo, rb_hi);
        const int cnt = rb_hi - rb_lo + 1, per = (cnt + split - 1) / split;
        const int r0 = rb_lo + g * per, r1 = st_min(rb_hi, r0 + per - 1);
        if (r0 <= r1) q = sym_tp2_sum<T>(D, j, tp2, npadw, r0, r1);
    }
    quarter[g][c] = q;
    __syncthreads();
    if (g == 0 && j < D.n) {
        T m = quarter[0][c];
        for (int t = 1; t < split; t++) m = el<T>::add(m, quarter[t][c]);
        T* p = out + vpos(j, D.n, inco);
        const bool beta0 = el<T>::is_zero(beta);
        *p = sym_finish_value<T>(j, nparts, part, npad, m, alpha, beta, beta0, beta0 ? el<T>::zero() : *p);
    }
}
template <typename T>
__global__ void smv_finish_kernel(int n, int nparts, const T* __restrict__ part, int64_t npad, const T* __restrict__ tpart, const T* __restrict__ vunit, T alpha,
                                  T beta, T* __restrict__ out, int64_t inco) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T* p = out + vpos(i, n, inco);
    const bool beta0 = el<T>::is_zero(beta);
    *p = finish_elem<T>(i, nparts, part, npad, tpart, vunit, alpha, beta, beta0, beta0 ? el<T>::zero() : *p);
}
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS) rank_kernel(Desc D, T* __restrict__ A, int rows, int ncols, int cpc, T alpha, const T* __restrict__ x, const T* __restrict__ y,
                                                   int mode) {
    const int i = blockIdx.x * ROW_THREADS + threadIdx.x;
    if (i >= rows) return;
    const int c0 = blockIdx.y * cpc, c1 = st_min(ncols, c0 + cpc);
    rank_row<T>(D, A, i, c0, c1, alpha, x, y, mode, i - (threadIdx.x & 31));
}
// x(i) -= N part of row i over the solved block's columns [b0,b1)      (rows [row0,row1) lie outside the block)
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS) solve_nupdate_kernel(Desc D, const T* __restrict__ A, T* x, int row0, int row1, int b0, int b1, int flags) {
    const int i = row0 + blockIdx.x * ROW_THREADS + threadIdx.x;
    if (i >= row1) return;
    x[i] = el<T>::sub(x[i], npart_row<T>(D, A, x, i, b0, b1, flags, i - (threadIdx.x & 31)));
}
// x(j) -= T part of column j over the solved block's rows [b0,b1)      (columns [col0,col1) lie outside the block)
template <typename T>
__global__ void __launch_bounds__(COL_WARPS * 32) solve_tupdate_kernel(Desc D, const T* __restrict__ A, T* x, int col0, int col1, int b0, int b1, int flags) {
    const int j = col0 + blockIdx.x * COL_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= col1) return;
    const T acc = warp_sum(tpart_lane<T>(D, A, x, j, lane, 32, b0, b1, flags));
    if (lane == 0) x[j] = el<T>::sub(x[j], acc);
}
// ---- triangular panel solve ----
// One warp solves the nb <= 32 unknowns of a diagonal block of op(S): lane l owns x(b0+l); each step broadcasts the solved
// unknown by shuffle and eliminates it from the lanes still waiting (x * (1/d) instead of x / d: one FP64 division per lane off
// the critical path instead of 32 in sequence; within the solves' stated tolerance).  See solve_panel_kernel.
// Row r of op(S) over the columns of one diagonal block: the stored range [lo, hi) of c and the address of element (r, c) as an
// affine walk from c = b0 (packed rows of the untransposed triangle: the closed form).  The generic stored() / off() pair costs a
// switch and 64-bit multiplies per ELEMENT; written per element in four fully unrolled 32-step loops it made this kernel 20 600
// instructions (320 KB of SASS): the single CTA spent its time on instruction fetch ("no instruction" + barrier stalls 15 of 20
// cycles per issue, 30 900 cycles per 32-block; ncu source page of run r02z).
struct OpRow { int lo, hi, r, n; int64_t a0, cs; bool packed_row, up; };
__device__ __forceinline__ OpRow op_row(const Desc& D, int r, int b0, bool trans) {
    OpRow w;
    const int rs = r < D.n ? r : D.n - 1;
    if (trans) col_rows(D, rs, w.lo, w.hi); else row_cols(D, rs, w.lo, w.hi);
    if (r >= D.n) w.lo = w.hi = 0;
    w.r = rs; w.n = D.n; w.up = D.upper != 0;
    w.packed_row = D.kind == K_PACKED && !trans;
    w.a0 = trans ? off(D, b0, rs) : off(D, rs, b0);
    w.cs = trans ? 1 : (D.kind == K_PACKED ? 0 : col_step(D, 0));
    return w;
}
template <bool PACKED_ROW> __device__ __forceinline__ int64_t op_row_addr(const OpRow& w, int b0, int c) {
    if (PACKED_ROW) return (int64_t)w.r + (w.up ? (int64_t)c * ((int64_t)c + 1) / 2 : (int64_t)c * (2 * (int64_t)w.n - c - 1) / 2);
    return w.a0 + (int64_t)(c - b0) * w.cs;
}
// (the rarely taken plain path, kept out of line: its inlined N / T bodies are thousands of instructions)
template <typename T> __device__ __noinline__ T panel_update_call(const Desc& D, const T* A, const T* x, int r, int b0, int b1, bool trans, int flags) {
    return panel_update<T>(D, A, x, r, b0, b1, trans, flags);
}
// One CTA solves the panel [p0,p1) of op(S) in place, 32-block by 32-block (structured.cuh: solve(), panel_block()).  The
// dependency chain through x is what bounds a triangular solve, so nothing that does NOT depend on x may sit on it, and the code
// on it must be short enough to stay in the instruction cache:
//   * warps 1 and 2 STAGE the next block's diagonal coefficients in shared memory -- already conjugated, masked (zero for lanes that
//     do not wait on that step) and scaled by the pivot reciprocals -- while warp 0 solves the current block from the buffer filled one
//     block earlier (two buffers take turns);
//   * warp 0's elimination is then a rolled loop of one shared-memory load, one broadcast and one FMA per step;
//   * warps 3..7 load the coefficients op(S)(r, b0..b1) of the panel row r they will update while warp 0 solves, and after the
//     barrier only multiply them with the 32 solved unknowns (shared memory) and subtract;
//   * rows beyond the 160 those warps hold (wide bands, the last rows of a packed panel) take the plain path (panel_update).
// History (DTBSV n = 2^18, k = 127; profiles/r01g_level2_struct_summary.txt, r02y_solves.txt): loads, solve, barrier, update loads,
// update, barrier in series 146 ms; loads off the chain 127 ms -- ncu showed the real cost: 20 600 instructions of unrolled
// per-element stored()/off() logic, the single CTA stalled on instruction fetch; affine address walks 57 ms; this form: see profiles/.
template <typename T>
__global__ void __launch_bounds__(SOLVE_THREADS) solve_panel_kernel(Desc D, const T* __restrict__ A, T* x, int p0, int p1, bool trans, bool conj, bool unit,
                                                                    bool forward) {
    __shared__ T xs[32];
    __shared__ T dco[2][32][33];                       // [buffer][step][lane]: masked diagonal coefficients in elimination order
    __shared__ T dinv[2][32];                          // pivot reciprocals (one for a unit diagonal)
    __shared__ T xn[32];                               // the next block's unknowns after this block's update
    bool xn_valid = false;
    const int nblk = (p1 - p0 + SOLVE_NB - 1) / SOLVE_NB, flags = conj ? F_CONJ : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int HELD = SOLVE_THREADS - 96;            // panel rows whose update coefficients are preloaded (warps 3..7)
    // warps 1 and 2, half the elimination steps each (one warp doing all 32 was the slowest warp of the CTA: ~2000 straight-line
    // instructions per block against ~250 on warp 0's chain -- ncu source page, run r02z2)
    auto stage_diag = [&](int bi, int half) {           // block bi's diagonal coefficients -> dco[bi & 1], dinv[bi & 1]
        int b0, b1, u0, u1;
        panel_block(D, p0, p1, forward, bi, b0, b1, u0, u1);
        const int nb = b1 - b0, r = b0 + lane, buf = bi & 1;
        const OpRow w = op_row(D, r, b0, trans);
        const int64_t dpos = off(D, w.r, w.r);           // the lane's own diagonal element: always stored
        T cf[16];
        unsigned okm = 0;
        auto body = [&](auto packed_tag) {
#pragma unroll
            for (int q = 0; q < 16; q++) {               // all loads in flight at once
                const int step = half * 16 + q;
                const int jj = forward ? step : nb - 1 - step;
                const int c = b0 + jj;
                const bool waiting = forward ? lane > jj : lane < jj;
                const bool ok = step < nb && lane < nb && waiting && c >= w.lo && c < w.hi;
                cf[q] = A[ok ? op_row_addr<decltype(packed_tag)::value>(w, b0, c) : dpos];
                okm |= (ok ? 1u : 0u) << q;
            }
        };
        if (w.packed_row) body(std::true_type{}); else body(std::false_type{});
        T piv = (unit || lane >= nb) ? el<T>::one() : A[dpos];
        if (conj) piv = el<T>::conj(piv);
        const T di = el<T>::div(el<T>::one(), piv);     // one division per lane, all lanes at once
        if (half == 0) dinv[buf][lane] = di;
#pragma unroll
        for (int q = 0; q < 16; q++) {                   // conjugate, mask and scale column jj by 1/pivot(jj): the solve then carries the
            const int step = half * 16 + q;              // UNSCALED unknowns y (x = y/d applied at the end, off the chain)
            const int jj = forward ? step : nb - 1 - step;
            const bool ok = (okm >> q) & 1u;
            const T a = conj ? el<T>::conj(cf[q]) : cf[q];
            const T dj = warp_bcast(di, jj & 31);
            dco[buf][step][lane] = ok ? el<T>::mul(a, dj) : el<T>::zero();
        }
    };
    if ((warp == 1 || warp == 2) && nblk > 0) stage_diag(0, warp - 1);
    __syncthreads();
    for (int bi = 0; bi < nblk; bi++) {
        int b0, b1, u0, u1;
        panel_block(D, p0, p1, forward, bi, b0, b1, u0, u1);
        const int nb = b1 - b0, buf = bi & 1;
        T cu[32];                                       // warps 2..7: the update row's coefficients
        unsigned uok = 0;
        const int r = u0 + (tid - 96);
        const bool mine = tid >= 96 && r < u1;
        if (warp == 0) {
            // per step on the chain: one broadcast and one FMA (coefficients pre-scaled by the pivot reciprocals; y = d x)
            const int rr = b0 + lane;
            T yv = lane < nb ? (xn_valid ? xn[lane] : x[rr]) : el<T>::zero();
#pragma unroll 4
            for (int step = 0; step < nb; step++) {
                const int jj = forward ? step : nb - 1 - step;
                const T c = dco[buf][step][lane];
                const T yj = warp_bcast(yv, jj);
                const bool waiting = forward ? (lane > jj && lane < nb) : lane < jj;
                if (waiting) yv = el<T>::sub(yv, el<T>::mul(c, yj));   // (guarded, not just masked: an Inf in x must not reach solved lanes as 0 * Inf)
            }
            const T xv = el<T>::mul(yv, dinv[buf][lane]);         // (1 for a unit diagonal)
            if (lane < nb) x[rr] = xv;
            xs[lane] = lane < nb ? xv : el<T>::zero();
        } else if (warp <= 2) {
            if (bi + 1 < nblk) stage_diag(bi + 1, warp - 1);
        } else if (mine) {
            const OpRow w = op_row(D, r, b0, trans);
            const int64_t safe = off(D, b0, b0);
            auto body = [&](auto packed_tag) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int c = b0 + j;
                    const bool ok = j < nb && c >= w.lo && c < w.hi;
                    cu[j] = A[ok ? op_row_addr<decltype(packed_tag)::value>(w, b0, c) : safe];
                    uok |= (ok ? 1u : 0u) << j;
                }
            };
            if (w.packed_row) body(std::true_type{}); else body(std::false_type{});
        }
        // the next block's unknowns also go to shared memory (saves warp 0 a round trip through L2 at the head of its chain) when
        // all of them are among the held rows -- a narrow band or a wide one leaves some to the plain path: then warp 0 reads x
        int nb0 = 0, nb1 = 0;
        bool cover = false;
        if (bi + 1 < nblk) {
            int t0, t1;
            panel_block(D, p0, p1, forward, bi + 1, nb0, nb1, t0, t1);
            cover = nb0 >= u0 && nb1 <= u1 && nb0 - u0 < HELD && nb1 - u0 <= HELD;
        }
        __syncthreads();
        if (mine) {
            T q0 = el<T>::zero(), q1 = q0, q2 = q0, q3 = q0;      // four chains of 8, not one of 32
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const T a0 = conj ? el<T>::conj(cu[j]) : cu[j], a1 = conj ? el<T>::conj(cu[j + 1]) : cu[j + 1];
                const T a2 = conj ? el<T>::conj(cu[j + 2]) : cu[j + 2], a3 = conj ? el<T>::conj(cu[j + 3]) : cu[j + 3];
                if ((uok >> j) & 1u) q0 = el<T>::mad(a0, xs[j], q0);
                if ((uok >> (j + 1)) & 1u) q1 = el<T>::mad(a1, xs[j + 1], q1);
                if ((uok >> (j + 2)) & 1u) q2 = el<T>::mad(a2, xs[j + 2], q2);
                if ((uok >> (j + 3)) & 1u) q3 = el<T>::mad(a3, xs[j + 3], q3);
            }
            const T nx = el<T>::sub(x[r], el<T>::add(el<T>::add(q0, q1), el<T>::add(q2, q3)));
            x[r] = nx;
            if (cover && r >= nb0 && r < nb1) xn[r - nb0] = nx;
        }
        xn_valid = cover;
        for (int r2 = u0 + HELD + tid; r2 < u1; r2 += SOLVE_THREADS) x[r2] = el<T>::sub(x[r2], panel_update_call<T>(D, A, x, r2, b0, b1, trans, flags));
        __syncthreads();
    }
}

// ------------------------------------------------ the device backend of the plans ------------------------------------------------
struct DeviceBackend {
    cudaStream_t s;
    explicit DeviceBackend(cudaStream_t st) : s(st) {}
    void* alloc(size_t bytes) { return ws_alloc(bytes); }
    int sm_target() const { return (sm_count() > 0 ? sm_count() : 148) * 8; }
    template <typename T> void gather(int n, const T* src, int64_t inc, T* dst, bool conj) { gather_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, src, inc, dst, conj); }
    template <typename T> void scatter(int n, const T* src, T* dst, int64_t inc) { scatter_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, src, dst, inc); }
    template <typename T>
    void npart(const Desc& D, const T* A, const T* v, int row0, int row1, int c_lo, int c_hi, int cpc, int nchunks, int flags, T* part, int64_t npad) {
        npart_kernel<T><<<dim3((row1 - row0 + ROW_THREADS - 1) / ROW_THREADS, nchunks), ROW_THREADS, 0, s>>>(D, A, v, row0, row1, c_lo, c_hi, cpc, flags, part, npad);
    }
    template <typename T> void tpart(const Desc& D, const T* A, const T* v, int col0, int col1, int r0, int r1, int flags, T* tp) {
        tpart_kernel<T><<<(col1 - col0 + COL_WARPS - 1) / COL_WARPS, COL_WARPS * 32, 0, s>>>(D, A, v, col0, col1, r0, r1, flags, tp);
    }
    template <typename T>
    void finish(int n, int nparts, const T* part, int64_t npad, const T* tp, const T* vunit, T alpha, T beta, T* out, int64_t inco) {
        smv_finish_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, nparts, part, npad, tp, vunit, alpha, beta, out, inco);
        last_variant = VAR_GENERIC_TILE;
    }
    // one-pass symmetric product: a strip of 1024 doubles (512 double-complex) per warp + v over the window = 40 KB per CTA, 5 CTAs per SM
    template <typename T> int sym_max_cols() const {
        static const bool off = getenv("B200BLAS_SYM_TWO_PASS") != nullptr;
        return off ? 0 : (int)(8192 / sizeof(T));
    }
    template <typename T>
    void sympart(const Desc& D, const T* A, const T* v, int cpc, int nchunks, int nflags, int tflags, T* part, int64_t npad, T* tp2, int64_t npadw) {
        const int64_t w = sym_width(D) < cpc ? sym_width(D) : (int64_t)cpc;
        const int wstride = (int)((w + 7) / 8 * 8);
        // 4 CTAs per SM (120 registers, no spills) or 5 (96 registers, a few spilled words in the loop).  Measured
        // (profiles/r02w_level2_sym_one_pass_v3.txt): packed / full triangles 80 / 83 % of the HBM peak with 4 against 55 / 74 % with 5;
        // bands (short rows, mostly masked steps) 57.5 % with 5 against 54 % with 4.  B200BLAS_SYM_MINB overrides.
        static const int minb_env = getenv("B200BLAS_SYM_MINB") ? atoi(getenv("B200BLAS_SYM_MINB")) : 0;
        const int minb = minb_env ? minb_env : (D.kind == K_BAND_TRI ? 5 : 4);
        const dim3 grid((D.n + ROW_THREADS - 1) / ROW_THREADS, nchunks);
        const size_t smem = (size_t)(ROW_THREADS / 32 + 1) * wstride * sizeof(T);
        if (minb == 4) sympart_kernel<T, 4><<<grid, ROW_THREADS, smem, s>>>(D, A, v, cpc, nflags, tflags, part, npad, tp2, npadw, wstride);
        else sympart_kernel<T, 5><<<grid, ROW_THREADS, smem, s>>>(D, A, v, cpc, nflags, tflags, part, npad, tp2, npadw, wstride);
    }
    template <typename T> void sym_finish(const Desc& D, int nparts, const T* part, int64_t npad, const T* tp2, int64_t npadw, T alpha, T beta, T* out, int64_t inco) {
        // a band column collects at most (k + 128) / 128 + 1 row blocks: one thread per column (the split cost 2.5 % there)
        const int split = D.kind == K_BAND_TRI ? 1 : FIN_SPLIT;
        sym_finish_kernel<T><<<(D.n + FIN_COLS - 1) / FIN_COLS, FIN_COLS * split, 0, s>>>(D, nparts, part, npad, tp2, npadw, alpha, beta, out, inco);
        last_variant = VAR_GENERIC_TILE;
    }
    template <typename T> void rank(const Desc& D, T* A, int rows, int ncols, int cpc, int nchunks, T alpha, const T* x, const T* y, int mode) {
        rank_kernel<T><<<dim3((rows + ROW_THREADS - 1) / ROW_THREADS, nchunks), ROW_THREADS, 0, s>>>(D, A, rows, ncols, cpc, alpha, x, y, mode);
        last_variant = VAR_GENERIC_TILE;
    }
    template <typename T> void solve_panel(const Desc& D, const T* A, T* x, int p0, int p1, bool trans, bool conj, bool unit, bool forward) {
        solve_panel_kernel<T><<<1, SOLVE_THREADS, 0, s>>>(D, A, x, p0, p1, trans, conj, unit, forward);
        last_variant = VAR_GENERIC_TILE;
    }
    template <typename T> void solve_nupdate(const Desc& D, const T* A, T* x, int row0, int row1, int b0, int b1, int flags) {
        solve_nupdate_kernel<T><<<(row1 - row0 + ROW_THREADS - 1) / ROW_THREADS, ROW_THREADS, 0, s>>>(D, A, x, row0, row1, b0, b1, flags);
    }
    template <typename T> void solve_tupdate(const Desc& D, const T* A, T* x, int col0, int col1, int b0, int b1, int flags) {
        solve_tupdate_kernel<T><<<(col1 - col0 + COL_WARPS - 1) / COL_WARPS, COL_WARPS * 32, 0, s>>>(D, A, x, col0, col1, b0, b1, flags);
    }
};

// real / complex full-storage TRSV through the panel solver above (fortran_l12.cu: trsv_entry picks it for large n)
template <typename T>
void trsv_struct_dev(cudaStream_t s, char uplo, char trans, char diag, int n, const T* A, int64_t lda, T* x, int64_t incx) {
    DeviceBackend be(s);
    plan_tri<T>(be, K_FULL_TRI, true, false, uplo == 'U', trans, diag == 'U', n, 0, A, lda, x, incx);
}
template void trsv_struct_dev<float>(cudaStream_t, char, char, char, int, const float*, int64_t, float*, int64_t);
template void trsv_struct_dev<double>(cudaStream_t, char, char, char, int, const double*, int64_t, double*, int64_t);
template void trsv_struct_dev<cuFloatComplex>(cudaStream_t, char, char, char, int, const cuFloatComplex*, int64_t, cuFloatComplex*, int64_t);
template void trsv_struct_dev<cuDoubleComplex>(cudaStream_t, char, char, char, int, const cuDoubleComplex*, int64_t, cuDoubleComplex*, int64_t);

}  // namespace b200

using namespace b200;

namespace {

typedef VecOperand Vec;   // runtime.h: strided vectors are written back element by element
inline char trans_code(const char* trans) { return lsame(trans, 'N') ? 'N' : (lsame(trans, 'T') ? 'T' : (lsame(trans, 'C') ? 'C' : '?')); }

// ---- GBMV: netlib info 1,2,3,4,5,8,10,13 ----
template <typename T>
void gbmv_entry(const char* name, bool rowmajor, const char* trans, int m, int n, int kl, int ku, const T* alpha, const T* a, int lda, const T* x, int incx,
                const T* beta, T* y, int incy) {
    const char op = trans_code(trans);
    int info = 0;
    if (op == '?') info = 1; else if (m < 0) info = 2; else if (n < 0) info = 3; else if (kl < 0) info = 4; else if (ku < 0) info = 5;
    else if (lda < kl + ku + 1) info = 8; else if (incx == 0) info = 10; else if (incy == 0) info = 13;
    if (info) { call_xerbla(name, info); return; }
    if (m == 0 || n == 0 || (is0(*alpha) && is1(*beta))) return;
    const int lenx = op == 'N' ? n : m, leny = op == 'N' ? m : n;
    CallScope scope(name);
    const bool a0 = is0(*alpha);   // netlib: A and x are not referenced
    Operand oa(a0 ? nullptr : a, kl + ku + 1, rowmajor ? m : n, lda, sizeof(T), ACC_IN);
    Vec ox(a0 ? nullptr : x, lenx, incx, sizeof(T), ACC_IN), oy(y, leny, incy, sizeof(T), ACC_INOUT);
    DeviceBackend be(current_stream());
    plan_gbmv<T>(be, rowmajor, op, m, n, kl, ku, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ox.dev(), incx, *beta, (T*)oy.dev(), incy);
    oy.release();
    log_exec(name, "%c%s m=%d n=%d kl=%d ku=%d lda=%d", op, rowmajor ? " row-major" : "", m, n, kl, ku, lda);
}

// ---- symmetric / Hermitian matrix-vector products: SBMV/HBMV (info 1,2,3,6,8,11), SPMV/HPMV (1,2,6,9), HEMV (1,2,5,7,10) ----
template <typename T>
void symv_like_entry(const char* name, int kind, bool herm, bool rowmajor, const char* uplo, int n, int k, const T* alpha, const T* a, int lda, const T* x,
                     int incx, const T* beta, T* y, int incy) {
    const bool upper = lsame(uplo, 'U');
    int info = 0;
    if (!upper && !lsame(uplo, 'L')) info = 1; else if (n < 0) info = 2;
    else if (kind == K_BAND_TRI) { if (k < 0) info = 3; else if (lda < k + 1) info = 6; else if (incx == 0) info = 8; else if (incy == 0) info = 11; }
    else if (kind == K_PACKED) { if (incx == 0) info = 6; else if (incy == 0) info = 9; }
    else { if (lda < imax(1, n)) info = 5; else if (incx == 0) info = 7; else if (incy == 0) info = 10; }
    if (info) { call_xerbla(name, info); return; }
    if (n == 0 || (is0(*alpha) && is1(*beta))) return;
    CallScope scope(name);
    const bool a0 = is0(*alpha);
    const int64_t arows = kind == K_BAND_TRI ? k + 1 : (kind == K_PACKED ? packed_len(n) : n), acols = kind == K_PACKED ? 1 : n;
    Operand oa(a0 ? nullptr : a, arows, acols, kind == K_PACKED ? packed_len(n) : lda, sizeof(T), ACC_IN);
    Vec ox(a0 ? nullptr : x, n, incx, sizeof(T), ACC_IN), oy(y, n, incy, sizeof(T), ACC_INOUT);
    DeviceBackend be(current_stream());
    plan_symv_like<T>(be, kind, herm, rowmajor, upper, n, k, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ox.dev(), incx, *beta, (T*)oy.dev(), incy);
    oy.release();
    log_exec(name, "%c%s n=%d k=%d", upper ? 'U' : 'L', rowmajor ? " row-major" : "", n, k);
}

// ---- triangular products and solves: TBMV/TBSV (info 1,2,3,4,5,7,9), TPMV/TPSV (1,2,3,4,7), TRMV (1,2,3,4,6,8) ----
template <typename T>
void tri_entry(const char* name, int kind, bool solve, bool rowmajor, const char* uplo, const char* trans, const char* diag, int n, int k, const T* a, int lda,
               T* x, int incx) {
    const bool upper = lsame(uplo, 'U'), unit = lsame(diag, 'U');
    const char op = trans_code(trans);
    int info = 0;
    if (!upper && !lsame(uplo, 'L')) info = 1; else if (op == '?') info = 2; else if (!unit && !lsame(diag, 'N')) info = 3;
    else if (n < 0) info = 4;
    else if (kind == K_BAND_TRI) { if (k < 0) info = 5; else if (lda < k + 1) info = 7; else if (incx == 0) info = 9; }
    else if (kind == K_PACKED) { if (incx == 0) info = 7; }
    else { if (lda < imax(1, n)) info = 6; else if (incx == 0) info = 8; }
    if (info) { call_xerbla(name, info); return; }
    if (n == 0) return;
    CallScope scope(name);
    const int64_t arows = kind == K_BAND_TRI ? k + 1 : (kind == K_PACKED ? packed_len(n) : n), acols = kind == K_PACKED ? 1 : n;
    Operand oa(a, arows, acols, kind == K_PACKED ? packed_len(n) : lda, sizeof(T), ACC_IN);
    Vec ox(x, n, incx, sizeof(T), ACC_INOUT);
    DeviceBackend be(current_stream());
    plan_tri<T>(be, kind, solve, rowmajor, upper, op, unit, n, k, (const T*)oa.dev(), oa.ld(), (T*)ox.dev(), incx);
    ox.release();
    log_exec(name, "%c%c%c%s n=%d k=%d", upper ? 'U' : 'L', op, unit ? 'U' : 'N', rowmajor ? " row-major" : "", n, k);
}

// ---- GERU / GERC: netlib info 1,2,5,7,9 ----
template <typename T>
void gerx_entry(const char* name, bool conjy, bool rowmajor, int m, int n, const T* alpha, const T* x, int incx, const T* y, int incy, T* a, int lda) {
    int info = 0;
    if (m < 0) info = 1; else if (n < 0) info = 2; else if (incx == 0) info = 5; else if (incy == 0) info = 7;
    else if (lda < imax(1, rowmajor ? n : m)) info = 9;
    if (info) { call_xerbla(name, info); return; }
    if (m == 0 || n == 0 || is0(*alpha)) return;
    CallScope scope(name);
    Vec ox(x, m, incx, sizeof(T), ACC_IN), oy(y, n, incy, sizeof(T), ACC_IN);
    Operand oa(a, rowmajor ? n : m, rowmajor ? m : n, lda, sizeof(T), ACC_INOUT);
    DeviceBackend be(current_stream());
    plan_ger<T>(be, conjy, rowmajor, m, n, *alpha, (const T*)ox.dev(), incx, (const T*)oy.dev(), incy, (T*)oa.dev(), oa.ld());
    oa.release();
    log_exec(name, "m=%d n=%d lda=%d%s", m, n, lda, rowmajor ? " row-major" : "");
}

// ---- symmetric / Hermitian rank-1 and rank-2 updates ----
// full storage: HER (info 1,2,5,7), HER2 / SYR2 (1,2,5,7,9); packed: SPR / HPR (1,2,5), SPR2 / HPR2 (1,2,5,7)
template <typename T>
void rank_sym_entry(const char* name, int kind, int mode, bool rowmajor, const char* uplo, int n, T alpha, const T* x, int incx, const T* y, int incy, T* a,
                    int lda) {
    const bool upper = lsame(uplo, 'U');
    const bool two = mode == R_SYR2 || mode == R_HER2;
    int info = 0;
    if (!upper && !lsame(uplo, 'L')) info = 1; else if (n < 0) info = 2; else if (incx == 0) info = 5;
    else if (two && incy == 0) info = 7;
    else if (kind == K_FULL_TRI && lda < imax(1, n)) info = two ? 9 : 7;
    if (info) { call_xerbla(name, info); return; }
    if (n == 0 || is0(alpha)) return;
    CallScope scope(name);
    Vec ox(x, n, incx, sizeof(T), ACC_IN), oy(two ? y : nullptr, n, incy, sizeof(T), ACC_IN);
    const int64_t arows = kind == K_PACKED ? packed_len(n) : n, acols = kind == K_PACKED ? 1 : n;
    Operand oa(a, arows, acols, kind == K_PACKED ? packed_len(n) : lda, sizeof(T), ACC_INOUT);
    DeviceBackend be(current_stream());
    plan_rank_sym<T>(be, kind, mode, rowmajor, upper, n, alpha, (const T*)ox.dev(), incx, (const T*)oy.dev(), incy, (T*)oa.dev(), oa.ld());
    oa.release();
    log_exec(name, "%c%s n=%d", upper ? 'U' : 'L', rowmajor ? " row-major" : "", n);
}

// ---- CBLAS row-major complex GEMV with ConjTrans: y = alpha*conj(S)*x + beta*y on the column-major m x n view S = A^T ----
template <typename T>
void gemv_conj_notrans(const char* name, int m, int n, const T* alpha, const T* a, int lda, const T* x, int incx, const T* beta, T* y, int incy) {
    CallScope scope(name);   // the caller has validated the arguments
    Operand oa(a, m, n, lda, sizeof(T), ACC_IN);
    Vec ox(x, n, incx, sizeof(T), ACC_IN), oy(y, m, incy, sizeof(T), ACC_INOUT);
    DeviceBackend be(current_stream());
    plan_gemv_conj<T>(be, m, n, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ox.dev(), incx, *beta, (T*)oy.dev(), incy);
    oy.release();
    log_exec(name, "R m=%d n=%d lda=%d", m, n, lda);
}

inline const char* uplo_c(enum CBLAS_UPLO u) { return u == CblasUpper ? "U" : (u == CblasLower ? "L" : "?"); }
inline const char* trans_c(enum CBLAS_TRANSPOSE t) { return t == CblasNoTrans ? "N" : (t == CblasTrans ? "T" : (t == CblasConjTrans ? "C" : "?")); }
inline const char* diag_c(enum CBLAS_DIAG d) { return d == CblasUnit ? "U" : (d == CblasNonUnit ? "N" : "?"); }
inline bool rm(enum CBLAS_ORDER o) { return o == CblasRowMajor; }
inline c32 mkc(float r) { return make_cuFloatComplex(r, 0.f); }
inline c64 mkc(double r) { return make_cuDoubleComplex(r, 0.0); }

}  // namespace

extern "C" {

// ------------------------------- real + complex families with identical signatures -------------------------------
#define B200_GBMV(P, T, CT, SC)                                                                                                                        \
    void P##gbmv_(const char* trans, const int* m, const int* n, const int* kl, const int* ku, const T* alpha, const T* a, const int* lda, const T* x,  \
                  const int* incx, const T* beta, T* y, const int* incy) {                                                                              \
        gbmv_entry<CT>(#P "gbmv_", false, trans, *m, *n, *kl, *ku, (const CT*)alpha, (const CT*)a, *lda, (const CT*)x, *incx, (const CT*)beta, (CT*)y, *incy); } \
    void P##tbmv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const T* a, const int* lda, T* x, const int* incx) { \
        tri_entry<CT>(#P "tbmv_", K_BAND_TRI, false, false, uplo, trans, diag, *n, *k, (const CT*)a, *lda, (CT*)x, *incx); }                            \
    void P##tbsv_(const char* uplo, const char* trans, const char* diag, const int* n, const int* k, const T* a, const int* lda, T* x, const int* incx) { \
        tri_entry<CT>(#P "tbsv_", K_BAND_TRI, true, false, uplo, trans, diag, *n, *k, (const CT*)a, *lda, (CT*)x, *incx); }                             \
    void P##tpmv_(const char* uplo, const char* trans, const char* diag, const int* n, const T* ap, T* x, const int* incx) {                           \
        tri_entry<CT>(#P "tpmv_", K_PACKED, false, false, uplo, trans, diag, *n, 0, (const CT*)ap, 1, (CT*)x, *incx); }                                 \
    void P##tpsv_(const char* uplo, const char* trans, const char* diag, const int* n, const T* ap, T* x, const int* incx) {                           \
        tri_entry<CT>(#P "tpsv_", K_PACKED, true, false, uplo, trans, diag, *n, 0, (const CT*)ap, 1, (CT*)x, *incx); }                                  \
    void cblas_##P##tbmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, enum CBLAS_TRANSPOSE t, enum CBLAS_DIAG d, int n, int k, const SC* a, int lda, SC* x, int incx) { \
        tri_entry<CT>(#P "tbmv_", K_BAND_TRI, false, rm(o), uplo_c(u), trans_c(t), diag_c(d), n, k, (const CT*)a, lda, (CT*)x, incx); }                 \
    void cblas_##P##tbsv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, enum CBLAS_TRANSPOSE t, enum CBLAS_DIAG d, int n, int k, const SC* a, int lda, SC* x, int incx) { \
        tri_entry<CT>(#P "tbsv_", K_BAND_TRI, true, rm(o), uplo_c(u), trans_c(t), diag_c(d), n, k, (const CT*)a, lda, (CT*)x, incx); }                  \
    void cblas_##P##tpmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, enum CBLAS_TRANSPOSE t, enum CBLAS_DIAG d, int n, const SC* ap, SC* x, int incx) {      \
        tri_entry<CT>(#P "tpmv_", K_PACKED, false, rm(o), uplo_c(u), trans_c(t), diag_c(d), n, 0, (const CT*)ap, 1, (CT*)x, incx); }                    \
    void cblas_##P##tpsv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, enum CBLAS_TRANSPOSE t, enum CBLAS_DIAG d, int n, const SC* ap, SC* x, int incx) {      \
        tri_entry<CT>(#P "tpsv_", K_PACKED, true, rm(o), uplo_c(u), trans_c(t), diag_c(d), n, 0, (const CT*)ap, 1, (CT*)x, incx); }
B200_GBMV(s, float, float, float)
B200_GBMV(d, double, double, double)
B200_GBMV(c, b200_c32, c32, void)
B200_GBMV(z, b200_c64, c64, void)
#undef B200_GBMV

void cblas_sgbmv(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE t, int m, int n, int kl, int ku, float alpha, const float* a, int lda, const float* x, int incx, float beta,
                 float* y, int incy) { gbmv_entry<float>("sgbmv_", rm(o), trans_c(t), m, n, kl, ku, &alpha, a, lda, x, incx, &beta, y, incy); }
void cblas_dgbmv(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE t, int m, int n, int kl, int ku, double alpha, const double* a, int lda, const double* x, int incx,
                 double beta, double* y, int incy) { gbmv_entry<double>("dgbmv_", rm(o), trans_c(t), m, n, kl, ku, &alpha, a, lda, x, incx, &beta, y, incy); }
void cblas_cgbmv(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE t, int m, int n, int kl, int ku, const void* alpha, const void* a, int lda, const void* x, int incx,
                 const void* beta, void* y, int incy) {
    gbmv_entry<c32>("cgbmv_", rm(o), trans_c(t), m, n, kl, ku, (const c32*)alpha, (const c32*)a, lda, (const c32*)x, incx, (const c32*)beta, (c32*)y, incy); }
void cblas_zgbmv(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE t, int m, int n, int kl, int ku, const void* alpha, const void* a, int lda, const void* x, int incx,
                 const void* beta, void* y, int incy) {
    gbmv_entry<c64>("zgbmv_", rm(o), trans_c(t), m, n, kl, ku, (const c64*)alpha, (const c64*)a, lda, (const c64*)x, incx, (const c64*)beta, (c64*)y, incy); }

// ------------------------------- real symmetric: SBMV, SPMV, SYR2, SPR, SPR2 -------------------------------
#define B200_REALSYM(P, T)                                                                                                                              \
    void P##sbmv_(const char* uplo, const int* n, const int* k, const T* alpha, const T* a, const int* lda, const T* x, const int* incx, const T* beta, T* y, \
                  const int* incy) { symv_like_entry<T>(#P "sbmv_", K_BAND_TRI, false, false, uplo, *n, *k, alpha, a, *lda, x, *incx, beta, y, *incy); } \
    void P##spmv_(const char* uplo, const int* n, const T* alpha, const T* ap, const T* x, const int* incx, const T* beta, T* y, const int* incy) {      \
        symv_like_entry<T>(#P "spmv_", K_PACKED, false, false, uplo, *n, 0, alpha, ap, 1, x, *incx, beta, y, *incy); }                                   \
    void P##syr2_(const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* a, const int* lda) {      \
        rank_sym_entry<T>(#P "syr2_", K_FULL_TRI, R_SYR2, false, uplo, *n, *alpha, x, *incx, y, *incy, a, *lda); }                                       \
    void P##spr_(const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, T* ap) {                                                   \
        rank_sym_entry<T>(#P "spr_", K_PACKED, R_SYR, false, uplo, *n, *alpha, x, *incx, nullptr, 1, ap, 1); }                                           \
    void P##spr2_(const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* ap) {                     \
        rank_sym_entry<T>(#P "spr2_", K_PACKED, R_SYR2, false, uplo, *n, *alpha, x, *incx, y, *incy, ap, 1); }                                           \
    void cblas_##P##sbmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, int k, T alpha, const T* a, int lda, const T* x, int incx, T beta, T* y, int incy) { \
        symv_like_entry<T>(#P "sbmv_", K_BAND_TRI, false, rm(o), uplo_c(u), n, k, &alpha, a, lda, x, incx, &beta, y, incy); }                            \
    void cblas_##P##spmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, T alpha, const T* ap, const T* x, int incx, T beta, T* y, int incy) {             \
        symv_like_entry<T>(#P "spmv_", K_PACKED, false, rm(o), uplo_c(u), n, 0, &alpha, ap, 1, x, incx, &beta, y, incy); }                               \
    void cblas_##P##syr2(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, T alpha, const T* x, int incx, const T* y, int incy, T* a, int lda) {             \
        rank_sym_entry<T>(#P "syr2_", K_FULL_TRI, R_SYR2, rm(o), uplo_c(u), n, alpha, x, incx, y, incy, a, lda); }                                       \
    void cblas_##P##spr(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, T alpha, const T* x, int incx, T* ap) {                                            \
        rank_sym_entry<T>(#P "spr_", K_PACKED, R_SYR, rm(o), uplo_c(u), n, alpha, x, incx, nullptr, 1, ap, 1); }                                         \
    void cblas_##P##spr2(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, T alpha, const T* x, int incx, const T* y, int incy, T* ap) {                     \
        rank_sym_entry<T>(#P "spr2_", K_PACKED, R_SYR2, rm(o), uplo_c(u), n, alpha, x, incx, y, incy, ap, 1); }
B200_REALSYM(s, float)
B200_REALSYM(d, double)
#undef B200_REALSYM

// ------------------------------- real, full storage: GER, SYR, SYMV, TRMV (reference blas_level2/ger.cc, syr.cc, symv.cc, trmv.cc) -------------------------------
#define B200_REALFULL(P, T)                                                                                                                             \
    void P##ger_(const int* m, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* a, const int* lda) {           \
        gerx_entry<T>(#P "ger_", false, false, *m, *n, alpha, x, *incx, y, *incy, a, *lda); }                                                            \
    void P##syr_(const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, T* a, const int* lda) {                                    \
        rank_sym_entry<T>(#P "syr_", K_FULL_TRI, R_SYR, false, uplo, *n, *alpha, x, *incx, nullptr, 1, a, *lda); }                                       \
    void P##symv_(const char* uplo, const int* n, const T* alpha, const T* a, const int* lda, const T* x, const int* incx, const T* beta, T* y,          \
                  const int* incy) { symv_like_entry<T>(#P "symv_", K_FULL_TRI, false, false, uplo, *n, 0, alpha, a, *lda, x, *incx, beta, y, *incy); }  \
    void P##trmv_(const char* uplo, const char* trans, const char* diag, const int* n, const T* a, const int* lda, T* x, const int* incx) {              \
        tri_entry<T>(#P "trmv_", K_FULL_TRI, false, false, uplo, trans, diag, *n, 0, a, *lda, x, *incx); }                                               \
    void cblas_##P##ger(enum CBLAS_ORDER o, int m, int n, T alpha, const T* x, int incx, const T* y, int incy, T* a, int lda) {                          \
        gerx_entry<T>(#P "ger_", false, rm(o), m, n, &alpha, x, incx, y, incy, a, lda); }                                                                \
    void cblas_##P##syr(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, T alpha, const T* x, int incx, T* a, int lda) {                                    \
        rank_sym_entry<T>(#P "syr_", K_FULL_TRI, R_SYR, rm(o), uplo_c(u), n, alpha, x, incx, nullptr, 1, a, lda); }                                      \
    void cblas_##P##symv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, T alpha, const T* a, int lda, const T* x, int incx, T beta, T* y, int incy) {     \
        symv_like_entry<T>(#P "symv_", K_FULL_TRI, false, rm(o), uplo_c(u), n, 0, &alpha, a, lda, x, incx, &beta, y, incy); }                            \
    void cblas_##P##trmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, enum CBLAS_TRANSPOSE t, enum CBLAS_DIAG d, int n, const T* a, int lda, T* x, int incx) { \
        tri_entry<T>(#P "trmv_", K_FULL_TRI, false, rm(o), uplo_c(u), trans_c(t), diag_c(d), n, 0, a, lda, x, incx); }
B200_REALFULL(s, float)
B200_REALFULL(d, double)
#undef B200_REALFULL

// ------------------------------- complex: HBMV, HPMV, HEMV, GERU, GERC, HER, HER2, HPR, HPR2, TRMV -------------------------------
#define B200_CPLX(P, T, CT, RT)                                                                                                                         \
    void P##hbmv_(const char* uplo, const int* n, const int* k, const T* alpha, const T* a, const int* lda, const T* x, const int* incx, const T* beta, T* y, \
                  const int* incy) {                                                                                                                     \
        symv_like_entry<CT>(#P "hbmv_", K_BAND_TRI, true, false, uplo, *n, *k, (const CT*)alpha, (const CT*)a, *lda, (const CT*)x, *incx, (const CT*)beta, (CT*)y, *incy); } \
    void P##hpmv_(const char* uplo, const int* n, const T* alpha, const T* ap, const T* x, const int* incx, const T* beta, T* y, const int* incy) {      \
        symv_like_entry<CT>(#P "hpmv_", K_PACKED, true, false, uplo, *n, 0, (const CT*)alpha, (const CT*)ap, 1, (const CT*)x, *incx, (const CT*)beta, (CT*)y, *incy); } \
    void P##hemv_(const char* uplo, const int* n, const T* alpha, const T* a, const int* lda, const T* x, const int* incx, const T* beta, T* y, const int* incy) { \
        symv_like_entry<CT>(#P "hemv_", K_FULL_TRI, true, false, uplo, *n, 0, (const CT*)alpha, (const CT*)a, *lda, (const CT*)x, *incx, (const CT*)beta, (CT*)y, *incy); } \
    void P##geru_(const int* m, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* a, const int* lda) {          \
        gerx_entry<CT>(#P "geru_", false, false, *m, *n, (const CT*)alpha, (const CT*)x, *incx, (const CT*)y, *incy, (CT*)a, *lda); }                    \
    void P##gerc_(const int* m, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* a, const int* lda) {          \
        gerx_entry<CT>(#P "gerc_", true, false, *m, *n, (const CT*)alpha, (const CT*)x, *incx, (const CT*)y, *incy, (CT*)a, *lda); }                     \
    void P##her_(const char* uplo, const int* n, const RT* alpha, const T* x, const int* incx, T* a, const int* lda) {                                   \
        rank_sym_entry<CT>(#P "her_", K_FULL_TRI, R_HER, false, uplo, *n, mkc(*alpha), (const CT*)x, *incx, nullptr, 1, (CT*)a, *lda); }                 \
    void P##her2_(const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* a, const int* lda) {      \
        rank_sym_entry<CT>(#P "her2_", K_FULL_TRI, R_HER2, false, uplo, *n, *(const CT*)alpha, (const CT*)x, *incx, (const CT*)y, *incy, (CT*)a, *lda); } \
    void P##hpr_(const char* uplo, const int* n, const RT* alpha, const T* x, const int* incx, T* ap) {                                                  \
        rank_sym_entry<CT>(#P "hpr_", K_PACKED, R_HER, false, uplo, *n, mkc(*alpha), (const CT*)x, *incx, nullptr, 1, (CT*)ap, 1); }                     \
    void P##hpr2_(const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* ap) {                     \
        rank_sym_entry<CT>(#P "hpr2_", K_PACKED, R_HER2, false, uplo, *n, *(const CT*)alpha, (const CT*)x, *incx, (const CT*)y, *incy, (CT*)ap, 1); }    \
    void P##trmv_(const char* uplo, const char* trans, const char* diag, const int* n, const T* a, const int* lda, T* x, const int* incx) {              \
        tri_entry<CT>(#P "trmv_", K_FULL_TRI, false, false, uplo, trans, diag, *n, 0, (const CT*)a, *lda, (CT*)x, *incx); }                              \
    void cblas_##P##hbmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, int k, const void* alpha, const void* a, int lda, const void* x, int incx,        \
                         const void* beta, void* y, int incy) {                                                                                          \
        symv_like_entry<CT>(#P "hbmv_", K_BAND_TRI, true, rm(o), uplo_c(u), n, k, (const CT*)alpha, (const CT*)a, lda, (const CT*)x, incx, (const CT*)beta, (CT*)y, incy); } \
    void cblas_##P##hpmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, const void* alpha, const void* ap, const void* x, int incx, const void* beta,     \
                         void* y, int incy) {                                                                                                            \
        symv_like_entry<CT>(#P "hpmv_", K_PACKED, true, rm(o), uplo_c(u), n, 0, (const CT*)alpha, (const CT*)ap, 1, (const CT*)x, incx, (const CT*)beta, (CT*)y, incy); } \
    void cblas_##P##hemv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, const void* alpha, const void* a, int lda, const void* x, int incx,               \
                         const void* beta, void* y, int incy) {                                                                                          \
        symv_like_entry<CT>(#P "hemv_", K_FULL_TRI, true, rm(o), uplo_c(u), n, 0, (const CT*)alpha, (const CT*)a, lda, (const CT*)x, incx, (const CT*)beta, (CT*)y, incy); } \
    void cblas_##P##geru(enum CBLAS_ORDER o, int m, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda) {      \
        gerx_entry<CT>(#P "geru_", false, rm(o), m, n, (const CT*)alpha, (const CT*)x, incx, (const CT*)y, incy, (CT*)a, lda); }                         \
    void cblas_##P##gerc(enum CBLAS_ORDER o, int m, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda) {      \
        gerx_entry<CT>(#P "gerc_", true, rm(o), m, n, (const CT*)alpha, (const CT*)x, incx, (const CT*)y, incy, (CT*)a, lda); }                          \
    void cblas_##P##her(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, RT alpha, const void* x, int incx, void* a, int lda) {                             \
        rank_sym_entry<CT>(#P "her_", K_FULL_TRI, R_HER, rm(o), uplo_c(u), n, mkc(alpha), (const CT*)x, incx, nullptr, 1, (CT*)a, lda); }                \
    void cblas_##P##her2(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* a, int lda) { \
        rank_sym_entry<CT>(#P "her2_", K_FULL_TRI, R_HER2, rm(o), uplo_c(u), n, *(const CT*)alpha, (const CT*)x, incx, (const CT*)y, incy, (CT*)a, lda); } \
    void cblas_##P##hpr(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, RT alpha, const void* x, int incx, void* ap) {                                     \
        rank_sym_entry<CT>(#P "hpr_", K_PACKED, R_HER, rm(o), uplo_c(u), n, mkc(alpha), (const CT*)x, incx, nullptr, 1, (CT*)ap, 1); }                   \
    void cblas_##P##hpr2(enum CBLAS_ORDER o, enum CBLAS_UPLO u, int n, const void* alpha, const void* x, int incx, const void* y, int incy, void* ap) {  \
        rank_sym_entry<CT>(#P "hpr2_", K_PACKED, R_HER2, rm(o), uplo_c(u), n, *(const CT*)alpha, (const CT*)x, incx, (const CT*)y, incy, (CT*)ap, 1); }  \
    void cblas_##P##trmv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, enum CBLAS_TRANSPOSE t, enum CBLAS_DIAG d, int n, const void* a, int lda, void* x, int incx) { \
        tri_entry<CT>(#P "trmv_", K_FULL_TRI, false, rm(o), uplo_c(u), trans_c(t), diag_c(d), n, 0, (const CT*)a, lda, (CT*)x, incx); }                  \
    /* column-major: the blocked TRSV of level2.cu; row-major (uplo flipped, ConjTrans as conjugate-no-transpose): the block solve above */                \
    void cblas_##P##trsv(enum CBLAS_ORDER o, enum CBLAS_UPLO u, enum CBLAS_TRANSPOSE t, enum CBLAS_DIAG d, int n, const void* a, int lda, void* x, int incx) { \
        if (!rm(o)) P##trsv_(uplo_c(u), trans_c(t), diag_c(d), &n, (const T*)a, &lda, (T*)x, &incx);                                                     \
        else tri_entry<CT>(#P "trsv_", K_FULL_TRI, true, true, uplo_c(u), trans_c(t), diag_c(d), n, 0, (const CT*)a, lda, (CT*)x, incx); }               \
    void cblas_##P##gemv(enum CBLAS_ORDER o, enum CBLAS_TRANSPOSE t, int m, int n, const void* alpha, const void* a, int lda, const void* x, int incx,   \
                         const void* beta, void* y, int incy) {                                                                                          \
        const char* tc = trans_c(t);                                                                                                                     \
        if (!rm(o)) { P##gemv_(tc, &m, &n, (const T*)alpha, (const T*)a, &lda, (const T*)x, &incx, (const T*)beta, (T*)y, &incy); return; }             \
        if (*tc != 'C') { const char* t2 = *tc == 'N' ? "T" : (*tc == 'T' ? "N" : "?");                                                                  \
            P##gemv_(t2, &n, &m, (const T*)alpha, (const T*)a, &lda, (const T*)x, &incx, (const T*)beta, (T*)y, &incy); return; }                        \
        /* row-major A^H x: conj(S) x on the column-major n x m view S; netlib xGEMV checks in the view's terms (info 2,3,6,8,11) */                       \
        int info = 0;                                                                                                                                    \
        if (n < 0) info = 2; else if (m < 0) info = 3; else if (lda < imax(1, n)) info = 6; else if (incx == 0) info = 8; else if (incy == 0) info = 11; \
        if (info) { call_xerbla(#P "gemv_", info); return; }                                                                                             \
        const CT al = *(const CT*)alpha, be = *(const CT*)beta;                                                                                          \
        if (m == 0 || n == 0 || (is0(al) && is1(be))) return;                                                                                            \
        if (is0(al)) { P##gemv_("N", &n, &m, (const T*)alpha, (const T*)a, &lda, (const T*)x, &incx, (const T*)beta, (T*)y, &incy); return; }           \
        gemv_conj_notrans<CT>(#P "gemv_", n, m, &al, (const CT*)a, lda, (const CT*)x, incx, &be, (CT*)y, incy); }
B200_CPLX(c, b200_c32, c32, float)
B200_CPLX(z, b200_c64, c64, double)
#undef B200_CPLX

}  // extern "C"
