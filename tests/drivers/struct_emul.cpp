// struct_emul.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles libgpublas_b200/csrc/structured.cuh (the index logic and the plans of the banded / packed / Hermitian Level-2
// routines) with g++ and supplies a backend that walks the kernels' (block, thread) grids on the CPU, lane by lane, in the
// launch shapes level2_struct.cu uses.  tests/test_level2_struct_cpu.py checks its results against the oracle, so the
// storage arithmetic, the flag choices per routine and the row-major mappings are verified where there is no GPU; the CUDA
// kernels themselves are verified on the GPU by tests/test_zz_level2_struct_gpu.py.  libb200blas.so never loads this file.
#include "../../libgpublas_b200/csrc/structured.cuh"

#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

using namespace b200::st;

namespace {

template <typename T> T butterfly(T* v) {   // the warp_sum of level2_struct.cu: xor offsets 16, 8, 4, 2, 1; lane 0's value
    T t[32];
    for (int o = 16; o > 0; o >>= 1) {
        for (int l = 0; l < 32; l++) t[l] = el<T>::add(v[l], v[l ^ o]);
        std::memcpy(v, t, sizeof t);
    }
    return v[0];
}

struct HostBackend {
    std::vector<std::unique_ptr<char[]>> pool;
    void* alloc(size_t bytes) { pool.emplace_back(new char[bytes ? bytes : 1]); return pool.back().get(); }
    int sm_target() const { return 148 * 8; }
    template <typename T> void gather(int n, const T* src, int64_t inc, T* dst, bool conj) {
        for (int bx = 0; bx < (n + 255) / 256; bx++)
            for (int tx = 0; tx < 256; tx++) {
                const int i = bx * 256 + tx;
                if (i >= n) continue;
                T v = src[vpos(i, n, inc)];
                dst[i] = conj ? el<T>::conj(v) : v;
            }
    }
    template <typename T> void scatter(int n, const T* src, T* dst, int64_t inc) {
        for (int i = 0; i < n; i++) dst[vpos(i, n, inc)] = src[i];
    }
    template <typename T>
    void npart(const Desc& D, const T* A, const T* v, int row0, int row1, int c_lo, int c_hi, int cpc, int nchunks, int flags, T* part, int64_t npad) {
        for (int by = 0; by < nchunks; by++)
            for (int bx = 0; bx < (row1 - row0 + ROW_THREADS - 1) / ROW_THREADS; bx++)
                for (int tx = 0; tx < ROW_THREADS; tx++) {
                    const int i = row0 + bx * ROW_THREADS + tx;
                    if (i >= row1) continue;
                    const int c0 = c_lo + by * cpc, c1 = st_min(c_hi, c0 + cpc);
                    part[(int64_t)by * npad + i] = npart_row<T>(D, A, v, i, c0, c1, flags, i - (tx & 31));
                }
    }
    template <typename T> void tpart(const Desc& D, const T* A, const T* v, int col0, int col1, int r0, int r1, int flags, T* tp) {
        for (int bx = 0; bx < (col1 - col0 + COL_WARPS - 1) / COL_WARPS; bx++)
            for (int w = 0; w < COL_WARPS; w++) {
                const int j = col0 + bx * COL_WARPS + w;
                if (j >= col1) continue;
                T lanes[32];
                for (int lane = 0; lane < 32; lane++) lanes[lane] = tpart_lane<T>(D, A, v, j, lane, 32, r0, r1, flags);
                tp[j] = butterfly(lanes);
            }
    }
    template <typename T>
    void finish(int n, int nparts, const T* part, int64_t npad, const T* tp, const T* vunit, T alpha, T beta, T* out, int64_t inco) {
        for (int i = 0; i < n; i++) {
            T* p = out + vpos(i, n, inco);
            const bool beta0 = el<T>::is_zero(beta);
            *p = finish_elem<T>(i, nparts, part, npad, tp, vunit, alpha, beta, beta0, beta0 ? el<T>::zero() : *p);
        }
    }
    // one-pass symmetric product: the grid and the per-warp strips of sympart_kernel; the lanes of a warp run one after the other, so
    // the sink adds into the strip where the device's butterfly stores one finished sum (the butterfly itself is checked on the GPU)
    template <typename T> struct HostStripSink {
        enum { NU = unroll_of<T>::N };
        T* strip; int cw0, cw1;
        void step(int j, T (&t)[NU]) {
            for (int u = 0; u < NU; u++)
                if (j + u >= cw0 && j + u < cw1) strip[j + u - cw0] = el<T>::add(strip[j + u - cw0], t[u]);
        }
    };
    int sym_two_pass = 0;
    template <typename T> int sym_max_cols() const { return sym_two_pass ? 0 : (int)(8192 / sizeof(T)); }
    template <typename T>
    void sympart(const Desc& D, const T* A, const T* v, int cpc, int nchunks, int nflags, int tflags, T* part, int64_t npad, T* tp2, int64_t npadw) {
        const int64_t w = sym_width(D) < cpc ? sym_width(D) : (int64_t)cpc;
        const int wstride = (int)((w + 7) / 8 * 8), nwarps = ROW_THREADS / 32;
        std::vector<T> strips((size_t)nwarps * wstride);
        for (int by = 0; by < nchunks; by++)
            for (int bx = 0; bx < (D.n + ROW_THREADS - 1) / ROW_THREADS; bx++) {
                const int r0 = bx * ROW_THREADS, c0 = by * cpc, c1 = st_min(D.n, c0 + cpc);
                int cw0, cw1;
                sym_window(D, r0, c0, c1, cw0, cw1);
                if (cw1 - cw0 > wstride) std::abort();                     // the strip must hold the CTA's window
                for (auto& x : strips) x = el<T>::zero();
                for (int tx = 0; tx < ROW_THREADS; tx++) {
                    const int i = r0 + tx, lane = tx & 31, warp = tx >> 5;
                    HostStripSink<T> sink = {strips.data() + (size_t)warp * wstride, cw0, cw1};
                    const T r = sym_row<T>(D, A, v, v, i, i - lane, c0, c1, nflags, tflags, sink);
                    if (i < D.n) part[(int64_t)by * npad + i] = r;
                }
                T* row = tp2 + (int64_t)bx * npadw - sym_jw0(D, r0);
                for (int t = 0; t < cw1 - cw0; t++) {
                    T sacc = strips[t];
                    for (int wv = 1; wv < nwarps; wv++) sacc = el<T>::add(sacc, strips[(size_t)wv * wstride + t]);
                    if (cw0 + t - sym_jw0(D, r0) < 0 || cw0 + t - sym_jw0(D, r0) >= npadw) std::abort();
                    row[cw0 + t] = sacc;
                }
            }
    }
    template <typename T> void sym_finish(const Desc& D, int nparts, const T* part, int64_t npad, const T* tp2, int64_t npadw, T alpha, T beta, T* out, int64_t inco) {
        for (int j = 0; j < D.n; j++) {
            T* p = out + vpos(j, D.n, inco);
            const bool beta0 = el<T>::is_zero(beta);
            *p = sym_finish_elem<T>(D, j, nparts, part, npad, tp2, npadw, alpha, beta, beta0, beta0 ? el<T>::zero() : *p);
        }
    }
    template <typename T> void rank(const Desc& D, T* A, int rows, int ncols, int cpc, int nchunks, T alpha, const T* x, const T* y, int mode) {
        for (int by = 0; by < nchunks; by++)
            for (int bx = 0; bx < (rows + ROW_THREADS - 1) / ROW_THREADS; bx++)
                for (int tx = 0; tx < ROW_THREADS; tx++) {
                    const int i = bx * ROW_THREADS + tx;
                    if (i >= rows) continue;
                    const int c0 = by * cpc, c1 = st_min(ncols, c0 + cpc);
                    rank_row<T>(D, A, i, c0, c1, alpha, x, y, mode, i - (tx & 31));
                }
    }
    // the warp of solve_diag_warp, lane by lane in lock step
    template <typename T> void solve_diag(const Desc& D, const T* A, T* x, int b0, int nb, bool trans, bool conj, bool unit, bool forward) {
        T coef[32][32], dinv[32], xv[32];
        for (int lane = 0; lane < 32; lane++) {
            const int r = b0 + lane;
            dinv[lane] = el<T>::one();
            for (int step = 0; step < 32; step++) {
                const int jj = forward ? step : nb - 1 - step;
                coef[lane][step] = el<T>::zero();
                if (step < nb && lane < nb) {
                    const bool waiting = forward ? lane > jj : lane < jj;
                    T a;
                    if ((waiting || (lane == jj && !unit)) && solve_coef<T>(D, A, r, b0 + jj, trans, conj, a)) {
                        if (lane == jj) dinv[lane] = el<T>::div(el<T>::one(), a); else coef[lane][step] = a;
                    }
                }
            }
            xv[lane] = lane < nb ? x[r] : el<T>::zero();
        }
        for (int step = 0; step < 32; step++) {
            if (step >= nb) continue;
            const int jj = forward ? step : nb - 1 - step;
            if (!unit) xv[jj] = el<T>::mul(xv[jj], dinv[jj]);
            const T xj = xv[jj];
            for (int lane = 0; lane < 32; lane++) {
                const bool waiting = forward ? (lane > jj && lane < nb) : lane < jj;
                if (waiting) xv[lane] = el<T>::sub(xv[lane], el<T>::mul(coef[lane][step], xj));
            }
        }
        for (int lane = 0; lane < nb; lane++) x[b0 + lane] = xv[lane];
    }
    // solve_panel_kernel: the CTA's phases in order; within the update phase every thread reads only solved unknowns of the
    // block and writes only its own rows, so the thread order does not matter
    template <typename T> void solve_panel(const Desc& D, const T* A, T* x, int p0, int p1, bool trans, bool conj, bool unit, bool forward) {
        const int nblk = (p1 - p0 + SOLVE_NB - 1) / SOLVE_NB, flags = conj ? F_CONJ : 0;
        for (int bi = 0; bi < nblk; bi++) {
            int b0, b1, u0, u1;
            panel_block(D, p0, p1, forward, bi, b0, b1, u0, u1);
            solve_diag<T>(D, A, x, b0, b1 - b0, trans, conj, unit, forward);
            for (int tx = 0; tx < SOLVE_THREADS; tx++)
                for (int r = u0 + tx; r < u1; r += SOLVE_THREADS) x[r] = el<T>::sub(x[r], panel_update<T>(D, A, x, r, b0, b1, trans, flags));
        }
    }
    template <typename T> void solve_nupdate(const Desc& D, const T* A, T* x, int row0, int row1, int b0, int b1, int flags) {
        for (int bx = 0; bx < (row1 - row0 + ROW_THREADS - 1) / ROW_THREADS; bx++)
            for (int tx = 0; tx < ROW_THREADS; tx++) {
                const int i = row0 + bx * ROW_THREADS + tx;
                if (i < row1) x[i] = el<T>::sub(x[i], npart_row<T>(D, A, x, i, b0, b1, flags, i - (tx & 31)));
            }
    }
    template <typename T> void solve_tupdate(const Desc& D, const T* A, T* x, int col0, int col1, int b0, int b1, int flags) {
        for (int j = col0; j < col1; j++) {
            T lanes[32];
            for (int lane = 0; lane < 32; lane++) lanes[lane] = tpart_lane<T>(D, A, x, j, lane, 32, b0, b1, flags);
            x[j] = el<T>::sub(x[j], butterfly(lanes));
        }
    }
};

inline cuFloatComplex mkc(float r) { cuFloatComplex c; c.x = r; c.y = 0; return c; }
inline cuDoubleComplex mkc(double r) { cuDoubleComplex c; c.x = r; c.y = 0; return c; }

}  // namespace

#define EMU(P, T)                                                                                                                                      \
    extern "C" void emu_##P##gbmv(int rowmajor, char trans, int m, int n, int kl, int ku, const T* alpha, const T* a, int lda, const T* x, int incx,  \
                                  const T* beta, T* y, int incy) {                                                                                     \
        HostBackend be; plan_gbmv<T>(be, rowmajor != 0, trans, m, n, kl, ku, *alpha, a, lda, x, incx, *beta, y, incy); }                               \
    extern "C" void emu_##P##symv_like(int kind, int herm, int rowmajor, char uplo, int n, int k, const T* alpha, const T* a, int lda, const T* x,     \
                                       int incx, const T* beta, T* y, int incy) {                                                                      \
        HostBackend be; plan_symv_like<T>(be, kind, herm != 0, rowmajor != 0, uplo == 'U', n, k, *alpha, a, kind == K_PACKED ? packed_len(n) : lda, x, incx, *beta, y, incy); } \
    extern "C" void emu_##P##tri(int kind, int solve, int rowmajor, char uplo, char trans, char diag, int n, int k, const T* a, int lda, T* x, int incx) { \
        HostBackend be; plan_tri<T>(be, kind, solve != 0, rowmajor != 0, uplo == 'U', trans, diag == 'U', n, k, a, kind == K_PACKED ? packed_len(n) : lda, x, incx); } \
    extern "C" void emu_##P##ger(int conjy, int rowmajor, int m, int n, const T* alpha, const T* x, int incx, const T* y, int incy, T* a, int lda) {   \
        HostBackend be; plan_ger<T>(be, conjy != 0, rowmajor != 0, m, n, *alpha, x, incx, y, incy, a, lda); }                                          \
    extern "C" void emu_##P##rank_sym(int kind, int mode, int rowmajor, char uplo, int n, const T* alpha, const T* x, int incx, const T* y, int incy,  \
                                      T* a, int lda) {                                                                                                 \
        HostBackend be; plan_rank_sym<T>(be, kind, mode, rowmajor != 0, uplo == 'U', n, *alpha, x, incx, y, incy, a, kind == K_PACKED ? packed_len(n) : lda); } \
    extern "C" void emu_##P##gemv_conj(int m, int n, const T* alpha, const T* a, int lda, const T* x, int incx, const T* beta, T* y, int incy) {       \
        HostBackend be; plan_gemv_conj<T>(be, m, n, *alpha, a, lda, x, incx, *beta, y, incy); }
EMU(s, float)
EMU(d, double)
EMU(c, cuFloatComplex)
EMU(z, cuDoubleComplex)
