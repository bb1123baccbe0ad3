// structured.cuh -- index logic shared by the banded / packed / Hermitian / complex Level-2 kernels (level2_struct.cu).
//
// the reference's dead wrappers blas_level2/{gbmv,bmv,pmv,hemv,her,her2,hpr,hpr2,spr,spr2,
// syr2,tbmv,tbsv,tpmv,tpsv,ger,trmv}.cc name these routines and forward them to cublas<t>{gbmv,sbmv,hbmv,spmv,hpmv,...}.
// Here every storage scheme is one descriptor (Desc) with three questions -- which rows of column j are stored, which
// columns of row i are stored, where does element (i,j) live -- and every routine is built from three per-thread /
// per-lane bodies over it:
//
//   npart_row   r(i)  = sum_j op(S(i,j)) v(j)   one thread per row i; adjacent threads read adjacent addresses (every
//                                                scheme is contiguous down a column), so the pass is coalesced
//   tpart_lane  r(j)  = sum_i op(S(i,j)) v(i)   one warp per column j, lanes stride the stored rows (coalesced), butterfly sum
//   rank_row    S(i,j) += a x(i) op(y(j)) [+ ...] one thread per row i over the stored columns
//
// Everything here is __host__ __device__ and free of CUDA runtime calls: tests/drivers/struct_emul.cpp compiles this same
// header with g++ and walks the (block, thread) grid on the CPU to check the index logic against the oracle where there
// is no GPU.  That harness is test infrastructure; the product library contains no host execution path for these routines.
#pragma once
#include <cuComplex.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ST_HD __host__ __device__ __forceinline__
#else
#define ST_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define ST_UNROLL _Pragma("unroll")     // device pass only: the host pass of these __host__ __device__ bodies does not know the pragma
#else
#define ST_UNROLL
#endif

namespace b200 {
namespace st {

enum Kind {
    K_FULL_GEN = 0,   // m x n, every element stored, leading dimension ld                       (GERU / GERC, CBLAS row-major GEMV)
    K_FULL_TRI = 1,   // n x n in full storage, one triangle referenced                           (HEMV, HER, HER2, SYR2, complex TRMV)
    K_BAND_GEN = 2,   // m x n with kl sub- and ku super-diagonals: S(i,j) at AB[ku+i-j + j*ld]   (GBMV)
    K_BAND_TRI = 3,   // n x n triangle with k off-diagonals: upper AB[k+i-j + j*ld], lower AB[i-j + j*ld]   (SBMV, HBMV, TBMV, TBSV)
    K_PACKED = 4      // n x n triangle packed by columns: upper AP[i + j(j+1)/2], lower AP[i + j(2n-j-1)/2]  (SPMV, HPMV, SPR*, HPR*, TPMV, TPSV)
};

// launch shapes shared by the kernels (level2_struct.cu) and their CPU emulation (tests/drivers/struct_emul.cpp)
enum {
    ROW_THREADS = 128,   // N part / rank row / solve update: one thread per row, this many rows per CTA
    COL_WARPS = 4        // T part / transposed solve update: one warp per column, this many columns per CTA
};

struct Desc {
    int kind, m, n, kl, ku, upper;   // K_BAND_TRI keeps its k in both kl and ku
    int64_t ld;
};

ST_HD int st_min(int a, int b) { return a < b ? a : b; }
ST_HD int st_max(int a, int b) { return a > b ? a : b; }

// stored rows [i0, i1) of column j
ST_HD void col_rows(const Desc& D, int j, int& i0, int& i1) {
    switch (D.kind) {
        case K_FULL_GEN: i0 = 0; i1 = D.m; break;
        case K_BAND_GEN: i0 = st_max(0, j - D.ku); i1 = st_min(D.m, j + D.kl + 1); break;
        case K_BAND_TRI:
            if (D.upper) { i0 = st_max(0, j - D.ku); i1 = j + 1; } else { i0 = j; i1 = st_min(D.n, j + D.kl + 1); }
            break;
        default:   // K_FULL_TRI, K_PACKED
            if (D.upper) { i0 = 0; i1 = j + 1; } else { i0 = j; i1 = D.n; }
            break;
    }
}
// stored columns [j0, j1) of row i
ST_HD void row_cols(const Desc& D, int i, int& j0, int& j1) {
    switch (D.kind) {
        case K_FULL_GEN: j0 = 0; j1 = D.n; break;
        case K_BAND_GEN: j0 = st_max(0, i - D.kl); j1 = st_min(D.n, i + D.ku + 1); break;
        case K_BAND_TRI:
            if (D.upper) { j0 = i; j1 = st_min(D.n, i + D.ku + 1); } else { j0 = st_max(0, i - D.kl); j1 = i + 1; }
            break;
        default:
            if (D.upper) { j0 = i; j1 = D.n; } else { j0 = 0; j1 = i + 1; }
            break;
    }
}
ST_HD bool stored(const Desc& D, int i, int j) {
    int i0, i1;
    col_rows(D, j, i0, i1);
    return i >= i0 && i < i1;
}
// offset (in elements) of stored element (i,j)
ST_HD int64_t off(const Desc& D, int i, int j) {
    switch (D.kind) {
        case K_BAND_GEN: return (int64_t)(D.ku + i - j) + (int64_t)j * D.ld;
        case K_BAND_TRI: return (int64_t)(D.upper ? D.ku + i - j : i - j) + (int64_t)j * D.ld;
        case K_PACKED:
            return D.upper ? (int64_t)i + (int64_t)j * ((int64_t)j + 1) / 2 : (int64_t)i + (int64_t)j * (2 * (int64_t)D.n - j - 1) / 2;
        default: return (int64_t)i + (int64_t)j * D.ld;
    }
}
// furthest |i-j| of a stored element (bounds how far a solved block reaches in the triangular solves)
ST_HD int reach(const Desc& D) { return (D.kind == K_BAND_GEN || D.kind == K_BAND_TRI) ? st_max(D.kl, D.ku) : st_max(D.m, D.n); }

// ---- arithmetic on the four element types ----
template <typename T> struct el;
template <> struct el<float> {
    typedef float real;
    static ST_HD float zero() { return 0.f; }
    static ST_HD float one() { return 1.f; }
    static ST_HD float conj(float a) { return a; }
    static ST_HD float realpart(float a) { return a; }
    static ST_HD float mul(float a, float b) { return a * b; }
    static ST_HD float mad(float a, float b, float c) { return a * b + c; }
    static ST_HD float add(float a, float b) { return a + b; }
    static ST_HD float sub(float a, float b) { return a - b; }
    static ST_HD float div(float a, float b) { return a / b; }
    static ST_HD float scale(float r, float a) { return r * a; }
    static ST_HD bool is_zero(float a) { return a == 0.f; }
};
template <> struct el<double> {
    typedef double real;
    static ST_HD double zero() { return 0.0; }
    static ST_HD double one() { return 1.0; }
    static ST_HD double conj(double a) { return a; }
    static ST_HD double realpart(double a) { return a; }
    static ST_HD double mul(double a, double b) { return a * b; }
    static ST_HD double mad(double a, double b, double c) { return a * b + c; }
    static ST_HD double add(double a, double b) { return a + b; }
    static ST_HD double sub(double a, double b) { return a - b; }
    static ST_HD double div(double a, double b) { return a / b; }
    static ST_HD double scale(double r, double a) { return r * a; }
    static ST_HD bool is_zero(double a) { return a == 0.0; }
};
template <typename C, typename Rl> struct el_cplx {
    typedef Rl real;
    static ST_HD C mk(Rl x, Rl y) { C c; c.x = x; c.y = y; return c; }
    static ST_HD C zero() { return mk(0, 0); }
    static ST_HD C one() { return mk(1, 0); }
    static ST_HD C conj(C a) { return mk(a.x, -a.y); }
    static ST_HD C realpart(C a) { return mk(a.x, 0); }
    static ST_HD C mul(C a, C b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
    static ST_HD C mad(C a, C b, C c) { return mk(c.x + (a.x * b.x - a.y * b.y), c.y + (a.x * b.y + a.y * b.x)); }
    static ST_HD C add(C a, C b) { return mk(a.x + b.x, a.y + b.y); }
    static ST_HD C sub(C a, C b) { return mk(a.x - b.x, a.y - b.y); }
    // netlib divides complex numbers with the compiler's complex division; Smith's form avoids spurious overflow likewise
    static ST_HD C div(C a, C b) {
        Rl abx = b.x < 0 ? -b.x : b.x, aby = b.y < 0 ? -b.y : b.y;
        if (abx >= aby) { Rl r = b.y / b.x, d = b.x + b.y * r; return mk((a.x + a.y * r) / d, (a.y - a.x * r) / d); }
        Rl r = b.x / b.y, d = b.x * r + b.y;
        return mk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
    }
    static ST_HD C scale(Rl r, C a) { return mk(r * a.x, r * a.y); }
    static ST_HD bool is_zero(C a) { return a.x == 0 && a.y == 0; }
};
template <> struct el<cuFloatComplex> : el_cplx<cuFloatComplex, float> {};
template <> struct el<cuDoubleComplex> : el_cplx<cuDoubleComplex, double> {};

// ---- flags of the matrix-vector bodies ----
enum {
    F_CONJ = 1,     // use conj(S(i,j))
    F_NODIAG = 2,   // skip i == j (unit-diagonal triangular products; the mirrored half of a symmetric product)
    F_HERM = 4      // the diagonal of a Hermitian matrix: imaginary part taken as zero (netlib xHEMV/xHBMV/xHPMV use DBLE(A(j,j)))
};

template <typename T> ST_HD T load_elem(const Desc& D, const T* A, int i, int j, int flags) {
    T a = A[off(D, i, j)];
    if ((flags & F_HERM) && i == j) a = el<T>::realpart(a);
    if (flags & F_CONJ) a = el<T>::conj(a);
    return a;
}

// offset(i, j+1) - offset(i, j): the walk along a row is a pointer bump (no 64-bit multiply per element)
ST_HD int64_t col_step(const Desc& D, int j) {
    switch (D.kind) {
        case K_BAND_GEN:
        case K_BAND_TRI: return D.ld - 1;
        case K_PACKED: return D.upper ? (int64_t)j + 1 : (int64_t)D.n - j - 1;
        default: return D.ld;
    }
}
// acc + op(a) * w with the flags applied to the loaded element a = S(i,j).  A skipped diagonal takes no part at all: its stored
// value (rogue, NaN) is dropped and no 0 * w product is formed, so an Inf in the vector cannot turn into a NaN (netlib never
// multiplies a unit diagonal either).
template <typename T> ST_HD T mad_elem(T a, T w, int i, int j, int flags, T acc) {
    if (i == j) {
        if (flags & F_NODIAG) return acc;
        if (flags & F_HERM) a = el<T>::realpart(a);
    }
    return el<T>::mad((flags & F_CONJ) ? el<T>::conj(a) : a, w, acc);
}
// r(i) = sum over the stored columns j of row i inside [c0, c1) of op(S(i,j)) * v(j).
// Memory behaviour is the point of this body (one thread per row):
//  * the threads of a warp must read the SAME column in the same step, or their addresses are a leading dimension apart:
//    the walk starts at the first stored column of the warp's first row (i_first; the start column never decreases with
//    the row) and a thread simply masks the columns left of its own range (upper triangles, bands);
//  * NU columns per step, all loads issued before the arithmetic: with one thread per row the bytes in flight per SM set
//    the bandwidth (one load per step: 24-33 % of the HBM peak, profiles/r01e_level2_struct_perf_v1.txt).
template <typename T> struct unroll_of { enum { N = sizeof(T) >= 16 ? 4 : 8 }; };
template <typename T> ST_HD T npart_row(const Desc& D, const T* A, const T* v, int i, int c0, int c1, int flags, int i_first) {
    enum { NU = unroll_of<T>::N };
    int j0, j1, jf0, jf1;
    row_cols(D, i, j0, j1);
    j0 = st_max(j0, c0); j1 = st_min(j1, c1);
    row_cols(D, i_first, jf0, jf1);
    const int jstart = st_min(j0, st_max(jf0, c0));
    T acc[NU];
ST_UNROLL
    for (int u = 0; u < NU; u++) acc[u] = el<T>::zero();
    if (j0 >= j1) return acc[0];
    const bool packed = D.kind == K_PACKED, up = D.upper != 0, cj = (flags & F_CONJ) != 0;
    const int64_t cs = packed ? 0 : col_step(D, 0);
    const T* p = A + off(D, i, jstart);   // may point outside the stored part for columns left of j0: never dereferenced there
    for (int j = jstart; j < j1; j += NU) {
        T a[NU], w[NU];
        int64_t s = 0;
        if (j >= j0 && j + NU <= j1) {   // interior step: no masks, a constant stride where the scheme has one
            if (packed) {
ST_UNROLL
                for (int u = 0; u < NU; u++) { a[u] = p[s]; s += up ? (int64_t)j + u + 1 : (int64_t)D.n - j - u - 1; }
            } else {
ST_UNROLL
                for (int u = 0; u < NU; u++) a[u] = p[u * cs];
                s = NU * cs;
            }
ST_UNROLL
            for (int u = 0; u < NU; u++) w[u] = v[j + u];
            p += s;
            if (i < j || i >= j + NU) {  // ... and no diagonal element in it: no flag tests either
ST_UNROLL
                for (int u = 0; u < NU; u++) acc[u] = el<T>::mad(cj ? el<T>::conj(a[u]) : a[u], w[u], acc[u]);
                continue;
            }
        } else {                         // head (columns left of this row's range) or tail
ST_UNROLL
            for (int u = 0; u < NU; u++) {
                const bool in = j + u >= j0 && j + u < j1;
                a[u] = in ? p[s] : el<T>::zero();
                w[u] = in ? v[j + u] : el<T>::zero();
                s += col_step(D, j + u);
            }
            p += s;
        }
ST_UNROLL
        for (int u = 0; u < NU; u++) acc[u] = mad_elem<T>(a[u], w[u], i, j + u, flags, acc[u]);
    }
    T r = acc[0];
ST_UNROLL
    for (int u = 1; u < NU; u++) r = el<T>::add(r, acc[u]);
    return r;
}
// this lane's share of r(j) = sum over the stored rows i of column j inside [r0, r1) of op(S(i,j)) * v(i); the stored rows
// of a column are contiguous in every scheme, so the lane walks a pointer in steps of nlanes, four loads in flight
template <typename T> ST_HD T tpart_lane(const Desc& D, const T* A, const T* v, int j, int lane, int nlanes, int r0, int r1, int flags) {
    int i0, i1;
    col_rows(D, j, i0, i1);
    i0 = st_max(i0, r0); i1 = st_min(i1, r1);
    T acc0 = el<T>::zero(), acc1 = el<T>::zero(), acc2 = el<T>::zero(), acc3 = el<T>::zero();
    int i = i0 + lane;
    if (i >= i1) return acc0;
    const T* p = A + off(D, i, j);
    for (; (int64_t)i + 3 * (int64_t)nlanes < i1; i += 4 * nlanes, p += 4 * nlanes) {
        const T a0 = p[0], a1 = p[nlanes], a2 = p[2 * nlanes], a3 = p[3 * nlanes];
        const T v0 = v[i], v1 = v[i + nlanes], v2 = v[i + 2 * nlanes], v3 = v[i + 3 * nlanes];
        acc0 = mad_elem<T>(a0, v0, i, j, flags, acc0);
        acc1 = mad_elem<T>(a1, v1, i + nlanes, j, flags, acc1);
        acc2 = mad_elem<T>(a2, v2, i + 2 * nlanes, j, flags, acc2);
        acc3 = mad_elem<T>(a3, v3, i + 3 * nlanes, j, flags, acc3);
    }
    for (; i < i1; i += nlanes, p += nlanes) acc0 = mad_elem<T>(*p, v[i], i, j, flags, acc0);
    return el<T>::add(el<T>::add(acc0, acc1), el<T>::add(acc2, acc3));
}
// ---- symmetric / Hermitian products in ONE pass over the stored triangle ----
// The two-pass form (N part over the triangle, T part over its strict half) reads every stored element twice: at most 50 % of
// the HBM peak (measured 41-44 %, profiles/r01g_level2_struct_summary.txt).  Here a CTA of ROW_THREADS rows x one column chunk
// walks its rows exactly like npart_row, and every loaded element S(i,j) also feeds the mirrored product op_t(S(i,j)) * v(i)
// for out(j).  The 32 rows of a warp hold NU such products per step, one per column: a transposing butterfly (sym_step in
// level2_struct.cu) leaves each column's sum over the warp's rows in one lane, which stores it in the warp's strip of shared
// memory -- every column is visited by exactly one step of a warp, so the strip needs no read-modify-write.  At the end the
// CTA adds its warps' strips and writes them to tp2[row block][column - jw0(row block)]; sym_finish_elem adds, for out(j), the
// row blocks that hold stored rows of column j.  Extra traffic: tp2 is written and read once, n/ROW_THREADS values per column
// against ~n/2 stored ones (1.6 %).  Deterministic: fixed grid, fixed order.
// first column of the tp2 row of the row block that starts at row r0 (bands keep only the block's window of columns)
ST_HD int sym_jw0(const Desc& D, int r0) {
    if (D.kind != K_BAND_TRI) return 0;
    int j0, j1;
    row_cols(D, r0, j0, j1);
    return j0;
}
// width of a tp2 row: the whole matrix, or (bands) the columns a row block can touch
ST_HD int64_t sym_width(const Desc& D) { return D.kind == K_BAND_TRI ? (int64_t)ROW_THREADS + st_max(D.kl, D.ku) : (int64_t)D.n; }
// columns [cw0, cw1) of chunk [c0, c1) that the rows [r0, r0 + ROW_THREADS) of an n x n triangle can touch (row ranges never move left as i grows)
ST_HD void sym_window(const Desc& D, int r0, int c0, int c1, int& cw0, int& cw1) {
    const int rl = st_min(r0 + ROW_THREADS, D.n) - 1;
    int ja0, ja1, jb0, jb1;
    row_cols(D, r0, ja0, ja1);
    row_cols(D, rl, jb0, jb1);
    cw0 = st_max(c0, ja0); cw1 = st_min(c1, jb1);
    if (cw1 < cw0) cw1 = cw0;
}
// One row's walk.  Every lane of a warp runs the SAME steps (the butterfly in sink.step needs all 32): the loop bounds come from
// the warp's first and last row, a lane masks the columns outside its own range (rows past n: everything).  Returns the N part
// of row i; sink.step(j, t) receives the NU mirrored products of the step, t[u] = op_t(S(i, j+u)) * v(i) (zero where masked).
// vcol[j] = v[j] for the columns of the CTA's window (the kernel keeps them in shared memory: the per-step v(j) loads were the
// largest stall of the first two versions -- long scoreboard 6.3 of 11 cycles per issued instruction, profiles/r02s_ncu_full_raw.csv).
template <typename T, typename SINK> ST_HD T sym_row(const Desc& D, const T* A, const T* v, const T* vcol, int i, int i_first, int c0, int c1, int nflags, int tflags, SINK& sink) {
    enum { NU = unroll_of<T>::N };
    const int n = D.n, i_last = st_min(i_first + 31, n - 1);
    if (i_first >= n) return el<T>::zero();                  // the whole warp is past the last row
    int j0 = 0, j1 = 0, jf0, jf1, jl0, jl1;
    if (i < n) row_cols(D, i, j0, j1);
    j0 = st_max(j0, c0); j1 = st_min(j1, c1);
    row_cols(D, i_first, jf0, jf1);
    row_cols(D, i_last, jl0, jl1);
    const int jstart = st_max(jf0, c0), jend = st_min(jl1, c1);
    // INTERIOR steps (warp-uniform test): all 32 rows exist, the step lies inside every row's range and meets no diagonal.  They
    // are nearly all of the work and run without masks, flag tests or per-column stride look-ups, with the next step's matrix
    // elements already in flight while this step's butterfly runs (the first version did all of that per element: 215 issue slots
    // per step, 57 % of the HBM peak with the DRAM pipe half idle; profiles/r02p_level2_sym_one_pass_v1.txt).
    const bool whole = i_first + 31 < n;
    const int jin0 = st_max(jl0, c0), jin1 = st_min(jf1, c1);          // the last row starts last, the first row ends first
    const bool packed = D.kind == K_PACKED, up = D.upper != 0;
    const int64_t cs = packed ? 0 : col_step(D, 0);
    const bool cn = (nflags & F_CONJ) != 0, ct = (tflags & F_CONJ) != 0;
    T acc[NU];
ST_UNROLL
    for (int u = 0; u < NU; u++) acc[u] = el<T>::zero();
    const T vi = i < n ? v[i] : el<T>::zero();
    const T* p = A + off(D, i < n ? i : i_first, jstart);   // outside the stored part it is never dereferenced
    int j = jstart;
    while (j < jend) {
        T a[NU], w[NU], t[NU];
        const bool interior = whole && j >= jin0 && j + NU <= jin1 && (j + NU <= i_first || j > i_last);
        if (interior) {
            T an[NU];
            auto fetch = [&](T (&dst)[NU], int jj) {       // NU loads along the row, p moves to column jj + NU
                if (packed) {
                    int64_t s = 0;
ST_UNROLL
                    for (int u = 0; u < NU; u++) { dst[u] = p[s]; s += up ? (int64_t)jj + u + 1 : (int64_t)n - jj - u - 1; }
                    p += s;
                } else {
ST_UNROLL
                    for (int u = 0; u < NU; u++) dst[u] = p[u * cs];
                    p += NU * cs;
                }
            };
            auto compute = [&](T (&x)[NU], int jj) {
ST_UNROLL
                for (int u = 0; u < NU; u++) w[u] = vcol[jj + u];
ST_UNROLL
                for (int u = 0; u < NU; u++) {
                    acc[u] = el<T>::mad(cn ? el<T>::conj(x[u]) : x[u], w[u], acc[u]);
                    t[u] = el<T>::mul(ct ? el<T>::conj(x[u]) : x[u], vi);
                }
                sink.step(jj, t);
            };
            auto is_interior = [&](int jj) { return jj < jend && jj >= jin0 && jj + NU <= jin1 && (jj + NU <= i_first || jj > i_last); };
            fetch(a, j);
            for (;;) {                                     // the two register sets take turns: no copies between steps
                bool more = is_interior(j + NU);
                if (more) fetch(an, j + NU);
                compute(a, j);
                j += NU;
                if (!more) break;
                more = is_interior(j + NU);
                if (more) fetch(a, j + NU);
                compute(an, j);
                j += NU;
                if (!more) break;
            }
            continue;
        }
        int64_t s = 0;
ST_UNROLL
        for (int u = 0; u < NU; u++) {
            const bool in = j + u >= j0 && j + u < j1;
            a[u] = in ? p[s] : el<T>::zero();
            w[u] = in ? vcol[j + u] : el<T>::zero();
            s += col_step(D, j + u);
        }
        p += s;
ST_UNROLL
        for (int u = 0; u < NU; u++) {
            acc[u] = mad_elem<T>(a[u], w[u], i, j + u, nflags, acc[u]);
            t[u] = mad_elem<T>(a[u], vi, i, j + u, tflags, el<T>::zero());
        }
        sink.step(j, t);
        j += NU;
    }
    T r = acc[0];
ST_UNROLL
    for (int u = 1; u < NU; u++) r = el<T>::add(r, acc[u]);
    return r;
}
// sum over the row blocks [rb0, rb1] of their mirrored partial for column j: four independent chains in a fixed order (one chain of
// dependent loads per thread left the finish pass latency-bound)
template <typename T> ST_HD T sym_tp2_sum(const Desc& D, int j, const T* tp2, int64_t npadw, int rb0, int rb1) {
    T q0 = el<T>::zero(), q1 = q0, q2 = q0, q3 = q0;
    auto at = [&](int rb) { return tp2[(int64_t)rb * npadw + (j - sym_jw0(D, rb * ROW_THREADS))]; };
    int rb = rb0;
    for (; rb + 3 <= rb1; rb += 4) {
        const T x0 = at(rb), x1 = at(rb + 1), x2 = at(rb + 2), x3 = at(rb + 3);
        q0 = el<T>::add(q0, x0); q1 = el<T>::add(q1, x1); q2 = el<T>::add(q2, x2); q3 = el<T>::add(q3, x3);
    }
    for (; rb <= rb1; rb++) q0 = el<T>::add(q0, at(rb));
    return el<T>::add(el<T>::add(q0, q1), el<T>::add(q2, q3));
}
// the row blocks that hold stored rows of column j
ST_HD void sym_col_blocks(const Desc& D, int j, int& rb_lo, int& rb_hi) {
    int i0, i1;
    col_rows(D, j, i0, i1);
    rb_lo = i0 / ROW_THREADS; rb_hi = (i1 - 1) / ROW_THREADS;
}
// out(j) = alpha*(N partials + mirrored) + beta*old
template <typename T>
ST_HD T sym_finish_value(int j, int nparts, const T* part, int64_t npad, T mirrored, T alpha, T beta, bool beta0, T old) {
    T s = el<T>::zero();
    for (int c = 0; c < nparts; c++) s = el<T>::add(s, part[(int64_t)c * npad + j]);
    s = el<T>::mul(alpha, el<T>::add(s, mirrored));
    return beta0 ? s : el<T>::mad(beta, old, s);
}
template <typename T>
ST_HD T sym_finish_elem(const Desc& D, int j, int nparts, const T* part, int64_t npad, const T* tp2, int64_t npadw, T alpha, T beta, bool beta0, T old) {
    int rb_lo, rb_hi;
    sym_col_blocks(D, j, rb_lo, rb_hi);
    return sym_finish_value<T>(j, nparts, part, npad, sym_tp2_sum<T>(D, j, tp2, npadw, rb_lo, rb_hi), alpha, beta, beta0, old);
}

// out = alpha*(sum of the partial rows + tpart + vunit) + beta*old   (beta == 0: old is never read, like netlib)
template <typename T>
ST_HD T finish_elem(int i, int nparts, const T* part, int64_t npad, const T* tpart, const T* vunit, T alpha, T beta, bool beta0, T old) {
    T s = el<T>::zero();
    for (int c = 0; c < nparts; c++) s = el<T>::add(s, part[(int64_t)c * npad + i]);
    if (tpart) s = el<T>::add(s, tpart[i]);
    if (vunit) s = el<T>::add(s, vunit[i]);
    s = el<T>::mul(alpha, s);
    return beta0 ? s : el<T>::mad(beta, old, s);
}

// ---- rank-1 / rank-2 updates of the stored part ----
enum RankMode {
    R_GERU = 0,   // S(i,j) += alpha x(i) y(j)
    R_GERC = 1,   // S(i,j) += alpha x(i) conj(y(j))
    R_SYR2 = 2,   // S(i,j) += alpha x(i) y(j) + alpha y(i) x(j)
    R_HER = 3,    // S(i,j) += alpha x(i) conj(x(j)), alpha real (passed with zero imaginary part); diagonal kept real
    R_HER2 = 4,   // S(i,j) += alpha x(i) conj(y(j)) + conj(alpha) y(i) conj(x(j)); diagonal kept real
    R_SYR = 5     // S(i,j) += alpha x(i) x(j)
};
template <typename T> ST_HD T rank_elem(T a, int i, int j, T axi, T ayi, const T* x, const T* y, int mode) {
    switch (mode) {
        case R_GERU: a = el<T>::mad(axi, y[j], a); break;
        case R_GERC: a = el<T>::mad(axi, el<T>::conj(y[j]), a); break;
        case R_SYR: a = el<T>::mad(axi, x[j], a); break;
        case R_SYR2: a = el<T>::mad(axi, y[j], a); a = el<T>::mad(ayi, x[j], a); break;
        case R_HER: a = el<T>::mad(axi, el<T>::conj(x[j]), a); break;
        default: a = el<T>::mad(axi, el<T>::conj(y[j]), a); a = el<T>::mad(ayi, el<T>::conj(x[j]), a); break;
    }
    if ((mode == R_HER || mode == R_HER2) && i == j) a = el<T>::realpart(a);
    return a;
}
template <typename T> ST_HD void rank_row(const Desc& D, T* A, int i, int c0, int c1, T alpha, const T* x, const T* y, int mode, int i_first) {
    enum { NU = unroll_of<T>::N };
    int j0, j1, jf0, jf1;
    row_cols(D, i, j0, j1);
    j0 = st_max(j0, c0); j1 = st_min(j1, c1);
    if (j0 >= j1) return;
    row_cols(D, i_first, jf0, jf1);
    const int jstart = st_min(j0, st_max(jf0, c0));   // the warp walks the same columns in the same steps (see npart_row)
    const T xi = x[i];
    const T yi = (mode == R_SYR2 || mode == R_HER2) ? y[i] : el<T>::zero();
    const T axi = el<T>::mul(alpha, xi);                                          // alpha x(i)
    const T ayi = el<T>::mul(mode == R_HER2 ? el<T>::conj(alpha) : alpha, yi);     // alpha y(i)  /  conj(alpha) y(i)
    const bool packed = D.kind == K_PACKED, up = D.upper != 0;
    const int64_t cs = packed ? 0 : col_step(D, 0);
    T* p = A + off(D, i, jstart);
    for (int j = jstart; j < j1; j += NU) {   // NU read-modify-writes per step, loads first
        T a[NU];
        int64_t st[NU], s = 0;
        if (j >= j0 && j + NU <= j1) {        // interior step: no masks, a constant stride where the scheme has one
            if (packed) {
ST_UNROLL
                for (int u = 0; u < NU; u++) { st[u] = s; a[u] = p[s]; s += up ? (int64_t)j + u + 1 : (int64_t)D.n - j - u - 1; }
            } else {
ST_UNROLL
                for (int u = 0; u < NU; u++) { st[u] = u * cs; a[u] = p[u * cs]; }
                s = NU * cs;
            }
ST_UNROLL
            for (int u = 0; u < NU; u++) p[st[u]] = rank_elem<T>(a[u], i, j + u, axi, ayi, x, y, mode);
        } else {                              // head (columns left of this row's range) or tail
ST_UNROLL
            for (int u = 0; u < NU; u++) {
                st[u] = s;
                a[u] = (j + u >= j0 && j + u < j1) ? p[s] : el<T>::zero();
                s += col_step(D, j + u);
            }
ST_UNROLL
            for (int u = 0; u < NU; u++)
                if (j + u >= j0 && j + u < j1) p[st[u]] = rank_elem<T>(a[u], i, j + u, axi, ayi, x, y, mode);
        }
        p += s;
    }
}

// ---- triangular solves: entry (r,c) of op(S), zero outside the stored part ----
// trans: op(S)(r,c) = S(c,r); conj applies on top.  unit: the diagonal is taken as one and never read.
template <typename T> ST_HD bool solve_coef(const Desc& D, const T* A, int r, int c, bool trans, bool conj, T& out) {
    const int i = trans ? c : r, j = trans ? r : c;
    if (!stored(D, i, j)) return false;
    out = load_elem<T>(D, A, i, j, conj ? F_CONJ : 0);
    return true;
}

// element i of a BLAS vector with increment inc (negative increments run backwards from the end)
ST_HD int64_t vpos(int64_t i, int64_t n, int64_t inc) { return inc >= 0 ? i * inc : (n - 1 - i) * (-inc); }

// =====================================================================================================================
// Plans: what each routine asks of the three bodies, written once over a backend BE.  level2_struct.cu supplies the
// backend that launches the kernels on the call's stream; tests/drivers/struct_emul.cpp supplies one that walks the same
// grids on the CPU.  All pointers are device-accessible (host arrays in the emulation), arguments already validated.
//   BE::alloc(bytes)                                   scratch valid until the call returns
//   BE::sm_target()                                    CTAs wanted per launch (SM count x 8)
//   BE::gather(n, src, inc, dst, conj) / scatter(n, src, dst, inc)
//   BE::npart(D, A, v, row0, row1, c_lo, c_hi, cpc, nchunks, flags, part, npad)
//   BE::tpart(D, A, v, col0, col1, r0, r1, flags, tpart)
//   BE::finish(n, nparts, part, npad, tpart, vunit, alpha, beta, out, inco)
//   BE::sym_max_cols<T>()                              widest column strip a one-pass symmetric CTA can hold (0: two-pass form)
//   BE::sympart(D, A, v, cpc, nchunks, nflags, tflags, part, npad, tp2, npadw) / sym_finish(D, nparts, part, npad, tp2, npadw, alpha, beta, out, inco)
//   BE::rank(D, A, rows, ncols, cpc, nchunks, alpha, x, y, mode)
//   BE::solve_panel(D, A, x, p0, p1, trans, conj, unit, forward)     one CTA: the whole panel [p0,p1) in place
//   BE::solve_nupdate(D, A, x, row0, row1, b0, b1, flags) / solve_tupdate(D, A, x, col0, col1, b0, b1, flags)
// =====================================================================================================================

// operation after the layout mapping: 'N', 'T', 'C' (conjugate transpose), 'R' (conjugate, no transpose).  A row-major
// array is the column-major storage of the transpose, so N <-> T and ConjTrans becomes conjugate-no-transpose.
inline char map_op(char t, bool rowmajor) {
    if (!rowmajor) return t;
    return t == 'N' ? 'T' : (t == 'T' ? 'N' : 'R');
}
inline bool op_is_n(char op) { return op == 'N' || op == 'R'; }
inline bool op_is_conj(char op) { return op == 'C' || op == 'R'; }
inline int64_t packed_len(int n) { return (int64_t)n * ((int64_t)n + 1) / 2; }

template <typename BE> inline int col_chunks(BE& be, const Desc& D, int rows, int ncols) {
    if (D.kind == K_BAND_GEN || D.kind == K_BAND_TRI) return 1;   // a row holds at most kl+ku+1 stored columns
    const int rb = (rows + ROW_THREADS - 1) / ROW_THREADS;
    int c = (be.sm_target() + rb - 1) / rb;
    const int maxc = (ncols + 63) / 64;
    if (c > maxc) c = maxc;
    return c < 1 ? 1 : c;
}
// contiguous (optionally conjugated) copy of a BLAS vector in the call's scratch; the vector itself if it already is one
template <typename T, typename BE> inline const T* gathered(BE& be, int n, const T* x, int64_t inc, bool conj, bool force_copy = false) {
    if (inc == 1 && !conj && !force_copy) return x;
    T* t = (T*)be.alloc((size_t)(n > 0 ? n : 1) * sizeof(T));
    if (n > 0) be.gather(n, x, inc, t, conj);
    return t;
}
// out := alpha*(N part(nflags) + T part(tflags) + vunit) + beta*out; a part is off when its flags are < 0.
// N part: rows of S index the result and v is indexed by columns; T part: the other way round.
template <typename T, typename BE>
inline void smv(BE& be, const Desc& D, const T* A, const T* v, int nflags, int tflags, const T* vunit, T alpha, T beta, int nout, T* out, int64_t inco) {
    const int64_t npad = ((int64_t)nout + 31) / 32 * 32;
    int nch = 0;
    T *part = nullptr, *tpart = nullptr;
    if (nflags >= 0) {
        const int ncols = D.n, ch = col_chunks(be, D, nout, ncols), cpc = (ncols + ch - 1) / ch;
        nch = (ncols + cpc - 1) / cpc;
        part = (T*)be.alloc((size_t)nch * npad * sizeof(T));
        be.npart(D, A, v, 0, nout, 0, ncols, cpc, nch, nflags, part, npad);
    }
    if (tflags >= 0) {
        tpart = (T*)be.alloc((size_t)npad * sizeof(T));
        const int nrows = (D.kind == K_FULL_GEN || D.kind == K_BAND_GEN) ? D.m : D.n;
        be.tpart(D, A, v, 0, nout, 0, nrows, tflags, tpart);
    }
    be.finish(nout, nch, part, npad, tpart, vunit, alpha, beta, out, inco);
}
// solve op(S) x = b in place on a contiguous x.  A panel of columns is solved by ONE CTA (32-wide diagonal blocks by a warp,
// the rest of the panel updated by the CTA between them, see panel_block); the rows the panel reaches outside itself are
// then updated by a grid-wide pass.  A narrow band never reaches outside a panel as wide as the matrix, so TBSV with a
// small k is a single launch; packed / full triangles take n/256 panel launches + updates.
// (first version: one launch per 32-block + one per update -- 31 ms for DTPSV n=32768, 286 ms for DTBSV n=2^18 k=127,
//  all launch latency; profiles/r01e_level2_struct_perf_v1.txt)
enum { SOLVE_NB = 32, SOLVE_PANEL = 256, SOLVE_BAND_REACH = 1024, SOLVE_THREADS = 256 };
// block bi (in dependency order) of panel [p0,p1): its columns [b0,b1) and the panel rows [u0,u1) it must update
ST_HD void panel_block(const Desc& D, int p0, int p1, bool forward, int bi, int& b0, int& b1, int& u0, int& u1) {
    const int nblk = (p1 - p0 + SOLVE_NB - 1) / SOLVE_NB, b = forward ? bi : nblk - 1 - bi, rch = reach(D);
    b0 = p0 + b * SOLVE_NB; b1 = st_min(p1, b0 + SOLVE_NB);
    const int64_t reach_end = (int64_t)b1 + rch;   // op(S)(r,c) != 0 needs |r-c| <= reach
    u0 = forward ? b1 : st_max(p0, b0 - rch);
    u1 = forward ? (int)(reach_end < p1 ? reach_end : p1) : b0;
}
// the in-panel update of one unknown r by the solved block [b0,b1)
template <typename T> ST_HD T panel_update(const Desc& D, const T* A, const T* x, int r, int b0, int b1, bool trans, int flags) {
    return trans ? tpart_lane<T>(D, A, x, r, 0, 1, b0, b1, flags) : npart_row<T>(D, A, x, r, b0, b1, flags, r);
}
template <typename T, typename BE> inline void solve(BE& be, const Desc& D, const T* A, T* x, bool trans, bool conj, bool unit) {
    const int n = D.n, rch = reach(D);
    const bool forward = (D.upper != 0) == trans;   // op(S) is lower triangular
    const int flags = conj ? F_CONJ : 0;
    const int pw = rch <= SOLVE_BAND_REACH ? n : SOLVE_PANEL;
    const int npan = (n + pw - 1) / pw;
    for (int pi = 0; pi < npan; pi++) {
        const int pb = forward ? pi : npan - 1 - pi;
        const int p0 = pb * pw, p1 = (int64_t)p0 + pw < n ? p0 + pw : n;
        be.solve_panel(D, A, x, p0, p1, trans, conj, unit, forward);
        const int64_t reach_end = (int64_t)p1 + rch;
        const int u0 = forward ? p1 : st_max(0, p0 - rch), u1 = forward ? (int)(reach_end < n ? reach_end : n) : p0;
        if (u1 <= u0) continue;
        if (!trans) be.solve_nupdate(D, A, x, u0, u1, p0, p1, flags);
        else be.solve_tupdate(D, A, x, u0, u1, p0, p1, flags);
    }
}

// GBMV: trans in {N,T,C} as the caller gave it; (m, n, kl, ku) in the caller's layout
template <typename T, typename BE>
inline void plan_gbmv(BE& be, bool rowmajor, char trans, int m, int n, int kl, int ku, T alpha, const T* a, int64_t lda, const T* x, int64_t incx, T beta,
                      T* y, int64_t incy) {
    const int lenx = trans == 'N' ? n : m, leny = trans == 'N' ? m : n;
    const char op = map_op(trans, rowmajor);
    if (rowmajor) { int t = m; m = n; n = t; t = kl; kl = ku; ku = t; }
    Desc D = {K_BAND_GEN, m, n, kl, ku, 0, lda};
    if (el<T>::is_zero(alpha)) { smv<T>(be, D, (const T*)nullptr, (const T*)nullptr, -1, -1, (const T*)nullptr, alpha, beta, leny, y, incy); return; }
    const T* xc = gathered<T>(be, lenx, x, incx, false);
    const int f = op_is_conj(op) ? F_CONJ : 0;
    smv<T>(be, D, a, xc, op_is_n(op) ? f : -1, op_is_n(op) ? -1 : f, (const T*)nullptr, alpha, beta, leny, y, incy);
}
// SBMV / HBMV / SPMV / HPMV / HEMV: the stored triangle as is + its strict part mirrored (with a conjugate when Hermitian);
// a row-major Hermitian array holds conj(A) in the column-major view, which toggles the conjugations
template <typename T, typename BE>
inline void plan_symv_like(BE& be, int kind, bool herm, bool rowmajor, bool upper, int n, int k, T alpha, const T* a, int64_t lda, const T* x, int64_t incx,
                           T beta, T* y, int64_t incy) {
    if (rowmajor) upper = !upper;
    Desc D = {kind, n, n, k, k, upper ? 1 : 0, lda};
    if (el<T>::is_zero(alpha)) { smv<T>(be, D, (const T*)nullptr, (const T*)nullptr, -1, -1, (const T*)nullptr, alpha, beta, n, y, incy); return; }
    const T* xc = gathered<T>(be, n, x, incx, false);
    const int nf = herm ? (F_HERM | (rowmajor ? F_CONJ : 0)) : 0;
    const int tf = F_NODIAG | ((herm && !rowmajor) ? F_CONJ : 0);
    // one pass over the triangle when a CTA's strip of mirrored sums fits in shared memory (bands: ROW_THREADS + k columns)
    const int wmax = be.template sym_max_cols<T>();
    if (wmax > 0 && (kind != K_BAND_TRI || (int64_t)ROW_THREADS + k <= wmax)) {
        int cpc = n, nch = 1;
        if (kind != K_BAND_TRI) {
            const int ch = st_max(col_chunks(be, D, n, n), (n + wmax - 1) / wmax);
            cpc = ((n + ch - 1) / ch + 7) / 8 * 8;
            if (cpc > wmax) cpc = wmax;
            nch = (n + cpc - 1) / cpc;
        }
        const int64_t npad = ((int64_t)n + 31) / 32 * 32, npadw = (sym_width(D) + 31) / 32 * 32;
        const int nrb = (n + ROW_THREADS - 1) / ROW_THREADS;
        T* part = (T*)be.alloc((size_t)nch * npad * sizeof(T));
        T* tp2 = (T*)be.alloc((size_t)nrb * npadw * sizeof(T));
        be.sympart(D, a, xc, cpc, nch, nf, tf, part, npad, tp2, npadw);
        be.sym_finish(D, nch, part, npad, tp2, npadw, alpha, beta, y, incy);
        return;
    }
    smv<T>(be, D, a, xc, nf, tf, (const T*)nullptr, alpha, beta, n, y, incy);
}
// TBMV / TPMV / TRMV (solve = false) and TBSV / TPSV (solve = true)
template <typename T, typename BE>
inline void plan_tri(BE& be, int kind, bool do_solve, bool rowmajor, bool upper, char trans, bool unit, int n, int k, const T* a, int64_t lda, T* x,
                     int64_t incx) {
    const char op = map_op(trans, rowmajor);
    if (rowmajor) upper = !upper;
    Desc D = {kind, n, n, k, k, upper ? 1 : 0, lda};
    if (do_solve) {
        T* xc = (T*)gathered<T>(be, n, x, incx, false);
        solve<T>(be, D, a, xc, !op_is_n(op), op_is_conj(op), unit);
        if (xc != x) be.scatter(n, xc, x, incx);
    } else {
        const T* xc = gathered<T>(be, n, x, incx, false, true);   // the product reads all of x: out of place
        const int f = (op_is_conj(op) ? F_CONJ : 0) | (unit ? F_NODIAG : 0);
        smv<T>(be, D, a, xc, op_is_n(op) ? f : -1, op_is_n(op) ? -1 : f, unit ? xc : (const T*)nullptr, el<T>::one(), el<T>::zero(), n, x, incx);
    }
}
// GERU / GERC; row-major: the array is the column-major n x m matrix A^T, A^T += alpha * [conj]y * x^T
template <typename T, typename BE>
inline void plan_ger(BE& be, bool conjy, bool rowmajor, int m, int n, T alpha, const T* x, int64_t incx, const T* y, int64_t incy, T* a, int64_t lda) {
    if (!rowmajor) {
        Desc D = {K_FULL_GEN, m, n, 0, 0, 0, lda};
        const T* xc = gathered<T>(be, m, x, incx, false);
        const T* yc = gathered<T>(be, n, y, incy, false);
        const int ch = col_chunks(be, D, m, n), cpc = (n + ch - 1) / ch;
        be.rank(D, a, m, n, cpc, (n + cpc - 1) / cpc, alpha, xc, yc, conjy ? R_GERC : R_GERU);
    } else {
        Desc D = {K_FULL_GEN, n, m, 0, 0, 0, lda};
        const T* xc = gathered<T>(be, m, x, incx, false);
        const T* yc = gathered<T>(be, n, y, incy, conjy);
        const int ch = col_chunks(be, D, n, m), cpc = (m + ch - 1) / ch;
        be.rank(D, a, n, m, cpc, (m + cpc - 1) / cpc, alpha, yc, xc, R_GERU);
    }
}
// SYR2 / SPR / SPR2 / HER / HER2 / HPR / HPR2.  Row-major Hermitian: the array holds conj(A);
// conj(A) += alpha conj(x) conj(x)^H, and for the rank-2 form the same update with (x, y) := (conj y, conj x)
template <typename T, typename BE>
inline void plan_rank_sym(BE& be, int kind, int mode, bool rowmajor, bool upper, int n, T alpha, const T* x, int64_t incx, const T* y, int64_t incy, T* a,
                          int64_t lda) {
    const bool two = mode == R_SYR2 || mode == R_HER2, herm = mode == R_HER || mode == R_HER2;
    if (rowmajor) upper = !upper;
    Desc D = {kind, n, n, 0, 0, upper ? 1 : 0, lda};
    const bool cj = herm && rowmajor;
    const T* xc = gathered<T>(be, n, x, incx, cj);
    const T* yc = two ? gathered<T>(be, n, y, incy, cj) : (const T*)nullptr;
    if (cj && two) { const T* t = xc; xc = yc; yc = t; }
    const int ch = col_chunks(be, D, n, n), cpc = (n + ch - 1) / ch;
    be.rank(D, a, n, n, cpc, (n + cpc - 1) / cpc, alpha, xc, yc, mode);
}
// y = alpha*conj(S)*x + beta*y on a full m x n S (CBLAS row-major GEMV with ConjTrans, S the column-major view)
template <typename T, typename BE>
inline void plan_gemv_conj(BE& be, int m, int n, T alpha, const T* a, int64_t lda, const T* x, int64_t incx, T beta, T* y, int64_t incy) {
    Desc D = {K_FULL_GEN, m, n, 0, 0, 0, lda};
    const T* xc = gathered<T>(be, n, x, incx, false);
    smv<T>(be, D, a, xc, F_CONJ, -1, (const T*)nullptr, alpha, beta, m, y, incy);
}

}  // namespace st
}  // namespace b200
