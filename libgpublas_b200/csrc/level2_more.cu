// level2_more.cu -- more of the Level-1/2 surface the reference's dead wrappers name (SURVEY.md section 8(f) rank 3:
// blas_level2/ger.cc, symv.cc, trmv.cc, syr.cc, blas_level1/rot.cc, rotg.cc forward to cublas<t>ger / symv / trmv /
// syr / rot / rotg): real types s/d, Fortran + CBLAS entry points, netlib argument checks.  All are one pass over the
// matrix (HBM-bound), coalesced along the column-major contiguous index, deterministic.
//
//   GER / SYR   thread per row, CTA per (256 rows x column chunk): A(i,j) += alpha*x(i)*y(j)   [SYR: one triangle]
//   SYMV / TRMV built from two masked matrix-vector kernels over the stored triangle:
//               "N part"  r(i) = sum_j [keep(i,j)] A(i,j) v(j)   (thread per row, column chunks, partials summed in order)
//               "T part"  r(j) = sum_i [keep(i,j)] A(i,j) v(i)   (CTA per column, block tree reduction)
//               SYMV lower: y = beta*y + alpha*(N part over i>=j  +  T part over i>j); TRMV picks one part by trans.
//   ROT         element-wise; ROTG is scalar work and stays on the host.
#include "abi_common.h"
#include "gemm_generic.cuh"
#include "../../include/b200blas.h"
#include <cmath>
#include <cstdlib>

namespace b200 {

enum TriKeep { KEEP_LI = 0 /* i >= j */, KEEP_LS = 1 /* i > j */, KEEP_UI = 2 /* i <= j */, KEEP_US = 3 /* i < j */ };
__device__ __forceinline__ bool keep_ij(int mode, int i, int j) {
    return mode == KEEP_LI ? i >= j : (mode == KEEP_LS ? i > j : (mode == KEEP_UI ? i <= j : i < j));
}
__device__ __forceinline__ int64_t vpos(int64_t i, int64_t n, int64_t inc) { return inc >= 0 ? i * inc : (n - 1 - i) * (-inc); }

template <typename T>
__global__ void __launch_bounds__(256) ger_kernel(int m, int n, T alpha, const T* __restrict__ x, int64_t incx, const T* __restrict__ y,
                                                  int64_t incy, T* __restrict__ A, int64_t lda, int cols_per_chunk, int mask /* 0 full, 1 lower, 2 upper */) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const int c0 = blockIdx.y * cols_per_chunk, c1 = min(n, c0 + cols_per_chunk);
    const T ax = alpha * x[vpos(i, m, incx)];
    for (int j = c0; j < c1; j++) {
        if (mask == 1 && i < j) continue;
        if (mask == 2 && i > j) continue;
        T* p = A + i + (int64_t)j * lda;
        *p = num<T>::fma(ax, y[vpos(j, n, incy)], *p);
    }
}

// r_part[chunk][i] = sum over the chunk's columns j with keep(i,j) of A(i,j)*v(j)      (v contiguous)
template <typename T>
__global__ void __launch_bounds__(128) trimv_n_kernel(int n, const T* __restrict__ A, int64_t lda, const T* __restrict__ v, int mode,
                                                      int cols_per_chunk, T* __restrict__ part, int64_t npad) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= n) return;
    int c0 = blockIdx.y * cols_per_chunk, c1 = min(n, c0 + cols_per_chunk);
    if (mode == KEEP_LI) c1 = min(c1, i + 1); else if (mode == KEEP_LS) c1 = min(c1, i);
    else if (mode == KEEP_UI) c0 = max(c0, i); else c0 = max(c0, i + 1);
    T acc = 0;
    for (int j = c0; j < c1; j++) acc = num<T>::fma(A[i + (int64_t)j * lda], v[j], acc);
    part[(int64_t)blockIdx.y * npad + i] = acc;
}
// r(j) = sum_i [keep(i,j)] A(i,j) v(i)
template <typename T>
__global__ void __launch_bounds__(128) trimv_t_kernel(int n, const T* __restrict__ A, int64_t lda, const T* __restrict__ v, int mode, T* __restrict__ r) {
    __shared__ T sm[128];
    const int j = blockIdx.x;
    int i0 = 0, i1 = n;
    if (mode == KEEP_LI) i0 = j; else if (mode == KEEP_LS) i0 = j + 1; else if (mode == KEEP_UI) i1 = j + 1; else i1 = j;
    T acc = 0;
    for (int i = i0 + threadIdx.x; i < i1; i += 128) acc = num<T>::fma(A[i + (int64_t)j * lda], v[i], acc);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) r[j] = sm[0];
}
// out(i) = alpha*(sum_chunks part(i) + tpart(i) + (unit ? v(i) : 0)) + beta*out_old(i)
template <typename T>
__global__ void trimv_finish_kernel(int n, int chunks, const T* __restrict__ part, int64_t npad, const T* __restrict__ tpart, const T* __restrict__ vunit,
                                    T alpha, T beta, T* __restrict__ out, int64_t inco) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T s = 0;
    for (int c = 0; c < chunks; c++) s += part[(int64_t)c * npad + i];
    if (tpart) s += tpart[i];
    if (vunit) s += vunit[i];
    T* p = out + vpos(i, n, inco);
    *p = (beta == T(0)) ? alpha * s : num<T>::fma(beta, *p, alpha * s);
}
template <typename T> __global__ void rot_kernel(int64_t n, T* x, int64_t incx, T* y, int64_t incy, T c, T s) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        T* px = x + vpos(i, n, incx); T* py = y + vpos(i, n, incy);
        const T xv = *px, yv = *py;
        *px = c * xv + s * yv;
        *py = c * yv - s * xv;
    }
}

static int chunks_for(int row_blocks, int n) {
    const int target = (sm_count() > 0 ? sm_count() : 148) * 8;
    int c = (target + row_blocks - 1) / row_blocks;
    const int maxc = (n + 63) / 64;
    if (c > maxc) c = maxc;
    return c < 1 ? 1 : c;
}

template <typename T> static const T* contiguous(cudaStream_t s, int n, const T* x, int64_t incx) {
    if (incx == 1) return x;
    T* t = (T*)ws_alloc((size_t)n * sizeof(T));
    copy_dev<T>(s, n, x, incx, t, 1);
    return t;
}

template <typename T>
void ger_dev(cudaStream_t s, int m, int n, T alpha, const T* x, int64_t incx, const T* y, int64_t incy, T* A, int64_t lda, int mask) {
    const int rb = (m + 255) / 256, ch = chunks_for(rb, n), cpc = (n + ch - 1) / ch;
    ger_kernel<T><<<dim3(rb, (n + cpc - 1) / cpc), 256, 0, s>>>(m, n, alpha, x, incx, y, incy, A, lda, cpc, mask);
    last_variant = VAR_GENERIC_TILE;
}

// out := alpha * M v + beta * out, M described by (npart mode | -1, tpart mode | -1) over the stored triangle of A
template <typename T>
static void trimv(cudaStream_t s, int n, const T* A, int64_t lda, const T* v, int nmode, int tmode, bool unit, T alpha, T beta, T* out, int64_t inco) {
    const int rb = (n + 127) / 128, ch = nmode >= 0 ? chunks_for(rb, n) : 0, cpc = ch ? (n + ch - 1) / ch : 0;
    const int nch = ch ? (n + cpc - 1) / cpc : 0;
    const int64_t npad = ((int64_t)n + 31) / 32 * 32;
    T* part = nch ? (T*)ws_alloc((size_t)nch * npad * sizeof(T)) : nullptr;
    T* tpart = tmode >= 0 ? (T*)ws_alloc((size_t)npad * sizeof(T)) : nullptr;
    if (nch) trimv_n_kernel<T><<<dim3(rb, nch), 128, 0, s>>>(n, A, lda, v, nmode, cpc, part, npad);
    if (tpart) trimv_t_kernel<T><<<n, 128, 0, s>>>(n, A, lda, v, tmode, tpart);
    trimv_finish_kernel<T><<<(n + 255) / 256, 256, 0, s>>>(n, nch, part, npad, tpart, unit ? v : nullptr, alpha, beta, out, inco);
    last_variant = VAR_GENERIC_TILE;
}

template <typename T>
void symv_dev(cudaStream_t s, char uplo, int n, T alpha, const T* A, int64_t lda, const T* x, int64_t incx, T beta, T* y, int64_t incy) {
    const T* xc = contiguous<T>(s, n, x, incx);
    const bool upper = uplo == 'U';
    trimv<T>(s, n, A, lda, xc, upper ? KEEP_UI : KEEP_LI, upper ? KEEP_US : KEEP_LS, false, alpha, beta, y, incy);
}
template <typename T>
void trmv_dev(cudaStream_t s, char uplo, char trans, char diag, int n, const T* A, int64_t lda, T* x, int64_t incx) {
    // out of place: the product reads all of x
    T* xc = (T*)ws_alloc((size_t)n * sizeof(T));
    copy_dev<T>(s, n, x, incx, xc, 1);
    const bool upper = uplo == 'U', unit = diag == 'U', nota = trans == 'N';
    const int mode = upper ? (unit ? KEEP_US : KEEP_UI) : (unit ? KEEP_LS : KEEP_LI);
    trimv<T>(s, n, A, lda, xc, nota ? mode : -1, nota ? -1 : mode, unit, T(1), T(0), x, incx);
}
template <typename T> void rot_dev(cudaStream_t s, int64_t n, T* x, int64_t incx, T* y, int64_t incy, T c, T sn) {
    int64_t b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16;
    rot_kernel<T><<<(int)b, 256, 0, s>>>(n, x, incx, y, incy, c, sn);
    last_variant = VAR_GENERIC_TILE;
}

}  // namespace b200

using namespace b200;

namespace {

template <typename T>
void ger_entry(const char* name, const int* m, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* a, const int* lda) {
    int info = 0;                                       // netlib xGER: info 1,2,5,7,9
    if (*m < 0) info = 1; else if (*n < 0) info = 2; else if (*incx == 0) info = 5; else if (*incy == 0) info = 7; else if (*lda < imax(1, *m)) info = 9;
    if (info) { call_xerbla(name, info); return; }
    if (*m == 0 || *n == 0 || *alpha == T(0)) return;
    CallScope scope(name);
    Operand ox(x, 1 + (int64_t)(*m - 1) * abs(*incx), 1, 1 + (int64_t)(*m - 1) * abs(*incx), sizeof(T), ACC_IN);
    Operand oy(y, 1 + (int64_t)(*n - 1) * abs(*incy), 1, 1 + (int64_t)(*n - 1) * abs(*incy), sizeof(T), ACC_IN);
    Operand oa(a, *m, *n, *lda, sizeof(T), ACC_INOUT);
    ger_dev<T>(current_stream(), *m, *n, *alpha, (const T*)ox.dev(), *incx, (const T*)oy.dev(), *incy, (T*)oa.dev(), oa.ld(), 0);
    oa.release();
    log_exec(name, "m=%d n=%d lda=%d", *m, *n, *lda);
}
template <typename T>
void syr_entry(const char* name, const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, T* a, const int* lda) {
    const bool upper = lsame(uplo, 'U');
    int info = 0;                                       // netlib xSYR: info 1,2,5,7
    if (!upper && !lsame(uplo, 'L')) info = 1; else if (*n < 0) info = 2; else if (*incx == 0) info = 5; else if (*lda < imax(1, *n)) info = 7;
    if (info) { call_xerbla(name, info); return; }
    if (*n == 0 || *alpha == T(0)) return;
    CallScope scope(name);
    Operand ox(x, 1 + (int64_t)(*n - 1) * abs(*incx), 1, 1 + (int64_t)(*n - 1) * abs(*incx), sizeof(T), ACC_IN);
    Operand oa(a, *n, *n, *lda, sizeof(T), ACC_INOUT);
    ger_dev<T>(current_stream(), *n, *n, *alpha, (const T*)ox.dev(), *incx, (const T*)ox.dev(), *incx, (T*)oa.dev(), oa.ld(), upper ? 2 : 1);
    oa.release();
    log_exec(name, "%c n=%d lda=%d", upper ? 'U' : 'L', *n, *lda);
}
template <typename T>
void symv_entry(const char* name, const char* uplo, const int* n, const T* alpha, const T* a, const int* lda, const T* x, const int* incx, const T* beta,
                T* y, const int* incy) {
    const bool upper = lsame(uplo, 'U');
    int info = 0;                                       // netlib xSYMV: info 1,2,5,7,10
    if (!upper && !lsame(uplo, 'L')) info = 1; else if (*n < 0) info = 2; else if (*lda < imax(1, *n)) info = 5; else if (*incx == 0) info = 7;
    else if (*incy == 0) info = 10;
    if (info) { call_xerbla(name, info); return; }
    if (*n == 0 || (*alpha == T(0) && *beta == T(1))) return;
    CallScope scope(name);
    const bool a0 = *alpha == T(0);
    Operand oa(a0 ? nullptr : a, *n, *n, *lda, sizeof(T), ACC_IN);
    Operand ox(a0 ? nullptr : x, 1 + (int64_t)(*n - 1) * abs(*incx), 1, 1 + (int64_t)(*n - 1) * abs(*incx), sizeof(T), ACC_IN);
    Operand oy(y, 1 + (int64_t)(*n - 1) * abs(*incy), 1, 1 + (int64_t)(*n - 1) * abs(*incy), sizeof(T), ACC_INOUT);
    if (*alpha == T(0))   // netlib: A and x are not referenced
        trimv<T>(current_stream(), *n, nullptr, 1, nullptr, -1, -1, false, T(0), *beta, (T*)oy.dev(), *incy);
    else
        symv_dev<T>(current_stream(), upper ? 'U' : 'L', *n, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ox.dev(), *incx, *beta, (T*)oy.dev(), *incy);
    oy.release();
    log_exec(name, "%c n=%d lda=%d", upper ? 'U' : 'L', *n, *lda);
}
template <typename T>
void trmv_entry(const char* name, const char* uplo, const char* trans, const char* diag, const int* n, const T* a, const int* lda, T* x, const int* incx) {
    int info = 0;                                       // netlib xTRMV: info 1,2,3,4,6,8
    if (!lsame(uplo, 'U') && !lsame(uplo, 'L')) info = 1;
    else if (!lsame(trans, 'N') && !lsame(trans, 'T') && !lsame(trans, 'C')) info = 2;
    else if (!lsame(diag, 'U') && !lsame(diag, 'N')) info = 3;
    else if (*n < 0) info = 4; else if (*lda < imax(1, *n)) info = 6; else if (*incx == 0) info = 8;
    if (info) { call_xerbla(name, info); return; }
    if (*n == 0) return;
    CallScope scope(name);
    Operand oa(a, *n, *n, *lda, sizeof(T), ACC_IN);
    Operand ox(x, 1 + (int64_t)(*n - 1) * abs(*incx), 1, 1 + (int64_t)(*n - 1) * abs(*incx), sizeof(T), ACC_INOUT);
    trmv_dev<T>(current_stream(), lsame(uplo, 'U') ? 'U' : 'L', lsame(trans, 'N') ? 'N' : 'T', lsame(diag, 'U') ? 'U' : 'N', *n, (const T*)oa.dev(), oa.ld(),
                (T*)ox.dev(), *incx);
    ox.release();
    log_exec(name, "n=%d lda=%d", *n, *lda);
}
template <typename T>
void rot_entry(const char* name, const int* n, T* x, const int* incx, T* y, const int* incy, const T* c, const T* s) {
    if (*n <= 0) return;
    CallScope scope(name);
    Operand ox(x, 1 + (int64_t)(*n - 1) * abs(*incx), 1, 1 + (int64_t)(*n - 1) * abs(*incx), sizeof(T), ACC_INOUT);
    Operand oy(y, 1 + (int64_t)(*n - 1) * abs(*incy), 1, 1 + (int64_t)(*n - 1) * abs(*incy), sizeof(T), ACC_INOUT);
    rot_dev<T>(current_stream(), *n, (T*)ox.dev(), *incx, (T*)oy.dev(), *incy, *c, *s);
    ox.release(); oy.release();
    log_exec(name, "n=%d", *n);
}
// netlib xROTG (reference BLAS 3.8 formulation): scalar work, host only
template <typename T> void rotg_host(T* a, T* b, T* c, T* s) {
    const T aa = std::fabs(*a), ab = std::fabs(*b);
    const T roe = aa > ab ? *a : *b, scale = aa + ab;
    T r, z;
    if (scale == T(0)) { *c = 1; *s = 0; r = 0; z = 0; }
    else {
        r = scale * std::sqrt((*a / scale) * (*a / scale) + (*b / scale) * (*b / scale));
        r = std::copysign(T(1), roe) * r;
        *c = *a / r; *s = *b / r; z = 1;
        if (aa > ab) z = *s;
        if (ab >= aa && *c != T(0)) z = T(1) / *c;
    }
    *a = r; *b = z;
}

}  // namespace

extern "C" {
#define B200_L2MORE(P, T)                                                                                                                     \
    void P##ger_(const int* m, const int* n, const T* alpha, const T* x, const int* incx, const T* y, const int* incy, T* a, const int* lda) {  \
        ger_entry<T>(#P "ger_", m, n, alpha, x, incx, y, incy, a, lda); }                                                                        \
    void P##syr_(const char* uplo, const int* n, const T* alpha, const T* x, const int* incx, T* a, const int* lda) {                           \
        syr_entry<T>(#P "syr_", uplo, n, alpha, x, incx, a, lda); }                                                                              \
    void P##symv_(const char* uplo, const int* n, const T* alpha, const T* a, const int* lda, const T* x, const int* incx, const T* beta, T* y,  \
                  const int* incy) { symv_entry<T>(#P "symv_", uplo, n, alpha, a, lda, x, incx, beta, y, incy); }                                \
    void P##trmv_(const char* uplo, const char* trans, const char* diag, const int* n, const T* a, const int* lda, T* x, const int* incx) {      \
        trmv_entry<T>(#P "trmv_", uplo, trans, diag, n, a, lda, x, incx); }                                                                      \
    void P##rot_(const int* n, T* x, const int* incx, T* y, const int* incy, const T* c, const T* s) { rot_entry<T>(#P "rot_", n, x, incx, y, incy, c, s); } \
    void P##rotg_(T* a, T* b, T* c, T* s) { rotg_host<T>(a, b, c, s); }                                                                          \
    void cblas_##P##ger(enum CBLAS_ORDER order, int m, int n, T alpha, const T* x, int incx, const T* y, int incy, T* a, int lda) {               \
        if (order == CblasColMajor) P##ger_(&m, &n, &alpha, x, &incx, y, &incy, a, &lda);                                                        \
        else P##ger_(&n, &m, &alpha, y, &incy, x, &incx, a, &lda); }                                                                             \
    void cblas_##P##syr(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, T alpha, const T* x, int incx, T* a, int lda) {                      \
        char u = (uplo == CblasUpper) == (order == CblasColMajor) ? 'U' : 'L'; P##syr_(&u, &n, &alpha, x, &incx, a, &lda); }                     \
    void cblas_##P##symv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, int n, T alpha, const T* a, int lda, const T* x, int incx, T beta, T* y,  \
                         int incy) {                                                                                                             \
        char u = (uplo == CblasUpper) == (order == CblasColMajor) ? 'U' : 'L'; P##symv_(&u, &n, &alpha, a, &lda, x, &incx, &beta, y, &incy); }   \
    void cblas_##P##trmv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const T* a,       \
                         int lda, T* x, int incx) {                                                                                              \
        char u = (uplo == CblasUpper) == (order == CblasColMajor) ? 'U' : 'L';                                                                   \
        char t = (trans == CblasNoTrans) == (order == CblasColMajor) ? 'N' : 'T', d = diag == CblasUnit ? 'U' : 'N';                             \
        P##trmv_(&u, &t, &d, &n, a, &lda, x, &incx); }                                                                                           \
    void cblas_##P##rot(int n, T* x, int incx, T* y, int incy, T c, T s) { P##rot_(&n, x, &incx, y, &incy, &c, &s); }                             \
    void cblas_##P##rotg(T* a, T* b, T* c, T* s) { rotg_host<T>(a, b, c, s); }
B200_L2MORE(s, float)
B200_L2MORE(d, double)
}
