// fortran_l12.cu -- Fortran-ABI and CBLAS Level-1 / Level-2 entry points.
// The reference's wrappers for these (blas_level1/*.cc, blas_level2/gemv.cc, trsv.cc) are dead code
// that names the routines; the calling convention follows the CPU BLAS (gfortran ABI: functions
// return their value, complex results by value in registers) because the reference's blas.h macros
// for Level 1/2 are unreliable.  Defects of the dead wrappers that are NOT
// inherited: axpy dropping y (axpy.cc:28-30), cblas_i?amax returning a 1-based index
// (amax.cc:25,33-36), nrm2 exported under a misspelt name (nrm2.cc:31-54).
#include "abi_common.h"
#include <cstdlib>
#include <type_traits>
#include "../../include/b200blas.h"

using namespace b200;

namespace {

// Scalar results: the reduction kernel's finishing block stores the value straight into pinned, device-mapped host
// memory (zero-copy), so returning it costs one stream synchronisation and no separate copy (these routines are
// synchronous by nature: they return a value).
template <typename R> R* scalar_slot() { return (R*)armed_scalar(sizeof(R)); }
template <typename R> R fetch_scalar(const void* slot) {
    // spins on the slot instead of synchronising the stream (runtime.h); float / complex-float results arrive as 4-byte stores
    wait_scalar(slot, sizeof(R), (sizeof(R) == 4 || std::is_same<R, cuFloatComplex>::value) ? 4 : 8);
    R r;
    memcpy(&r, slot, sizeof(R));     // pinned host memory, written by the kernel's finishing block
    return r;
}

template <typename T> T dot_entry(const char* name, const int* n, const T* x, const int* incx, const T* y, const int* incy, bool conj) {
    T zero; memset(&zero, 0, sizeof zero);
    if (*n <= 0) return zero;
    CallScope scope(name);
    VecOperand ox(x, *n, *incx, sizeof(T), ACC_IN), oy(y, *n, *incy, sizeof(T), ACC_IN);
    T* out = scalar_slot<T>();
    dot_dev<T>(current_stream(), *n, (const T*)ox.dev(), *incx, (const T*)oy.dev(), *incy, out, conj);
    T r = fetch_scalar<T>(out);
    log_exec(name, "n=%d incx=%d incy=%d", *n, *incx, *incy);
    return r;
}
template <typename T, typename R> R nrm2_entry(const char* name, const int* n, const T* x, const int* incx, bool asum) {
    if (*n < 1 || *incx < 1) return R(0);
    CallScope scope(name);
    VecOperand ox(x, *n, *incx, sizeof(T), ACC_IN);
    R* out = scalar_slot<R>();
    if (asum) asum_dev<T, R>(current_stream(), *n, (const T*)ox.dev(), *incx, out);
    else nrm2_dev<T, R>(current_stream(), *n, (const T*)ox.dev(), *incx, out);
    R r = fetch_scalar<R>(out);
    log_exec(name, "n=%d incx=%d", *n, *incx);
    return r;
}
template <typename T> int iamax_entry(const char* name, const int* n, const T* x, const int* incx) {
    if (*n < 1 || *incx <= 0) return 0;
    CallScope scope(name);
    VecOperand ox(x, *n, *incx, sizeof(T), ACC_IN);
    long long* out = scalar_slot<long long>();
    iamax_dev<T>(current_stream(), *n, (const T*)ox.dev(), *incx, out);
    long long r = fetch_scalar<long long>(out);
    log_exec(name, "n=%d incx=%d", *n, *incx);
    return (int)(r + 1);   // Fortran: 1-based
}
template <typename T> void axpy_entry(const char* name, const int* n, const T* alpha, const T* x, const int* incx, T* y, const int* incy) {
    if (*n <= 0 || is0(*alpha)) return;
    CallScope scope(name);
    VecOperand ox(x, *n, *incx, sizeof(T), ACC_IN), oy(y, *n, *incy, sizeof(T), ACC_INOUT);
    axpy_dev<T>(current_stream(), *n, *alpha, (const T*)ox.dev(), *incx, (T*)oy.dev(), *incy);
    oy.release();
    log_exec(name, "n=%d incx=%d incy=%d", *n, *incx, *incy);
}
template <typename T, typename S> void scal_entry(const char* name, const int* n, const S* alpha, T* x, const int* incx) {
    if (*n <= 0 || *incx <= 0) return;
    CallScope scope(name);
    VecOperand ox(x, *n, *incx, sizeof(T), ACC_INOUT);
    scal_dev<T, S>(current_stream(), *n, *alpha, (T*)ox.dev(), *incx);
    ox.release();
    log_exec(name, "n=%d incx=%d", *n, *incx);
}
template <typename T> void copy_entry(const char* name, const int* n, const T* x, const int* incx, T* y, const int* incy, bool swap) {
    if (*n <= 0) return;
    CallScope scope(name);
    VecOperand ox(x, *n, *incx, sizeof(T), swap ? ACC_INOUT : ACC_IN);
    // strided copy targets keep their gaps: read-modify-write staging unless contiguous
    VecOperand oy(y, *n, *incy, sizeof(T), (swap || (*incy != 1 && *incy != -1)) ? ACC_INOUT : ACC_OUT);
    if (swap) swap_dev<T>(current_stream(), *n, (T*)ox.dev(), *incx, (T*)oy.dev(), *incy);
    else copy_dev<T>(current_stream(), *n, (const T*)ox.dev(), *incx, (T*)oy.dev(), *incy);
    if (swap) ox.release();
    oy.release();
    log_exec(name, "n=%d incx=%d incy=%d", *n, *incx, *incy);
}

// netlib xGEMV preamble: info 1,2,3,6,8,11
template <typename T>
void gemv_entry(const char* name, const char* trans, const int* m, const int* n, const T* alpha, const T* a, const int* lda,
                const T* x, const int* incx, const T* beta, T* y, const int* incy) {
    const bool nota = lsame(trans, 'N');
    int info = 0;
    if (!nota && !lsame(trans, 'T') && !lsame(trans, 'C')) info = 1;
    else if (*m < 0) info = 2;
    else if (*n < 0) info = 3;
    else if (*lda < imax(1, *m)) info = 6;
    else if (*incx == 0) info = 8;
    else if (*incy == 0) info = 11;
    if (info) { call_xerbla(name, info); return; }
    if (*m == 0 || *n == 0 || (is0(*alpha) && is1(*beta))) return;
    const int lenx = nota ? *n : *m, leny = nota ? *m : *n;
    const char t = nota ? 'N' : (lsame(trans, 'T') ? 'T' : 'C');
    CallScope scope(name);
    Operand oa(is0(*alpha) ? nullptr : a, *m, *n, *lda, sizeof(T), ACC_IN);
    VecOperand ox(is0(*alpha) ? nullptr : x, lenx, *incx, sizeof(T), ACC_IN);
    const bool strided_y = (*incy != 1 && *incy != -1);
    VecOperand oy(y, leny, *incy, sizeof(T), (is0(*beta) && !strided_y) ? ACC_OUT : ACC_INOUT);
    gemv_dev<T>(current_stream(), t, *m, *n, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ox.dev(), *incx, *beta, (T*)oy.dev(), *incy);
    oy.release();
    log_exec(name, "%c m=%d n=%d lda=%d incx=%d incy=%d", t, *m, *n, *lda, *incx, *incy);
}
// netlib xTRSV preamble: info 1,2,3,4,6,8
template <typename T>
void trsv_entry(const char* name, const char* uplo, const char* trans, const char* diag, const int* n, const T* a, const int* lda,
                T* x, const int* incx) {
    int info = 0;
    if (!lsame(uplo, 'U') && !lsame(uplo, 'L')) info = 1;
    else if (!lsame(trans, 'N') && !lsame(trans, 'T') && !lsame(trans, 'C')) info = 2;
    else if (!lsame(diag, 'U') && !lsame(diag, 'N')) info = 3;
    else if (*n < 0) info = 4;
    else if (*lda < imax(1, *n)) info = 6;
    else if (*incx == 0) info = 8;
    if (info) { call_xerbla(name, info); return; }
    if (*n == 0) return;
    CallScope scope(name);
    Operand oa(a, *n, *n, *lda, sizeof(T), ACC_IN);
    VecOperand ox(x, *n, *incx, sizeof(T), ACC_INOUT);
    const char u = lsame(uplo, 'U') ? 'U' : 'L', t = lsame(trans, 'N') ? 'N' : (lsame(trans, 'T') ? 'T' : 'C'), d = lsame(diag, 'U') ? 'U' : 'N';
    // netlib: with a negative increment, element i lives at x[(n-1-i)*|incx|]
    // The panel solver of level2_struct.cu (32-wide diagonal blocks staged ahead of the dependency chain, one CTA per 256-column
    // panel) against this file's diagonal kernel + GEMV update per 64 columns: 0.33 vs 0.73 ms at n = 2048, 1.96 vs 3.54 at 8192,
    // 8.1 vs 16.3 at 32768 (profiles/r02y7_trsv_panel_solver.txt).  B200BLAS_TRSV_STRUCT=0/1 forces one or the other.
    static const int st_env = getenv("B200BLAS_TRSV_STRUCT") ? atoi(getenv("B200BLAS_TRSV_STRUCT")) : -1;
    const bool panel = st_env >= 0 ? st_env != 0 : *n >= 512;
    if (panel) trsv_struct_dev<T>(current_stream(), u, t, d, *n, (const T*)oa.dev(), oa.ld(), (T*)ox.dev(), *incx);
    else trsv_dev<T>(current_stream(), u, t, d, *n, (const T*)oa.dev(), oa.ld(), (T*)ox.dev(), *incx);
    ox.release();
    log_exec(name, "%c%c%c n=%d lda=%d incx=%d", u, t, d, *n, *lda, *incx);
}

typedef cuFloatComplex c32;
typedef cuDoubleComplex c64;
}  // namespace

extern "C" {
// ---- DOT ----
float sdot_(const int* n, const float* x, const int* incx, const float* y, const int* incy) { return dot_entry<float>("sdot_", n, x, incx, y, incy, false); }
double ddot_(const int* n, const double* x, const int* incx, const double* y, const int* incy) { return dot_entry<double>("ddot_", n, x, incx, y, incy, false); }
b200_c32 cdotu_(const int* n, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy) { c32 r = dot_entry<c32>("cdotu_", n, (const c32*)x, incx, (const c32*)y, incy, false); return b200_c32{r.x, r.y}; }
b200_c32 cdotc_(const int* n, const b200_c32* x, const int* incx, const b200_c32* y, const int* incy) { c32 r = dot_entry<c32>("cdotc_", n, (const c32*)x, incx, (const c32*)y, incy, true); return b200_c32{r.x, r.y}; }
b200_c64 zdotu_(const int* n, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy) { c64 r = dot_entry<c64>("zdotu_", n, (const c64*)x, incx, (const c64*)y, incy, false); return b200_c64{r.x, r.y}; }
b200_c64 zdotc_(const int* n, const b200_c64* x, const int* incx, const b200_c64* y, const int* incy) { c64 r = dot_entry<c64>("zdotc_", n, (const c64*)x, incx, (const c64*)y, incy, true); return b200_c64{r.x, r.y}; }
// ---- NRM2 / ASUM ----
float snrm2_(const int* n, const float* x, const int* incx) { return nrm2_entry<float, float>("snrm2_", n, x, incx, false); }
double dnrm2_(const int* n, const double* x, const int* incx) { return nrm2_entry<double, double>("dnrm2_", n, x, incx, false); }
float scnrm2_(const int* n, const b200_c32* x, const int* incx) { return nrm2_entry<c32, float>("scnrm2_", n, (const c32*)x, incx, false); }
double dznrm2_(const int* n, const b200_c64* x, const int* incx) { return nrm2_entry<c64, double>("dznrm2_", n, (const c64*)x, incx, false); }
float sasum_(const int* n, const float* x, const int* incx) { return nrm2_entry<float, float>("sasum_", n, x, incx, true); }
double dasum_(const int* n, const double* x, const int* incx) { return nrm2_entry<double, double>("dasum_", n, x, incx, true); }
float scasum_(const int* n, const b200_c32* x, const int* incx) { return nrm2_entry<c32, float>("scasum_", n, (const c32*)x, incx, true); }
double dzasum_(const int* n, const b200_c64* x, const int* incx) { return nrm2_entry<c64, double>("dzasum_", n, (const c64*)x, incx, true); }
// ---- I?AMAX (1-based; 0 if n < 1 or incx <= 0) ----
int isamax_(const int* n, const float* x, const int* incx) { return iamax_entry<float>("isamax_", n, x, incx); }
int idamax_(const int* n, const double* x, const int* incx) { return iamax_entry<double>("idamax_", n, x, incx); }
int icamax_(const int* n, const b200_c32* x, const int* incx) { return iamax_entry<c32>("icamax_", n, (const c32*)x, incx); }
int izamax_(const int* n, const b200_c64* x, const int* incx) { return iamax_entry<c64>("izamax_", n, (const c64*)x, incx); }
// ---- AXPY / SCAL / COPY / SWAP ----
void saxpy_(const int* n, const float* alpha, const float* x, const int* incx, float* y, const int* incy) { axpy_entry<float>("saxpy_", n, alpha, x, incx, y, incy); }
void daxpy_(const int* n, const double* alpha, const double* x, const int* incx, double* y, const int* incy) { axpy_entry<double>("daxpy_", n, alpha, x, incx, y, incy); }
void caxpy_(const int* n, const b200_c32* alpha, const b200_c32* x, const int* incx, b200_c32* y, const int* incy) { axpy_entry<c32>("caxpy_", n, (const c32*)alpha, (const c32*)x, incx, (c32*)y, incy); }
void zaxpy_(const int* n, const b200_c64* alpha, const b200_c64* x, const int* incx, b200_c64* y, const int* incy) { axpy_entry<c64>("zaxpy_", n, (const c64*)alpha, (const c64*)x, incx, (c64*)y, incy); }
void sscal_(const int* n, const float* alpha, float* x, const int* incx) { scal_entry<float, float>("sscal_", n, alpha, x, incx); }
void dscal_(const int* n, const double* alpha, double* x, const int* incx) { scal_entry<double, double>("dscal_", n, alpha, x, incx); }
void cscal_(const int* n, const b200_c32* alpha, b200_c32* x, const int* incx) { scal_entry<c32, c32>("cscal_", n, (const c32*)alpha, (c32*)x, incx); }
void zscal_(const int* n, const b200_c64* alpha, b200_c64* x, const int* incx) { scal_entry<c64, c64>("zscal_", n, (const c64*)alpha, (c64*)x, incx); }
void csscal_(const int* n, const float* alpha, b200_c32* x, const int* incx) { scal_entry<c32, float>("csscal_", n, alpha, (c32*)x, incx); }
void zdscal_(const int* n, const double* alpha, b200_c64* x, const int* incx) { scal_entry<c64, double>("zdscal_", n, alpha, (c64*)x, incx); }
void scopy_(const int* n, const float* x, const int* incx, float* y, const int* incy) { copy_entry<float>("scopy_", n, x, incx, y, incy, false); }
void dcopy_(const int* n, const double* x, const int* incx, double* y, const int* incy) { copy_entry<double>("dcopy_", n, x, incx, y, incy, false); }
void ccopy_(const int* n, const b200_c32* x, const int* incx, b200_c32* y, const int* incy) { copy_entry<c32>("ccopy_", n, (const c32*)x, incx, (c32*)y, incy, false); }
void zcopy_(const int* n, const b200_c64* x, const int* incx, b200_c64* y, const int* incy) { copy_entry<c64>("zcopy_", n, (const c64*)x, incx, (c64*)y, incy, false); }
void sswap_(const int* n, float* x, const int* incx, float* y, const int* incy) { copy_entry<float>("sswap_", n, x, incx, y, incy, true); }
void dswap_(const int* n, double* x, const int* incx, double* y, const int* incy) { copy_entry<double>("dswap_", n, x, incx, y, incy, true); }
void cswap_(const int* n, b200_c32* x, const int* incx, b200_c32* y, const int* incy) { copy_entry<c32>("cswap_", n, (const c32*)x, incx, (c32*)y, incy, true); }
void zswap_(const int* n, b200_c64* x, const int* incx, b200_c64* y, const int* incy) { copy_entry<c64>("zswap_", n, (const c64*)x, incx, (c64*)y, incy, true); }
// ---- GEMV / TRSV ----
void sgemv_(const char* trans, const int* m, const int* n, const float* alpha, const float* a, const int* lda, const float* x, const int* incx, const float* beta, float* y, const int* incy) { gemv_entry<float>("sgemv_", trans, m, n, alpha, a, lda, x, incx, beta, y, incy); }
void dgemv_(const char* trans, const int* m, const int* n, const double* alpha, const double* a, const int* lda, const double* x, const int* incx, const double* beta, double* y, const int* incy) { gemv_entry<double>("dgemv_", trans, m, n, alpha, a, lda, x, incx, beta, y, incy); }
void cgemv_(const char* trans, const int* m, const int* n, const b200_c32* alpha, const b200_c32* a, const int* lda, const b200_c32* x, const int* incx, const b200_c32* beta, b200_c32* y, const int* incy) { gemv_entry<c32>("cgemv_", trans, m, n, (const c32*)alpha, (const c32*)a, lda, (const c32*)x, incx, (const c32*)beta, (c32*)y, incy); }
void zgemv_(const char* trans, const int* m, const int* n, const b200_c64* alpha, const b200_c64* a, const int* lda, const b200_c64* x, const int* incx, const b200_c64* beta, b200_c64* y, const int* incy) { gemv_entry<c64>("zgemv_", trans, m, n, (const c64*)alpha, (const c64*)a, lda, (const c64*)x, incx, (const c64*)beta, (c64*)y, incy); }
void strsv_(const char* uplo, const char* trans, const char* diag, const int* n, const float* a, const int* lda, float* x, const int* incx) { trsv_entry<float>("strsv_", uplo, trans, diag, n, a, lda, x, incx); }
void dtrsv_(const char* uplo, const char* trans, const char* diag, const int* n, const double* a, const int* lda, double* x, const int* incx) { trsv_entry<double>("dtrsv_", uplo, trans, diag, n, a, lda, x, incx); }
void ctrsv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c32* a, const int* lda, b200_c32* x, const int* incx) { trsv_entry<c32>("ctrsv_", uplo, trans, diag, n, (const c32*)a, lda, (c32*)x, incx); }
void ztrsv_(const char* uplo, const char* trans, const char* diag, const int* n, const b200_c64* a, const int* lda, b200_c64* x, const int* incx) { trsv_entry<c64>("ztrsv_", uplo, trans, diag, n, (const c64*)a, lda, (c64*)x, incx); }

// ---------------------------------- CBLAS Level 1 / 2 ----------------------------------
float cblas_sdot(int n, const float* x, int incx, const float* y, int incy) { return sdot_(&n, x, &incx, y, &incy); }
double cblas_ddot(int n, const double* x, int incx, const double* y, int incy) { return ddot_(&n, x, &incx, y, &incy); }
void cblas_cdotu_sub(int n, const void* x, int incx, const void* y, int incy, void* out) { *(b200_c32*)out = cdotu_(&n, (const b200_c32*)x, &incx, (const b200_c32*)y, &incy); }
void cblas_cdotc_sub(int n, const void* x, int incx, const void* y, int incy, void* out) { *(b200_c32*)out = cdotc_(&n, (const b200_c32*)x, &incx, (const b200_c32*)y, &incy); }
void cblas_zdotu_sub(int n, const void* x, int incx, const void* y, int incy, void* out) { *(b200_c64*)out = zdotu_(&n, (const b200_c64*)x, &incx, (const b200_c64*)y, &incy); }
void cblas_zdotc_sub(int n, const void* x, int incx, const void* y, int incy, void* out) { *(b200_c64*)out = zdotc_(&n, (const b200_c64*)x, &incx, (const b200_c64*)y, &incy); }
float cblas_snrm2(int n, const float* x, int incx) { return snrm2_(&n, x, &incx); }
double cblas_dnrm2(int n, const double* x, int incx) { return dnrm2_(&n, x, &incx); }
float cblas_scnrm2(int n, const void* x, int incx) { return scnrm2_(&n, (const b200_c32*)x, &incx); }
double cblas_dznrm2(int n, const void* x, int incx) { return dznrm2_(&n, (const b200_c64*)x, &incx); }
float cblas_sasum(int n, const float* x, int incx) { return sasum_(&n, x, &incx); }
double cblas_dasum(int n, const double* x, int incx) { return dasum_(&n, x, &incx); }
// CBLAS index is 0-based (the reference's wrapper gets this wrong, amax.cc:25,33-36)
CBLAS_INDEX cblas_isamax(int n, const float* x, int incx) { int r = isamax_(&n, x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
CBLAS_INDEX cblas_idamax(int n, const double* x, int incx) { int r = idamax_(&n, x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
CBLAS_INDEX cblas_icamax(int n, const void* x, int incx) { int r = icamax_(&n, (const b200_c32*)x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
CBLAS_INDEX cblas_izamax(int n, const void* x, int incx) { int r = izamax_(&n, (const b200_c64*)x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
void cblas_saxpy(int n, float alpha, const float* x, int incx, float* y, int incy) { saxpy_(&n, &alpha, x, &incx, y, &incy); }
void cblas_daxpy(int n, double alpha, const double* x, int incx, double* y, int incy) { daxpy_(&n, &alpha, x, &incx, y, &incy); }
void cblas_caxpy(int n, const void* alpha, const void* x, int incx, void* y, int incy) { caxpy_(&n, (const b200_c32*)alpha, (const b200_c32*)x, &incx, (b200_c32*)y, &incy); }
void cblas_zaxpy(int n, const void* alpha, const void* x, int incx, void* y, int incy) { zaxpy_(&n, (const b200_c64*)alpha, (const b200_c64*)x, &incx, (b200_c64*)y, &incy); }
void cblas_sscal(int n, float alpha, float* x, int incx) { sscal_(&n, &alpha, x, &incx); }
void cblas_dscal(int n, double alpha, double* x, int incx) { dscal_(&n, &alpha, x, &incx); }
void cblas_scopy(int n, const float* x, int incx, float* y, int incy) { scopy_(&n, x, &incx, y, &incy); }
void cblas_dcopy(int n, const double* x, int incx, double* y, int incy) { dcopy_(&n, x, &incx, y, &incy); }
void cblas_sswap(int n, float* x, int incx, float* y, int incy) { sswap_(&n, x, &incx, y, &incy); }
void cblas_dswap(int n, double* x, int incx, double* y, int incy) { dswap_(&n, x, &incx, y, &incy); }

static inline char tr2(enum CBLAS_TRANSPOSE t) { return t == CblasNoTrans ? 'N' : (t == CblasTrans ? 'T' : (t == CblasConjTrans ? 'C' : '?')); }
// row-major A(m x n) is the column-major n x m matrix A^T: swap m/n and flip the transpose
// (real types; the reference's wrappers instead call an undefined transpose() helper, gemv.cc:39-46)
void cblas_sgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, float alpha, const float* a, int lda, const float* x, int incx, float beta, float* y, int incy) {
    char t = tr2(trans);
    if (order == CblasColMajor) sgemv_(&t, &m, &n, &alpha, a, &lda, x, &incx, &beta, y, &incy);
    else { char t2 = (t == 'N') ? 'T' : (t == '?' ? '?' : 'N'); sgemv_(&t2, &n, &m, &alpha, a, &lda, x, &incx, &beta, y, &incy); }
}
void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans, int m, int n, double alpha, const double* a, int lda, const double* x, int incx, double beta, double* y, int incy) {
    char t = tr2(trans);
    if (order == CblasColMajor) dgemv_(&t, &m, &n, &alpha, a, &lda, x, &incx, &beta, y, &incy);
    else { char t2 = (t == 'N') ? 'T' : (t == '?' ? '?' : 'N'); dgemv_(&t2, &n, &m, &alpha, a, &lda, x, &incx, &beta, y, &incy); }
}
void cblas_strsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const float* a, int lda, float* x, int incx) {
    char u = uplo == CblasUpper ? 'U' : (uplo == CblasLower ? 'L' : '?'), t = tr2(trans), d = diag == CblasUnit ? 'U' : (diag == CblasNonUnit ? 'N' : '?');
    if (order == CblasRowMajor) { u = (u == 'U') ? 'L' : (u == 'L' ? 'U' : '?'); t = (t == 'N') ? 'T' : (t == '?' ? '?' : 'N'); }
    strsv_(&u, &t, &d, &n, a, &lda, x, &incx);
}
void cblas_dtrsv(enum CBLAS_ORDER order, enum CBLAS_UPLO uplo, enum CBLAS_TRANSPOSE trans, enum CBLAS_DIAG diag, int n, const double* a, int lda, double* x, int incx) {
    char u = uplo == CblasUpper ? 'U' : (uplo == CblasLower ? 'L' : '?'), t = tr2(trans), d = diag == CblasUnit ? 'U' : (diag == CblasNonUnit ? 'N' : '?');
    if (order == CblasRowMajor) { u = (u == 'U') ? 'L' : (u == 'L' ? 'U' : '?'); t = (t == 'N') ? 'T' : (t == '?' ? '?' : 'N'); }
    dtrsv_(&u, &t, &d, &n, a, &lda, x, &incx);
}
}  // extern "C"
