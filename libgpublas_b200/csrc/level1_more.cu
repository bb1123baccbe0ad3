// level1_more.cu -- the rest of the Level-1 surface the reference's dead wrappers name:
// blas_level1/rot.cc, rotg.cc (cublas<t>rot / rotg), rotm.cc:10-57 (cublas<t>rotm), rotmg.cc:11-28 (cublas<t>rotmg), amin.cc:10-56 (cublasI<t>amin), the
// DSDOT / SDSDOT prototypes of cblas.h, CSROT / ZDROT, and the cblas_ forms of the complex copy / swap / scal / asum that
// fortran_l12.cu exports as Fortran symbols only.  Element-wise kernels are one coalesced pass; ROTMG is scalar work and
// stays on the host, like ROTG.
#include "abi_common.h"
#include "../../include/b200blas.h"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace b200 {

__device__ __forceinline__ int64_t l1m_pos(int64_t i, int64_t n, int64_t inc) { return inc >= 0 ? i * inc : (n - 1 - i) * (-inc); }

// netlib xROTM: (x_i, y_i) := H (x_i, y_i), H given by its four entries (the flag has been decoded on the host)
template <typename T>
__global__ void rotm_kernel(int64_t n, T* x, int64_t incx, T* y, int64_t incy, T h11, T h21, T h12, T h22) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        T* px = x + l1m_pos(i, n, incx); T* py = y + l1m_pos(i, n, incy);
        const T w = *px, z = *py;
        *px = w * h11 + z * h12;
        *py = w * h21 + z * h22;
    }
}
// netlib CSROT / ZDROT: plane rotation of complex vectors with real cosine and sine
template <typename CT, typename R>
__global__ void crot_kernel(int64_t n, CT* x, int64_t incx, CT* y, int64_t incy, R c, R s) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        CT* px = x + l1m_pos(i, n, incx); CT* py = y + l1m_pos(i, n, incy);
        const CT xv = *px, yv = *py;
        CT nx, ny;
        nx.x = c * xv.x + s * yv.x; nx.y = c * xv.y + s * yv.y;
        ny.x = c * yv.x - s * xv.x; ny.y = c * yv.y - s * xv.y;
        *px = nx; *py = ny;
    }
}
// netlib xROT (reference blas_level1/rot.cc): real plane rotation
template <typename T> __global__ void rot_kernel(int64_t n, T* x, int64_t incx, T* y, int64_t incy, T c, T s) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        T* px = x + l1m_pos(i, n, incx); T* py = y + l1m_pos(i, n, incy);
        const T xv = *px, yv = *py;
        *px = c * xv + s * yv;
        *py = c * yv - s * xv;
    }
}
template <typename T> void rot_dev(cudaStream_t s, int64_t n, T* x, int64_t incx, T* y, int64_t incy, T c, T sn) {
    int64_t b = (n + 255) / 256; if (b > 148 * 16) b = 148 * 16;
    rot_kernel<T><<<(int)b, 256, 0, s>>>(n, x, incx, y, incy, c, sn);
    last_variant = VAR_GENERIC_TILE;
}

static int l1m_blocks(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int cap = (sm_count() > 0 ? sm_count() : 148) * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace b200

using namespace b200;

namespace {

typedef VecOperand Vec;   // runtime.h: strided vectors are written back element by element
template <typename R> R fetch_scalar(const void* slot) {   // the kernel's finishing block writes into pinned, device-mapped host memory
    wait_scalar(slot, sizeof(R), sizeof(R) == 4 ? 4 : 8);
    R r;
    memcpy(&r, slot, sizeof(R));
    return r;
}

// netlib xROTM: param[0] = flag; -2: H = I (nothing to do); -1: full H; 0: unit diagonal; 1: h12 = 1, h21 = -1
template <typename T> void rotm_entry(const char* name, const int* n, T* x, const int* incx, T* y, const int* incy, const T* param) {
    const T flag = param[0];
    if (*n <= 0 || flag == T(-2)) return;
    T h11, h21, h12, h22;
    if (flag < T(0)) { h11 = param[1]; h21 = param[2]; h12 = param[3]; h22 = param[4]; }
    else if (flag == T(0)) { h11 = 1; h21 = param[2]; h12 = param[3]; h22 = 1; }
    else { h11 = param[1]; h21 = -1; h12 = 1; h22 = param[4]; }
    CallScope scope(name);
    Vec ox(x, *n, *incx, sizeof(T), ACC_INOUT), oy(y, *n, *incy, sizeof(T), ACC_INOUT);
    rotm_kernel<T><<<l1m_blocks(*n), 256, 0, current_stream()>>>(*n, (T*)ox.dev(), *incx, (T*)oy.dev(), *incy, h11, h21, h12, h22);
    last_variant = VAR_GENERIC_TILE;
    ox.release(); oy.release();
    log_exec(name, "n=%d flag=%g", *n, (double)flag);
}

// netlib xROTMG (reference BLAS 3.8): builds the modified Givens transformation that zeroes the second component of
// (sqrt(d1) x1, sqrt(d2) y1); rescaling keeps d1, d2 within [gam^-2, gam^2].
template <typename T> void rotmg_host(T* d1, T* d2, T* x1, const T* y1, T* param) {
    const T gam = 4096, gamsq = gam * gam, rgamsq = T(1) / gamsq;
    T flag, h11 = 0, h12 = 0, h21 = 0, h22 = 0;
    if (*d1 < 0) { flag = -1; *d1 = 0; *d2 = 0; *x1 = 0; }
    else {
        const T p2 = *d2 * *y1;
        if (p2 == 0) { param[0] = -2; return; }
        const T p1 = *d1 * *x1, q2 = p2 * *y1, q1 = p1 * *x1;
        if (std::fabs(q1) > std::fabs(q2)) {
            h21 = -*y1 / *x1; h12 = p2 / p1;
            const T u = 1 - h12 * h21;
            if (u > 0) { flag = 0; *d1 /= u; *d2 /= u; *x1 *= u; }
            else { flag = -1; h11 = h12 = h21 = h22 = 0; *d1 = 0; *d2 = 0; *x1 = 0; }
        } else if (q2 < 0) { flag = -1; h11 = h12 = h21 = h22 = 0; *d1 = 0; *d2 = 0; *x1 = 0; }
        else {
            flag = 1; h11 = p1 / p2; h22 = *x1 / *y1;
            const T u = 1 + h11 * h22, t = *d2 / u;
            *d2 = *d1 / u; *d1 = t; *x1 = *y1 * u;
        }
        if (*d1 != 0)
            while (*d1 <= rgamsq || *d1 >= gamsq) {
                if (flag == 0) { h11 = 1; h22 = 1; flag = -1; } else if (flag > 0) { h21 = -1; h12 = 1; flag = -1; }
                if (*d1 <= rgamsq) { *d1 *= gamsq; *x1 /= gam; h11 /= gam; h12 /= gam; }
                else { *d1 /= gamsq; *x1 *= gam; h11 *= gam; h12 *= gam; }
            }
        if (*d2 != 0)
            while (std::fabs(*d2) <= rgamsq || std::fabs(*d2) >= gamsq) {
                if (flag == 0) { h11 = 1; h22 = 1; flag = -1; } else if (flag > 0) { h21 = -1; h12 = 1; flag = -1; }
                if (std::fabs(*d2) <= rgamsq) { *d2 *= gamsq; h21 /= gam; h22 /= gam; }
                else { *d2 /= gamsq; h21 *= gam; h22 *= gam; }
            }
    }
    if (flag < 0) { param[1] = h11; param[2] = h21; param[3] = h12; param[4] = h22; }
    else if (flag == 0) { param[2] = h21; param[3] = h12; }
    else { param[1] = h11; param[4] = h22; }
    param[0] = flag;
}

// netlib CROTG / ZROTG (reference BLAS 3.8 formulation): complex Givens rotation, scalar host work
template <typename CT, typename R> void crotg_host(CT* ca, const CT* cb, R* c, CT* s) {
    const R aa = std::hypot(ca->re, ca->im), ab = std::hypot(cb->re, cb->im);
    if (aa == R(0)) { *c = 0; s->re = 1; s->im = 0; *ca = *cb; return; }
    const R scale = aa + ab;
    const R norm = scale * std::sqrt((aa / scale) * (aa / scale) + (ab / scale) * (ab / scale));
    const R alr = ca->re / aa, ali = ca->im / aa;                 // alpha = ca / |ca|
    *c = aa / norm;
    // s = alpha * conj(cb) / norm
    s->re = (alr * cb->re + ali * cb->im) / norm;
    s->im = (ali * cb->re - alr * cb->im) / norm;
    ca->re = alr * norm; ca->im = ali * norm;
}

template <typename CT, typename R> void crot_entry(const char* name, const int* n, CT* x, const int* incx, CT* y, const int* incy, const R* c, const R* s) {
    if (*n <= 0) return;
    CallScope scope(name);
    Vec ox(x, *n, *incx, sizeof(CT), ACC_INOUT), oy(y, *n, *incy, sizeof(CT), ACC_INOUT);
    crot_kernel<CT, R><<<l1m_blocks(*n), 256, 0, current_stream()>>>(*n, (CT*)ox.dev(), *incx, (CT*)oy.dev(), *incy, *c, *s);
    last_variant = VAR_GENERIC_TILE;
    ox.release(); oy.release();
    log_exec(name, "n=%d", *n);
}

template <typename T> int iamin_entry(const char* name, const int* n, const T* x, const int* incx) {
    if (*n < 1 || *incx <= 0) return 0;
    CallScope scope(name);
    Vec ox(x, *n, *incx, sizeof(T), ACC_IN);
    long long* out = (long long*)armed_scalar(sizeof(long long));
    iamin_dev<T>(current_stream(), *n, (const T*)ox.dev(), *incx, out);
    const long long r = fetch_scalar<long long>(out);
    log_exec(name, "n=%d incx=%d", *n, *incx);
    return (int)(r + 1);   // Fortran: 1-based
}

double dsdot_entry(const char* name, const int* n, float sb, const float* x, const int* incx, const float* y, const int* incy, bool out_double) {
    if (*n <= 0) return out_double ? (double)sb : (double)(float)sb;
    CallScope scope(name);
    Vec ox(x, *n, *incx, sizeof(float), ACC_IN), oy(y, *n, *incy, sizeof(float), ACC_IN);
    void* out = armed_scalar(out_double ? sizeof(double) : sizeof(float));
    dsdot_dev(current_stream(), *n, (const float*)ox.dev(), *incx, (const float*)oy.dev(), *incy, (double)sb, out, out_double);
    const double r = out_double ? fetch_scalar<double>(out) : (double)fetch_scalar<float>(out);
    log_exec(name, "n=%d incx=%d incy=%d", *n, *incx, *incy);
    return r;
}

template <typename T>
void rot_entry(const char* name, const int* n, T* x, const int* incx, T* y, const int* incy, const T* c, const T* s) {
    if (*n <= 0) return;
    CallScope scope(name);
    Vec ox(x, *n, *incx, sizeof(T), ACC_INOUT), oy(y, *n, *incy, sizeof(T), ACC_INOUT);
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

typedef cuFloatComplex c32;
typedef cuDoubleComplex c64;

}  // namespace

extern "C" {
#define B200_ROT(P, T)                                                                                                                        \
    void P##rot_(const int* n, T* x, const int* incx, T* y, const int* incy, const T* c, const T* s) { rot_entry<T>(#P "rot_", n, x, incx, y, incy, c, s); } \
    void P##rotg_(T* a, T* b, T* c, T* s) { rotg_host<T>(a, b, c, s); }                                                                          \
    void cblas_##P##rot(int n, T* x, int incx, T* y, int incy, T c, T s) { P##rot_(&n, x, &incx, y, &incy, &c, &s); }                             \
    void cblas_##P##rotg(T* a, T* b, T* c, T* s) { rotg_host<T>(a, b, c, s); }
B200_ROT(s, float)
B200_ROT(d, double)
#undef B200_ROT
void srotm_(const int* n, float* x, const int* incx, float* y, const int* incy, const float* param) { rotm_entry<float>("srotm_", n, x, incx, y, incy, param); }
void drotm_(const int* n, double* x, const int* incx, double* y, const int* incy, const double* param) { rotm_entry<double>("drotm_", n, x, incx, y, incy, param); }
void srotmg_(float* d1, float* d2, float* x1, const float* y1, float* param) { rotmg_host<float>(d1, d2, x1, y1, param); }
void drotmg_(double* d1, double* d2, double* x1, const double* y1, double* param) { rotmg_host<double>(d1, d2, x1, y1, param); }
void csrot_(const int* n, b200_c32* x, const int* incx, b200_c32* y, const int* incy, const float* c, const float* s) { crot_entry<c32, float>("csrot_", n, (c32*)x, incx, (c32*)y, incy, c, s); }
void zdrot_(const int* n, b200_c64* x, const int* incx, b200_c64* y, const int* incy, const double* c, const double* s) { crot_entry<c64, double>("zdrot_", n, (c64*)x, incx, (c64*)y, incy, c, s); }
void crotg_(b200_c32* ca, const b200_c32* cb, float* c, b200_c32* s) { crotg_host<b200_c32, float>(ca, cb, c, s); }
void zrotg_(b200_c64* ca, const b200_c64* cb, double* c, b200_c64* s) { crotg_host<b200_c64, double>(ca, cb, c, s); }
void cblas_crotg(void* a, void* b, float* c, void* s) { crotg_host<b200_c32, float>((b200_c32*)a, (const b200_c32*)b, c, (b200_c32*)s); }
void cblas_zrotg(void* a, void* b, double* c, void* s) { crotg_host<b200_c64, double>((b200_c64*)a, (const b200_c64*)b, c, (b200_c64*)s); }
// I?AMIN (1-based; 0 if n < 1 or incx <= 0) -- not in netlib; the reference's amin.cc and the CPU BLAS (OpenBLAS) export it
int isamin_(const int* n, const float* x, const int* incx) { return iamin_entry<float>("isamin_", n, x, incx); }
int idamin_(const int* n, const double* x, const int* incx) { return iamin_entry<double>("idamin_", n, x, incx); }
int icamin_(const int* n, const b200_c32* x, const int* incx) { return iamin_entry<c32>("icamin_", n, (const c32*)x, incx); }
int izamin_(const int* n, const b200_c64* x, const int* incx) { return iamin_entry<c64>("izamin_", n, (const c64*)x, incx); }
double dsdot_(const int* n, const float* x, const int* incx, const float* y, const int* incy) { return dsdot_entry("dsdot_", n, 0.f, x, incx, y, incy, true); }
float sdsdot_(const int* n, const float* sb, const float* x, const int* incx, const float* y, const int* incy) { return (float)dsdot_entry("sdsdot_", n, *sb, x, incx, y, incy, false); }

void cblas_srotm(int n, float* x, int incx, float* y, int incy, const float* param) { srotm_(&n, x, &incx, y, &incy, param); }
void cblas_drotm(int n, double* x, int incx, double* y, int incy, const double* param) { drotm_(&n, x, &incx, y, &incy, param); }
void cblas_srotmg(float* d1, float* d2, float* x1, float y1, float* param) { rotmg_host<float>(d1, d2, x1, &y1, param); }
void cblas_drotmg(double* d1, double* d2, double* x1, double y1, double* param) { rotmg_host<double>(d1, d2, x1, &y1, param); }
void cblas_csrot(int n, void* x, int incx, void* y, int incy, float c, float s) { csrot_(&n, (b200_c32*)x, &incx, (b200_c32*)y, &incy, &c, &s); }
void cblas_zdrot(int n, void* x, int incx, void* y, int incy, double c, double s) { zdrot_(&n, (b200_c64*)x, &incx, (b200_c64*)y, &incy, &c, &s); }
// CBLAS index is 0-based
CBLAS_INDEX cblas_isamin(int n, const float* x, int incx) { int r = isamin_(&n, x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
CBLAS_INDEX cblas_idamin(int n, const double* x, int incx) { int r = idamin_(&n, x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
CBLAS_INDEX cblas_icamin(int n, const void* x, int incx) { int r = icamin_(&n, (const b200_c32*)x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
CBLAS_INDEX cblas_izamin(int n, const void* x, int incx) { int r = izamin_(&n, (const b200_c64*)x, &incx); return r > 0 ? (CBLAS_INDEX)(r - 1) : 0; }
double cblas_dsdot(int n, const float* x, int incx, const float* y, int incy) { return dsdot_(&n, x, &incx, y, &incy); }
float cblas_sdsdot(int n, float sb, const float* x, int incx, const float* y, int incy) { return sdsdot_(&n, &sb, x, &incx, y, &incy); }
// cblas_ forms of the complex Level 1 that fortran_l12.cu exports as Fortran symbols
void cblas_ccopy(int n, const void* x, int incx, void* y, int incy) { ccopy_(&n, (const b200_c32*)x, &incx, (b200_c32*)y, &incy); }
void cblas_zcopy(int n, const void* x, int incx, void* y, int incy) { zcopy_(&n, (const b200_c64*)x, &incx, (b200_c64*)y, &incy); }
void cblas_cswap(int n, void* x, int incx, void* y, int incy) { cswap_(&n, (b200_c32*)x, &incx, (b200_c32*)y, &incy); }
void cblas_zswap(int n, void* x, int incx, void* y, int incy) { zswap_(&n, (b200_c64*)x, &incx, (b200_c64*)y, &incy); }
void cblas_cscal(int n, const void* alpha, void* x, int incx) { cscal_(&n, (const b200_c32*)alpha, (b200_c32*)x, &incx); }
void cblas_zscal(int n, const void* alpha, void* x, int incx) { zscal_(&n, (const b200_c64*)alpha, (b200_c64*)x, &incx); }
void cblas_csscal(int n, float alpha, void* x, int incx) { csscal_(&n, &alpha, (b200_c32*)x, &incx); }
void cblas_zdscal(int n, double alpha, void* x, int incx) { zdscal_(&n, &alpha, (b200_c64*)x, &incx); }
float cblas_scasum(int n, const void* x, int incx) { return scasum_(&n, (const b200_c32*)x, &incx); }
double cblas_dzasum(int n, const void* x, int incx) { return dzasum_(&n, (const b200_c64*)x, &incx); }
}  // extern "C"
