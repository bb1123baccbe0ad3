// level2_more.cu -- ROT and ROTG (reference blas_level1/rot.cc, rotg.cc forward to cublas<t>rot / rotg): real types s/d,
// Fortran + CBLAS entry points.  ROT is one element-wise pass; ROTG is scalar work and stays on the host.
// (GER / SYR / SYMV / TRMV, first built here on their own kernels, now run on the structured Level-2 bodies of
// level2_struct.cu: their "N part" walked upper triangles from a different column per thread, i.e. uncoalesced.)
#include "abi_common.h"
#include "../../include/b200blas.h"
#include <cmath>
#include <cstdlib>

namespace b200 {

__device__ __forceinline__ int64_t vpos(int64_t i, int64_t n, int64_t inc) { return inc >= 0 ? i * inc : (n - 1 - i) * (-inc); }

template <typename T> __global__ void rot_kernel(int64_t n, T* x, int64_t incx, T* y, int64_t incy, T c, T s) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        T* px = x + vpos(i, n, incx); T* py = y + vpos(i, n, incy);
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

}  // namespace b200

using namespace b200;

namespace {

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
#define B200_ROT(P, T)                                                                                                                        \
    void P##rot_(const int* n, T* x, const int* incx, T* y, const int* incy, const T* c, const T* s) { rot_entry<T>(#P "rot_", n, x, incx, y, incy, c, s); } \
    void P##rotg_(T* a, T* b, T* c, T* s) { rotg_host<T>(a, b, c, s); }                                                                          \
    void cblas_##P##rot(int n, T* x, int incx, T* y, int incy, T c, T s) { P##rot_(&n, x, &incx, y, &incy, &c, &s); }                             \
    void cblas_##P##rotg(T* a, T* b, T* c, T* s) { rotg_host<T>(a, b, c, s); }
B200_ROT(s, float)
B200_ROT(d, double)
}
