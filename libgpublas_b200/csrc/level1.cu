// level1.cu -- bandwidth-bound Level-1 kernels (reference blas_level1/*.cc are dead wrappers that
// forward to cublas<t>{dot,axpy,nrm2,amax,...}).
//
// All kernels: 128-bit vectorised coalesced loads on the unit-stride fast path (when both pointers
// are 16-byte aligned; a scalar path covers strides and misalignment), several independent loads in
// flight per thread, grid = a multiple of the SM count.  Reductions are deterministic: each block
// reduces with warp shuffles + shared memory, writes one partial, and the last block to finish
// (ticket counter) combines the partials in block order.
#include "common.cuh"
#include "kernels.h"
#include "gemm_generic.cuh"
#include "runtime.h"

namespace b200 {

constexpr int L1_THREADS = 256;
constexpr int L1_MAX_BLOCKS = 148 * 8;

static inline int l1_blocks(int64_t n, int per_thread) {
    int64_t b = (n + (int64_t)L1_THREADS * per_thread - 1) / ((int64_t)L1_THREADS * per_thread);
    int cap = sm_count() > 0 ? sm_count() * 8 : L1_MAX_BLOCKS;
    if (cap > L1_MAX_BLOCKS) cap = L1_MAX_BLOCKS;
    if (b < 1) b = 1;
    return (int)(b > cap ? cap : b);
}
// Grid of a grid-stride REDUCTION kernel: exactly one resident wave (SMs x CTAs that fit per SM, from the occupancy
// calculator, cached per kernel).  With 8 CTAs per SM requested and 2-5 resident, the old grid ran in 2-4 waves whose
// ramps and tails cost ~5 us of a 165 us DDOT, and the finishing block folded 1184 partials instead of ~300-700.
static int l1_reduce_grid(const void* kernel, int64_t n, int per_thread) {
    struct Entry { const void* k; int per_sm; };
    static Entry cache[64];
    static int ncache = 0;
    static const int waves = getenv("B200BLAS_L1_WAVES") ? atoi(getenv("B200BLAS_L1_WAVES")) : 1;
    int per_sm = 0;
    for (int i = 0; i < ncache; i++) if (cache[i].k == kernel) per_sm = cache[i].per_sm;
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, L1_THREADS, 0) != cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 2; }
        if (ncache < 64) { cache[ncache].k = kernel; cache[ncache].per_sm = per_sm; __atomic_fetch_add(&ncache, 1, __ATOMIC_RELEASE); }
    }
    int64_t b = (n + (int64_t)L1_THREADS * per_thread - 1) / ((int64_t)L1_THREADS * per_thread);
    int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * per_sm * (waves > 0 ? waves : 1);
    if (cap > L1_MAX_BLOCKS) cap = L1_MAX_BLOCKS;
    if (b < 1) b = 1;
    return (int)(b > cap ? cap : b);
}

// BLAS element i of a strided vector (negative increments walk backwards from the end)
__device__ __forceinline__ int64_t vix(int64_t i, int64_t n, int64_t inc) { return inc >= 0 ? i * inc : (n - 1 - i) * (-inc); }

// streaming 128-bit load (read once: do not allocate in L1)
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------------
// generic deterministic block + grid reduction.  V must be trivially copyable; Op::combine(a,b)
// must be associative enough for the caller (sum / argmax); order is fixed.
template <typename V, typename Op>
__device__ __forceinline__ V block_reduce(V v, V* smem /* >= 32 entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Op::combine(v, Op::shfl_down(v, o));
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (blockDim.x >> 5) ? smem[lane] : Op::identity();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = Op::combine(v, Op::shfl_down(v, o));
    }
    return v;   // valid in thread 0
}

template <typename V, typename Op, typename Fin>
__device__ __forceinline__ void grid_finish(V block_val, V* partials, unsigned int* ticket, V* smem, Fin fin) {
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = block_val;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed order: thread t folds partials t, t+T, t+2T, ... then the block tree combines in lane order
    V v = Op::identity();
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v = Op::combine(v, *((volatile V*)&partials[i]));
    __syncthreads();
    v = block_reduce<V, Op>(v, smem);
    if (threadIdx.x == 0) {
        // Ticket first, result second: the host acts on the result the moment it sees it (wait_scalar spins on the zero-copy
        // slot), possibly launching the next reduction on another stream; by then the ticket must already read zero.  Every
        // other block has taken its ticket (this block drew the last one), so nobody else still needs the old value.
        *ticket = 0;
        __threadfence_system();
        fin(v);
    }
}

template <typename T> struct SumOp {
    static __device__ T identity() { return T(0); }
    static __device__ T combine(T a, T b) { return a + b; }
    static __device__ T shfl_down(T v, int o) { return __shfl_down_sync(0xffffffffu, v, o); }
};
struct Sum2 { double x, y; };
struct Sum2Op {
    static __device__ Sum2 identity() { return Sum2{0.0, 0.0}; }
    static __device__ Sum2 combine(Sum2 a, Sum2 b) { return Sum2{a.x + b.x, a.y + b.y}; }
    static __device__ Sum2 shfl_down(Sum2 v, int o) { return Sum2{__shfl_down_sync(0xffffffffu, v.x, o), __shfl_down_sync(0xffffffffu, v.y, o)}; }
};
struct Sum3 { double a, b, c; };
struct Sum3Op {
    static __device__ Sum3 identity() { return Sum3{0.0, 0.0, 0.0}; }
    static __device__ Sum3 combine(Sum3 p, Sum3 q) { return Sum3{p.a + q.a, p.b + q.b, p.c + q.c}; }
    static __device__ Sum3 shfl_down(Sum3 v, int o) {
        return Sum3{__shfl_down_sync(0xffffffffu, v.a, o), __shfl_down_sync(0xffffffffu, v.b, o), __shfl_down_sync(0xffffffffu, v.c, o)};
    }
};
struct ArgMax { double v; long long i; };   // i < 0: empty
struct ArgMaxOp {
    static __device__ ArgMax identity() { return ArgMax{-1.0, -1}; }
    // larger |x| wins; ties -> smaller index (netlib I?AMAX returns the FIRST maximal element)
    static __device__ ArgMax combine(ArgMax a, ArgMax b) {
        if (b.i < 0) return a;
        if (a.i < 0) return b;
        if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
        return a;
    }
    static __device__ ArgMax shfl_down(ArgMax v, int o) { return ArgMax{__shfl_down_sync(0xffffffffu, v.v, o), __shfl_down_sync(0xffffffffu, v.i, o)}; }
};

struct L1Scratch { void* partials; unsigned int* ticket; };
static L1Scratch l1_scratch(size_t elem) {
    L1Scratch s;
    s.partials = ws_alloc(L1_MAX_BLOCKS * elem);
    s.ticket = (unsigned int*)((char*)device_scalar() + 128);
    return s;
}

// ------------------------------------------ DOT ------------------------------------------
template <bool CONJ> __device__ __forceinline__ void cfma(Sum2& acc, double ar, double ai, double br, double bi) {
    // acc += (CONJ ? conj(a) : a) * b
    if (CONJ) { acc.x = fma(ar, br, acc.x); acc.x = fma(ai, bi, acc.x); acc.y = fma(ar, bi, acc.y); acc.y = fma(-ai, br, acc.y); }
    else      { acc.x = fma(ar, br, acc.x); acc.x = fma(-ai, bi, acc.x); acc.y = fma(ar, bi, acc.y); acc.y = fma(ai, br, acc.y); }
}

__global__ void __launch_bounds__(L1_THREADS, 4) ddot_kernel(int64_t n, const double* __restrict__ x, int64_t incx,
                                                         const double* __restrict__ y, int64_t incy, double* partials,
                                                         unsigned int* ticket, double* out, bool vec) {
    __shared__ double sm[32];
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        const double2* x2 = (const double2*)x; const double2* y2 = (const double2*)y;
        const int64_t n2 = n >> 1;
        int64_t i = tid;
#pragma unroll 1
        for (; i + 3 * nth < n2; i += 4 * nth) {
            double2 xa = ldg_stream(x2 + i), xb = ldg_stream(x2 + i + nth), xc = ldg_stream(x2 + i + 2 * nth), xd = ldg_stream(x2 + i + 3 * nth);
            double2 ya = ldg_stream(y2 + i), yb = ldg_stream(y2 + i + nth), yc = ldg_stream(y2 + i + 2 * nth), yd = ldg_stream(y2 + i + 3 * nth);
            a0 = fma(xa.x, ya.x, a0); a0 = fma(xa.y, ya.y, a0);
            a1 = fma(xb.x, yb.x, a1); a1 = fma(xb.y, yb.y, a1);
            a2 = fma(xc.x, yc.x, a2); a2 = fma(xc.y, yc.y, a2);
            a3 = fma(xd.x, yd.x, a3); a3 = fma(xd.y, yd.y, a3);
        }
        for (; i < n2; i += nth) { double2 xa = x2[i], ya = y2[i]; a0 = fma(xa.x, ya.x, a0); a0 = fma(xa.y, ya.y, a0); }
        if ((n & 1) && tid == 0) a1 = fma(x[n - 1], y[n - 1], a1);
    } else {
        for (int64_t i = tid; i < n; i += nth) a0 = fma(x[vix(i, n, incx)], y[vix(i, n, incy)], a0);
    }
    double v = block_reduce<double, SumOp<double>>((a0 + a1) + (a2 + a3), sm);
    grid_finish<double, SumOp<double>>(v, partials, ticket, sm, [=](double r) { *out = r; });
}

template <typename OutT>
__global__ void __launch_bounds__(L1_THREADS) sdot_kernel(int64_t n, const float* __restrict__ x, int64_t incx,
                                                         const float* __restrict__ y, int64_t incy, double* partials,
                                                         unsigned int* ticket, OutT* out, bool vec, double addend) {
    // float inputs, double accumulation (tighter than the CPU BLAS's float sum; exactly what netlib DSDOT / SDSDOT specify)
    __shared__ double sm[32];
    double a0 = 0, a1 = 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        const float4* x4 = (const float4*)x; const float4* y4 = (const float4*)y;
        const int64_t n4 = n >> 2;
        int64_t i = tid;
        for (; i + nth < n4; i += 2 * nth) {
            float4 xa = ldg_stream(x4 + i), xb = ldg_stream(x4 + i + nth), ya = ldg_stream(y4 + i), yb = ldg_stream(y4 + i + nth);
            a0 += (double)xa.x * ya.x + (double)xa.y * ya.y + (double)xa.z * ya.z + (double)xa.w * ya.w;
            a1 += (double)xb.x * yb.x + (double)xb.y * yb.y + (double)xb.z * yb.z + (double)xb.w * yb.w;
        }
        for (; i < n4; i += nth) { float4 xa = x4[i], ya = y4[i]; a0 += (double)xa.x * ya.x + (double)xa.y * ya.y + (double)xa.z * ya.z + (double)xa.w * ya.w; }
        if (tid == 0) for (int64_t r = n4 << 2; r < n; r++) a1 += (double)x[r] * y[r];
    } else {
        for (int64_t i = tid; i < n; i += nth) a0 += (double)x[vix(i, n, incx)] * y[vix(i, n, incy)];
    }
    double v = block_reduce<double, SumOp<double>>(a0 + a1, sm);
    grid_finish<double, SumOp<double>>(v, partials, ticket, sm, [=](double r) { *out = (OutT)(r + addend); });
}

template <typename CT, bool CONJ>
__global__ void __launch_bounds__(L1_THREADS) cdot_kernel(int64_t n, const CT* __restrict__ x, int64_t incx, const CT* __restrict__ y,
                                                         int64_t incy, Sum2* partials, unsigned int* ticket, CT* out) {
    __shared__ Sum2 sm[32];
    Sum2 a = {0.0, 0.0}, b = {0.0, 0.0};
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    int64_t i = tid;
    for (; i + nth < n; i += 2 * nth) {
        CT xa = x[vix(i, n, incx)], ya = y[vix(i, n, incy)], xb = x[vix(i + nth, n, incx)], yb = y[vix(i + nth, n, incy)];
        cfma<CONJ>(a, xa.x, xa.y, ya.x, ya.y);
        cfma<CONJ>(b, xb.x, xb.y, yb.x, yb.y);
    }
    for (; i < n; i += nth) { CT xa = x[vix(i, n, incx)], ya = y[vix(i, n, incy)]; cfma<CONJ>(a, xa.x, xa.y, ya.x, ya.y); }
    Sum2 v = block_reduce<Sum2, Sum2Op>(Sum2{a.x + b.x, a.y + b.y}, sm);
    grid_finish<Sum2, Sum2Op>(v, partials, ticket, sm, [=](Sum2 r) { out->x = r.x; out->y = r.y; });
}

static inline bool vec_ok(const void* a, const void* b, int64_t inca, int64_t incb) {
    return inca == 1 && incb == 1 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
}

template <> void dot_dev<double>(cudaStream_t s, int64_t n, const double* x, int64_t incx, const double* y, int64_t incy, double* out, bool) {
    L1Scratch sc = l1_scratch(sizeof(double));
    ddot_kernel<<<l1_reduce_grid((const void*)ddot_kernel, n, 8), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (double*)sc.partials, sc.ticket, out, vec_ok(x, y, incx, incy));
}
template <> void dot_dev<float>(cudaStream_t s, int64_t n, const float* x, int64_t incx, const float* y, int64_t incy, float* out, bool) {
    L1Scratch sc = l1_scratch(sizeof(double));
    sdot_kernel<float><<<l1_blocks(n, 8), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (double*)sc.partials, sc.ticket, out, vec_ok(x, y, incx, incy), 0.0);
}
// netlib DSDOT (double result) / SDSDOT (sb + dot, rounded to float once): float operands, double accumulation
void dsdot_dev(cudaStream_t s, int64_t n, const float* x, int64_t incx, const float* y, int64_t incy, double sb, void* out, bool out_double) {
    L1Scratch sc = l1_scratch(sizeof(double));
    if (out_double) sdot_kernel<double><<<l1_blocks(n, 8), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (double*)sc.partials, sc.ticket, (double*)out, vec_ok(x, y, incx, incy), sb);
    else sdot_kernel<float><<<l1_blocks(n, 8), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (double*)sc.partials, sc.ticket, (float*)out, vec_ok(x, y, incx, incy), sb);
}
template <> void dot_dev<cuDoubleComplex>(cudaStream_t s, int64_t n, const cuDoubleComplex* x, int64_t incx, const cuDoubleComplex* y,
                                          int64_t incy, cuDoubleComplex* out, bool conj_x) {
    L1Scratch sc = l1_scratch(sizeof(Sum2));
    if (conj_x) cdot_kernel<cuDoubleComplex, true><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (Sum2*)sc.partials, sc.ticket, out);
    else        cdot_kernel<cuDoubleComplex, false><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (Sum2*)sc.partials, sc.ticket, out);
}
template <> void dot_dev<cuFloatComplex>(cudaStream_t s, int64_t n, const cuFloatComplex* x, int64_t incx, const cuFloatComplex* y,
                                         int64_t incy, cuFloatComplex* out, bool conj_x) {
    L1Scratch sc = l1_scratch(sizeof(Sum2));
    if (conj_x) cdot_kernel<cuFloatComplex, true><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (Sum2*)sc.partials, sc.ticket, out);
    else        cdot_kernel<cuFloatComplex, false><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, x, incx, y, incy, (Sum2*)sc.partials, sc.ticket, out);
}

// ------------------------------------------ NRM2 / ASUM ------------------------------------------
// Blue's one-pass scaled accumulation (the algorithm of netlib's current DNRM2): three sums with
// scaling constants keep the sum of squares in range for any finite input.
constexpr double BLUE_TSML = 1.4916681462400413e-154;   // 2^-511
constexpr double BLUE_TBIG = 1.9979190722022350e+146;   // 2^486
constexpr double BLUE_SSML = 4.4989137945431964e+161;   // 2^537
constexpr double BLUE_SBIG = 1.1113793747425387e-162;   // 2^-538

__device__ __forceinline__ void blue_add(Sum3& acc, double v) {
    double a = fabs(v);
    if (a > BLUE_TBIG) { double t = a * BLUE_SBIG; acc.c = fma(t, t, acc.c); }
    else if (a < BLUE_TSML) { double t = a * BLUE_SSML; acc.a = fma(t, t, acc.a); }
    else acc.b = fma(a, a, acc.b);
}
__device__ __forceinline__ double blue_finish(Sum3 r) {
    // netlib DNRM2 combination of the three accumulators
    double scl, sumsq;
    if (r.c > 0.0) {
        if (r.b > 0.0 || r.b != r.b) r.c += (r.b * BLUE_SBIG) * BLUE_SBIG;
        scl = 1.0 / BLUE_SBIG; sumsq = r.c;
    } else if (r.a > 0.0) {
        if (r.b > 0.0 || r.b != r.b) {
            double amed = sqrt(r.b), asml = sqrt(r.a) / BLUE_SSML;
            double ymin = amed < asml ? amed : asml, ymax = amed < asml ? asml : amed;
            scl = 1.0; sumsq = ymax * ymax * (1.0 + (ymin / ymax) * (ymin / ymax));
        } else { scl = 1.0 / BLUE_SSML; sumsq = r.a; }
    } else { scl = 1.0; sumsq = r.b; }
    return scl * sqrt(sumsq);
}

// NC = real components per element (1 real, 2 complex); T = component type
template <typename T, int NC, typename R>
__global__ void __launch_bounds__(L1_THREADS) nrm2_kernel(int64_t n, const T* __restrict__ x, int64_t incx, Sum3* partials,
                                                         unsigned int* ticket, R* out, bool vec) {
    __shared__ Sum3 sm[32];
    Sum3 a = {0, 0, 0}, b = {0, 0, 0};
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (vec && sizeof(T) == 8) {
        const double2* x2 = (const double2*)x;
        const int64_t tot = n * NC, n2 = tot >> 1;
        int64_t i = tid;
#pragma unroll 1
        for (; i + 7 * nth < n2; i += 8 * nth) {       // 8 independent 16-byte loads in flight per thread
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) v[u] = ldg_stream(x2 + i + u * nth);
#pragma unroll
            for (int u = 0; u < 8; u += 2) { blue_add(a, v[u].x); blue_add(a, v[u].y); blue_add(b, v[u + 1].x); blue_add(b, v[u + 1].y); }
        }
        for (; i + 3 * nth < n2; i += 4 * nth) {
            double2 p = ldg_stream(x2 + i), q = ldg_stream(x2 + i + nth), r = ldg_stream(x2 + i + 2 * nth), t = ldg_stream(x2 + i + 3 * nth);
            blue_add(a, p.x); blue_add(a, p.y); blue_add(b, q.x); blue_add(b, q.y);
            blue_add(a, r.x); blue_add(a, r.y); blue_add(b, t.x); blue_add(b, t.y);
        }
        for (; i < n2; i += nth) { double2 p = x2[i]; blue_add(a, p.x); blue_add(a, p.y); }
        if ((tot & 1) && tid == 0) blue_add(b, (double)x[tot - 1]);
    } else {
        for (int64_t i = tid; i < n; i += nth) {
            const T* p = x + i * incx * NC;
#pragma unroll
            for (int c = 0; c < NC; c++) blue_add(a, (double)p[c]);
        }
    }
    Sum3 v = block_reduce<Sum3, Sum3Op>(Sum3{a.a + b.a, a.b + b.b, a.c + b.c}, sm);
    grid_finish<Sum3, Sum3Op>(v, partials, ticket, sm, [=](Sum3 r) { *out = (R)blue_finish(r); });
}

template <typename T, int NC, typename R>
__global__ void __launch_bounds__(L1_THREADS) asum_kernel(int64_t n, const T* __restrict__ x, int64_t incx, double* partials,
                                                         unsigned int* ticket, R* out) {
    __shared__ double sm[32];
    double a = 0, b = 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    int64_t i = tid;
    for (; i + nth < n; i += 2 * nth) {
        const T* p = x + i * incx * NC; const T* q = x + (i + nth) * incx * NC;
#pragma unroll
        for (int c = 0; c < NC; c++) { a += fabs((double)p[c]); b += fabs((double)q[c]); }
    }
    for (; i < n; i += nth) {
        const T* p = x + i * incx * NC;
#pragma unroll
        for (int c = 0; c < NC; c++) a += fabs((double)p[c]);
    }
    double v = block_reduce<double, SumOp<double>>(a + b, sm);
    grid_finish<double, SumOp<double>>(v, partials, ticket, sm, [=](double r) { *out = (R)r; });
}

#define B200_NRM2(T, CT, NC, R)                                                                                          \
    template <> void nrm2_dev<T, R>(cudaStream_t s, int64_t n, const T* x, int64_t incx, R* out) {                      \
        L1Scratch sc = l1_scratch(sizeof(Sum3));                                                                         \
        nrm2_kernel<CT, NC, R><<<l1_reduce_grid((const void*)nrm2_kernel<CT, NC, R>, n* NC, 8), L1_THREADS, 0, s>>>(n, (const CT*)x, incx, (Sum3*)sc.partials, sc.ticket, out, \
                                                                          incx == 1 && (uintptr_t)x % 16 == 0);          \
    }                                                                                                                    \
    template <> void asum_dev<T, R>(cudaStream_t s, int64_t n, const T* x, int64_t incx, R* out) {                      \
        L1Scratch sc = l1_scratch(sizeof(double));                                                                       \
        asum_kernel<CT, NC, R><<<l1_blocks(n* NC, 8), L1_THREADS, 0, s>>>(n, (const CT*)x, incx, (double*)sc.partials, sc.ticket, out); \
    }
B200_NRM2(double, double, 1, double)
B200_NRM2(float, float, 1, float)
B200_NRM2(cuDoubleComplex, double, 2, double)
B200_NRM2(cuFloatComplex, float, 2, float)
#undef B200_NRM2

// ------------------------------------------ I?AMAX ------------------------------------------
template <typename T, int NC>
__global__ void __launch_bounds__(L1_THREADS, 4) iamax_kernel(int64_t n, const T* __restrict__ x, int64_t incx, ArgMax* partials,
                                                          unsigned int* ticket, long long* out, bool vec) {
    __shared__ ArgMax sm[32];
    ArgMax best = {-1.0, -1};
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    // each thread visits its elements in increasing index order and keeps strictly larger values only,
    // so it holds the first maximum of its subsequence; ties across threads resolve to the smaller index.
    if (vec && NC == 1 && sizeof(T) == 8) {
        const double2* x2 = (const double2*)x;
        const int64_t n2 = n >> 1;
        int64_t i = tid;
#pragma unroll 1
        for (; i + 3 * nth < n2; i += 4 * nth) {
            double2 p = ldg_stream(x2 + i), q = ldg_stream(x2 + i + nth), r = ldg_stream(x2 + i + 2 * nth), t = ldg_stream(x2 + i + 3 * nth);
            double v;
            v = fabs(p.x); if (v > best.v) { best.v = v; best.i = 2 * i; }
            v = fabs(p.y); if (v > best.v) { best.v = v; best.i = 2 * i + 1; }
            v = fabs(q.x); if (v > best.v) { best.v = v; best.i = 2 * (i + nth); }
            v = fabs(q.y); if (v > best.v) { best.v = v; best.i = 2 * (i + nth) + 1; }
            v = fabs(r.x); if (v > best.v) { best.v = v; best.i = 2 * (i + 2 * nth); }
            v = fabs(r.y); if (v > best.v) { best.v = v; best.i = 2 * (i + 2 * nth) + 1; }
            v = fabs(t.x); if (v > best.v) { best.v = v; best.i = 2 * (i + 3 * nth); }
            v = fabs(t.y); if (v > best.v) { best.v = v; best.i = 2 * (i + 3 * nth) + 1; }
        }
        for (; i < n2; i += nth) {
            double2 p = x2[i]; double v;
            v = fabs(p.x); if (v > best.v) { best.v = v; best.i = 2 * i; }
            v = fabs(p.y); if (v > best.v) { best.v = v; best.i = 2 * i + 1; }
        }
        if ((n & 1) && tid == 0) { double v = fabs((double)x[n - 1]); if (v > best.v) { best.v = v; best.i = n - 1; } }
    } else {
        for (int64_t i = tid; i < n; i += nth) {
            const T* p = x + i * incx * NC;
            double v = fabs((double)p[0]);
            if (NC == 2) v += fabs((double)p[NC - 1]);   // netlib DCABS1: |re| + |im|
            if (v > best.v) { best.v = v; best.i = i; }
        }
    }
    ArgMax v = block_reduce<ArgMax, ArgMaxOp>(best, sm);
    grid_finish<ArgMax, ArgMaxOp>(v, partials, ticket, sm, [=](ArgMax r) { *out = r.i; });
}
#define B200_IAMAX(T, CT, NC)                                                                                  \
    template <> void iamax_dev<T>(cudaStream_t s, int64_t n, const T* x, int64_t incx, long long* out) {       \
        L1Scratch sc = l1_scratch(sizeof(ArgMax));                                                              \
        iamax_kernel<CT, NC><<<l1_reduce_grid((const void*)iamax_kernel<CT, NC>, n, 8), L1_THREADS, 0, s>>>(n, (const CT*)x, incx, (ArgMax*)sc.partials, sc.ticket, out, \
                                                                    incx == 1 && (uintptr_t)x % 16 == 0);       \
    }
B200_IAMAX(double, double, 1)
B200_IAMAX(float, float, 1)
B200_IAMAX(cuDoubleComplex, double, 2)
B200_IAMAX(cuFloatComplex, float, 2)
#undef B200_IAMAX

// ------------------------------------------ I?AMIN ------------------------------------------
// reference blas_level1/amin.cc:10-56 (cublasI<t>amin): index of the first element of smallest |x| (|re|+|im| for complex)
struct ArgMinOp {
    static __device__ ArgMax identity() { return ArgMax{0.0, -1}; }
    static __device__ ArgMax combine(ArgMax a, ArgMax b) {
        if (b.i < 0) return a;
        if (a.i < 0) return b;
        if (b.v < a.v || (b.v == a.v && b.i < a.i)) return b;
        return a;
    }
    static __device__ ArgMax shfl_down(ArgMax v, int o) { return ArgMax{__shfl_down_sync(0xffffffffu, v.v, o), __shfl_down_sync(0xffffffffu, v.i, o)}; }
};
template <typename T, int NC>
__global__ void __launch_bounds__(L1_THREADS) iamin_kernel(int64_t n, const T* __restrict__ x, int64_t incx, ArgMax* partials, unsigned int* ticket,
                                                          long long* out) {
    __shared__ ArgMax sm[32];
    ArgMax best = {0.0, -1};
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n; i += nth) {   // increasing index per thread: strictly smaller values only => first minimum
        const T* p = x + i * incx * NC;
        double v = fabs((double)p[0]);
        if (NC == 2) v += fabs((double)p[NC - 1]);
        if (best.i < 0 || v < best.v) { best.v = v; best.i = i; }
    }
    ArgMax v = block_reduce<ArgMax, ArgMinOp>(best, sm);
    grid_finish<ArgMax, ArgMinOp>(v, partials, ticket, sm, [=](ArgMax r) { *out = r.i; });
}
#define B200_IAMIN(T, CT, NC)                                                                                  \
    template <> void iamin_dev<T>(cudaStream_t s, int64_t n, const T* x, int64_t incx, long long* out) {       \
        L1Scratch sc = l1_scratch(sizeof(ArgMax));                                                              \
        iamin_kernel<CT, NC><<<l1_blocks(n, 8), L1_THREADS, 0, s>>>(n, (const CT*)x, incx, (ArgMax*)sc.partials, sc.ticket, out); \
    }
B200_IAMIN(double, double, 1)
B200_IAMIN(float, float, 1)
B200_IAMIN(cuDoubleComplex, double, 2)
B200_IAMIN(cuFloatComplex, float, 2)
#undef B200_IAMIN

// ------------------------------------------ AXPY / SCAL / COPY / SWAP ------------------------------------------
__global__ void __launch_bounds__(L1_THREADS) daxpy_vec_kernel(int64_t n, double alpha, const double* __restrict__ x, double* __restrict__ y) {
    const double2* x2 = (const double2*)x; double2* y2 = (double2*)y;
    const int64_t n2 = n >> 1, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    int64_t i = tid;
    for (; i + 3 * nth < n2; i += 4 * nth) {
        double2 xa = ldg_stream(x2 + i), xb = ldg_stream(x2 + i + nth), xc = ldg_stream(x2 + i + 2 * nth), xd = ldg_stream(x2 + i + 3 * nth);
        double2 ya = y2[i], yb = y2[i + nth], yc = y2[i + 2 * nth], yd = y2[i + 3 * nth];
        ya.x = fma(alpha, xa.x, ya.x); ya.y = fma(alpha, xa.y, ya.y); yb.x = fma(alpha, xb.x, yb.x); yb.y = fma(alpha, xb.y, yb.y);
        yc.x = fma(alpha, xc.x, yc.x); yc.y = fma(alpha, xc.y, yc.y); yd.x = fma(alpha, xd.x, yd.x); yd.y = fma(alpha, xd.y, yd.y);
        y2[i] = ya; y2[i + nth] = yb; y2[i + 2 * nth] = yc; y2[i + 3 * nth] = yd;
    }
    for (; i < n2; i += nth) { double2 xa = x2[i], ya = y2[i]; ya.x = fma(alpha, xa.x, ya.x); ya.y = fma(alpha, xa.y, ya.y); y2[i] = ya; }
    if ((n & 1) && tid == 0) y[n - 1] = fma(alpha, x[n - 1], y[n - 1]);
}
template <typename T>
__global__ void __launch_bounds__(L1_THREADS) axpy_kernel(int64_t n, T alpha, const T* __restrict__ x, int64_t incx, T* __restrict__ y, int64_t incy) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        T* p = y + vix(i, n, incy);
        *p = num<T>::fma(alpha, x[vix(i, n, incx)], *p);
    }
}
template <typename T> void axpy_dev(cudaStream_t s, int64_t n, T alpha, const T* x, int64_t incx, T* y, int64_t incy) {
    axpy_kernel<T><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, alpha, x, incx, y, incy);
}
template <> void axpy_dev<double>(cudaStream_t s, int64_t n, double alpha, const double* x, int64_t incx, double* y, int64_t incy) {
    if (vec_ok(x, y, incx, incy)) daxpy_vec_kernel<<<l1_blocks(n, 8), L1_THREADS, 0, s>>>(n, alpha, x, y);
    else axpy_kernel<double><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, alpha, x, incx, y, incy);
}
template void axpy_dev<float>(cudaStream_t, int64_t, float, const float*, int64_t, float*, int64_t);
template void axpy_dev<cuFloatComplex>(cudaStream_t, int64_t, cuFloatComplex, const cuFloatComplex*, int64_t, cuFloatComplex*, int64_t);
template void axpy_dev<cuDoubleComplex>(cudaStream_t, int64_t, cuDoubleComplex, const cuDoubleComplex*, int64_t, cuDoubleComplex*, int64_t);

template <typename T, typename S> __device__ __forceinline__ T scal_mul(S a, T v);
template <> __device__ __forceinline__ float scal_mul(float a, float v) { return a * v; }
template <> __device__ __forceinline__ double scal_mul(double a, double v) { return a * v; }
template <> __device__ __forceinline__ cuFloatComplex scal_mul(cuFloatComplex a, cuFloatComplex v) { return num<cuFloatComplex>::mul(a, v); }
template <> __device__ __forceinline__ cuDoubleComplex scal_mul(cuDoubleComplex a, cuDoubleComplex v) { return num<cuDoubleComplex>::mul(a, v); }
template <> __device__ __forceinline__ cuFloatComplex scal_mul(float a, cuFloatComplex v) { return make_cuFloatComplex(a * v.x, a * v.y); }
template <> __device__ __forceinline__ cuDoubleComplex scal_mul(double a, cuDoubleComplex v) { return make_cuDoubleComplex(a * v.x, a * v.y); }

template <typename T, typename S>
__global__ void __launch_bounds__(L1_THREADS) scal_kernel(int64_t n, S alpha, T* __restrict__ x, int64_t incx) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) x[i * incx] = scal_mul<T, S>(alpha, x[i * incx]);
}
template <typename T, typename S> void scal_dev(cudaStream_t s, int64_t n, S alpha, T* x, int64_t incx) {
    scal_kernel<T, S><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, alpha, x, incx);
}
template void scal_dev<float, float>(cudaStream_t, int64_t, float, float*, int64_t);
template void scal_dev<double, double>(cudaStream_t, int64_t, double, double*, int64_t);
template void scal_dev<cuFloatComplex, cuFloatComplex>(cudaStream_t, int64_t, cuFloatComplex, cuFloatComplex*, int64_t);
template void scal_dev<cuDoubleComplex, cuDoubleComplex>(cudaStream_t, int64_t, cuDoubleComplex, cuDoubleComplex*, int64_t);
template void scal_dev<cuFloatComplex, float>(cudaStream_t, int64_t, float, cuFloatComplex*, int64_t);
template void scal_dev<cuDoubleComplex, double>(cudaStream_t, int64_t, double, cuDoubleComplex*, int64_t);

template <typename T>
__global__ void __launch_bounds__(L1_THREADS) copy_kernel(int64_t n, const T* __restrict__ x, int64_t incx, T* __restrict__ y, int64_t incy) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) y[vix(i, n, incy)] = x[vix(i, n, incx)];
}
template <typename T> void copy_dev(cudaStream_t s, int64_t n, const T* x, int64_t incx, T* y, int64_t incy) {
    if (incx == 1 && incy == 1) { B200_CUDA(cudaMemcpyAsync(y, x, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s)); return; }
    copy_kernel<T><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, x, incx, y, incy);
}
template void copy_dev<float>(cudaStream_t, int64_t, const float*, int64_t, float*, int64_t);
template void copy_dev<double>(cudaStream_t, int64_t, const double*, int64_t, double*, int64_t);
template void copy_dev<cuFloatComplex>(cudaStream_t, int64_t, const cuFloatComplex*, int64_t, cuFloatComplex*, int64_t);
template void copy_dev<cuDoubleComplex>(cudaStream_t, int64_t, const cuDoubleComplex*, int64_t, cuDoubleComplex*, int64_t);

template <typename T>
__global__ void __launch_bounds__(L1_THREADS) swap_kernel(int64_t n, T* __restrict__ x, int64_t incx, T* __restrict__ y, int64_t incy) {
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += nth) {
        T* p = x + vix(i, n, incx); T* q = y + vix(i, n, incy);
        T t = *p; *p = *q; *q = t;
    }
}
template <typename T> void swap_dev(cudaStream_t s, int64_t n, T* x, int64_t incx, T* y, int64_t incy) {
    swap_kernel<T><<<l1_blocks(n, 4), L1_THREADS, 0, s>>>(n, x, incx, y, incy);
}
template void swap_dev<float>(cudaStream_t, int64_t, float*, int64_t, float*, int64_t);
template void swap_dev<double>(cudaStream_t, int64_t, double*, int64_t, double*, int64_t);
template void swap_dev<cuFloatComplex>(cudaStream_t, int64_t, cuFloatComplex*, int64_t, cuFloatComplex*, int64_t);
template void swap_dev<cuDoubleComplex>(cudaStream_t, int64_t, cuDoubleComplex*, int64_t, cuDoubleComplex*, int64_t);

}  // namespace b200
