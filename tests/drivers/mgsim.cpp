// mgsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A CUDA stream / event SIMULATOR for the multi-device drivers.  libgpublas_b200/csrc/multi_gemm.cu (partitioned ?gemm_, the
// Cholesky workload) and multi_level3.cu (partitioned ?syrk_, ?trsm_, ?trmm_) contain no kernels: they are host code that queues
// copies, kernels, event records and event waits on many streams of many devices.  Here the two files are compiled UNCHANGED with
// g++ and linked against this stand-in for the CUDA runtime and for the single-GPU launchers they call:
//   * a stream is a FIFO of operations; cudaEventRecord / cudaStreamWaitEvent have CUDA's semantics (a wait captures the event's
//     latest record at the time of the call; waiting on a never-recorded event is a no-op);
//   * nothing executes when it is queued.  At a synchronisation point the simulator runs the queued operations one at a time,
//     each time picking -- by a seeded policy: random, kernels first, or copies first -- one stream whose head operation may run
//     (a wait may run once its event has completed; the flag-polling DGEMM may run once every panel flag it polls has reached the
//     call's epoch).  Any interleaving the real hardware could choose is a possible schedule, so a missing wait shows up as a
//     kernel that reads a panel (allocated full of NaN) before its pieces have landed, a buffer overwritten before its reader
//     ran, or a deadlock;
//   * "device memory" of every device is host memory; the single-GPU kernels are OpenBLAS calls (cblas_dgemm / dtrsm / dtrmm,
//     dpotrf_) applied when their operation runs.
// The driver below calls the product's multi_gemm / multi_syrk / multi_trxm / multi_cholesky_lower for 2, 3, 4, 6 and 8 devices,
// device- and host-resident operands, under every policy, and compares with OpenBLAS on the whole problem.
// tests/test_multigpu_cpu.py builds and runs it; libb200blas.so never sees this file.
#include <cuda_runtime_api.h>
#include <cuda.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../../libgpublas_b200/csrc/runtime.h"
#include "../../libgpublas_b200/csrc/kernels.h"
#include "../../libgpublas_b200/csrc/multi_gemm.h"
#include "../../libgpublas_b200/csrc/tracker.h"

// ------------------------------------------------------------------------------------------------ the simulator
namespace sim {
struct Event { long recorded = 0, completed = 0; };
enum Kind { RUN = 0, RECORD = 1, WAIT = 2 };
struct Op { int kind; bool kernel; std::function<void()> fn; std::function<bool()> ready; Event* ev; long seq; const char* what; };
struct Stream { int dev; int id; std::deque<Op> q; };
std::vector<Stream*> streams;
std::map<int, Stream*> default_stream;      // per device
int cur_dev = 0, ndev = 8;
int policy = 0;                              // 0 random, 1 kernels first, 2 copies first
std::mt19937_64 rng(1);
long executed = 0, deadlocks = 0;

Stream* new_stream(int dev) { Stream* s = new Stream{dev, (int)streams.size(), {}}; streams.push_back(s); return s; }
Stream* resolve(cudaStream_t s) {
    if (s) return (Stream*)s;
    auto it = default_stream.find(cur_dev);
    if (it == default_stream.end()) it = default_stream.emplace(cur_dev, new_stream(cur_dev)).first;
    return it->second;
}
bool runnable(const Op& o) {
    if (o.kind == WAIT) return o.ev->completed >= o.seq;
    if (o.kind == RUN && o.ready) return o.ready();
    return true;
}
void run_all() {
    for (;;) {
        std::vector<Stream*> cand, pending;
        for (Stream* s : streams) {
            if (s->q.empty()) continue;
            pending.push_back(s);
            if (runnable(s->q.front())) cand.push_back(s);
        }
        if (pending.empty()) return;
        if (cand.empty()) {
            fprintf(stderr, "mgsim: DEADLOCK -- %zu streams hold operations, none may run\n", pending.size());
            for (Stream* s : pending) fprintf(stderr, "  stream %d (device %d): %zu ops, head = %s\n", s->id, s->dev, s->q.size(), s->q.front().what);
            deadlocks++;
            for (Stream* s : pending) s->q.clear();
            return;
        }
        // bookkeeping operations (records, satisfied waits) are free to run at once under every policy; the policy orders the work
        std::vector<Stream*> pref;
        if (policy == 1) { for (Stream* s : cand) if (s->q.front().kind == RUN && s->q.front().kernel) pref.push_back(s); }
        else if (policy == 2) { for (Stream* s : cand) if (s->q.front().kind == RUN && !s->q.front().kernel) pref.push_back(s); }
        std::vector<Stream*>& from = pref.empty() ? cand : pref;
        Stream* s = from[(size_t)(rng() % from.size())];
        Op o = std::move(s->q.front());
        s->q.pop_front();
        if (o.kind == RUN) { if (o.fn) o.fn(); }
        else if (o.kind == RECORD) { if (o.ev->completed < o.seq) o.ev->completed = o.seq; }
        executed++;
    }
}
void push(cudaStream_t st, Op o) { resolve(st)->q.push_back(std::move(o)); }

// allocations made through cudaMalloc: "device" memory (the residency classifier below asks)
std::map<const char*, size_t> device_allocs;
std::map<const char*, std::pair<size_t, int>> registered;     // harness-declared residency of host buffers
}  // namespace sim

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = sim::ndev; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { sim::cur_dev = d; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = sim::cur_dev; return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -5; return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { sim::run_all(); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { sim::run_all(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)sim::new_stream(sim::cur_dev); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)sim::new_stream(sim::cur_dev); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t) new sim::Event(); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    sim::Event* ev = (sim::Event*)e;
    sim::push(s, sim::Op{sim::RECORD, false, nullptr, nullptr, ev, ++ev->recorded, "event record"});
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    sim::Event* ev = (sim::Event*)e;
    if (ev->recorded == 0) return cudaSuccess;                  // never recorded: no-op, as in CUDA
    sim::push(s, sim::Op{sim::WAIT, false, nullptr, nullptr, ev, ev->recorded, "event wait"});
    return cudaSuccess;
}
cudaError_t cudaEventQuery(cudaEvent_t) { sim::run_all(); return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t bytes) {
    sim::run_all();
    *p = malloc(bytes ? bytes : 1);
    memset(*p, 0xFF, bytes);                                    // NaN everywhere: a panel read before its pieces landed poisons the result
    sim::device_allocs[(const char*)*p] = bytes;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) { sim::run_all(); if (p) { sim::device_allocs.erase((const char*)p); free(p); } return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t bytes) { *p = malloc(bytes ? bytes : 1); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { sim::run_all(); memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { sim::run_all(); memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t s) {
    sim::push(s, sim::Op{sim::RUN, false, [=] { memset(d, v, n); }, nullptr, nullptr, 0, "memset"});
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st) {
    sim::push(st, sim::Op{sim::RUN, false, [=] { memcpy(d, s, n); }, nullptr, nullptr, 0, "copy"});
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t st) {
    sim::push(st, sim::Op{sim::RUN, false, [=] { for (size_t r = 0; r < h; r++) memcpy((char*)d + r * dp, (const char*)s + r * sp, w); }, nullptr, nullptr, 0, "2-D copy"});
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "mgsim"; }
void tracker_enter(void) {}
void tracker_leave(void) {}
void b200_writef(int fd, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n > 0 && write(fd, buf, (size_t)(n < (int)sizeof buf ? n : (int)sizeof buf - 1)) < 0) {}
}
}

// ------------------------------------------------------------------------------------------------ OpenBLAS (the kernels' stand-in)
namespace ob {
typedef void (*dgemm_t)(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);
typedef void (*dtrxm_t)(int, int, int, int, int, int, int, double, const double*, int, double*, int);
typedef void (*dpotrf_t)(const char*, const int*, double*, const int*, int*);
dgemm_t dgemm; dtrxm_t dtrsm, dtrmm; dpotrf_t dpotrf;
enum { ColMajor = 102, NoTrans = 111, Trans = 112, ConjTrans = 113, Upper = 121, Lower = 122, NonUnit = 131, Unit = 132, Left = 141, Right = 142 };
int tr(char t) { return t == 'N' ? NoTrans : (t == 'T' ? Trans : ConjTrans); }
bool load() {
    const char* path = getenv("MGSIM_OPENBLAS");
    void* h = path ? dlopen(path, RTLD_NOW | RTLD_LOCAL) : nullptr;
    if (!h) { fprintf(stderr, "mgsim: cannot load OpenBLAS (%s): %s\n", path ? path : "MGSIM_OPENBLAS unset", dlerror()); return false; }
    dgemm = (dgemm_t)dlsym(h, "cblas_dgemm"); dtrsm = (dtrxm_t)dlsym(h, "cblas_dtrsm"); dtrmm = (dtrxm_t)dlsym(h, "cblas_dtrmm");
    dpotrf = (dpotrf_t)dlsym(h, "dpotrf_");
    return dgemm && dtrsm && dtrmm && dpotrf;
}
// C := alpha op(A) op(B) + beta C on the masked part (mask: local i >= j / i <= j, as the kernels define it)
void gemm_masked(char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                 const double* Cin, int64_t ldc, double* D, int64_t ldd, int mask) {
    if (mask == b200::MASK_FULL && Cin == D) { dgemm(ColMajor, tr(ta), tr(tb), m, n, k, alpha, A, (int)lda, B, (int)ldb, beta, D, (int)ldd); return; }
    std::vector<double> P((size_t)m * n);
    dgemm(ColMajor, tr(ta), tr(tb), m, n, k, alpha, A, (int)lda, B, (int)ldb, 0.0, P.data(), m);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < m; i++) {
            if (!(mask == b200::MASK_FULL || (mask == b200::MASK_LOWER ? i >= j : i <= j))) continue;
            D[i + (int64_t)j * ldd] = P[i + (size_t)j * m] + (beta == 0.0 ? 0.0 : beta * Cin[i + (int64_t)j * ldc]);
        }
}
}  // namespace ob

// ------------------------------------------------------------------------------------------------ the library's own runtime, mocked
namespace b200 {
Options g_opts;
Stats g_stats = {};
thread_local int last_variant = 0;
int force_variant = 0;
thread_local const char* t_call_name = "sim";
void fatal(const char* what, const char* file, int line, const char* detail) { fprintf(stderr, "mgsim fatal: %s %s:%d %s\n", what, file, line, detail); abort(); }
void ensure_init() {}
int home_device() { return 0; }
int current_device() { return sim::cur_dev; }
DeviceScope::DeviceScope(int dev) : prev_(sim::cur_dev) { sim::cur_dev = dev; }
DeviceScope::~DeviceScope() { sim::cur_dev = prev_; }
cudaStream_t current_stream() { return (cudaStream_t)sim::resolve(nullptr); }
void ws_reset() {}
void make_resident(const void*, size_t, cudaStream_t) {}
static std::map<int, void*> g_scalars, g_pinned;
void* device_scalar() { void*& p = g_scalars[sim::cur_dev]; if (!p) p = calloc(1, 4096); return p; }
void* pinned_scalar() { void*& p = g_pinned[0]; if (!p) p = calloc(1, 4096); return p; }
Residency classify(const void* p) {
    const char* c = (const char*)p;
    auto it = sim::device_allocs.upper_bound(c);
    if (it != sim::device_allocs.begin()) { --it; if (c < it->first + it->second) return RES_DEVICE; }
    auto jt = sim::registered.upper_bound(c);
    if (jt != sim::registered.begin()) { --jt; if (c < jt->first + jt->second.first) return (Residency)jt->second.second; }
    return RES_HOST_PAGEABLE;
}

// ---- the single-GPU launchers: an operation on the stream that applies OpenBLAS when it runs ----
static thread_local const uint32_t* t_af = nullptr; static thread_local const uint32_t* t_bf = nullptr;
static thread_local int t_ag = 0, t_bg = 0; static thread_local uint32_t t_epoch = 0;
void dgemm_set_panel_flags(const uint32_t* aflags, int a_group, const uint32_t* bflags, int b_group, uint32_t epoch) { t_af = aflags; t_ag = a_group; t_bf = bflags; t_bg = b_group; t_epoch = epoch; }
void dgemm_out_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                   const double* C, int64_t ldc, double* D, int64_t ldd, int mask) {
    const uint32_t* af = t_af; const uint32_t* bf = t_bf; const int ag = t_ag, bg = t_bg; const uint32_t ep = t_epoch;
    t_af = t_bf = nullptr;
    std::function<bool()> ready;
    if (af || bf)      // the flag-polling kernel: a tile may read its panels once their flags have reached the epoch
        ready = [=] {
            if (af) for (int g = 0; g * ag < m; g++) if (af[g] < ep) return false;
            if (bf) for (int h = 0; h * bg < n; h++) if (bf[h] < ep) return false;
            return true;
        };
    sim::push(s, sim::Op{sim::RUN, true, [=] { ob::gemm_masked(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, D, ldd, mask); }, ready, nullptr, 0, "dgemm (flags)"});
}
void dgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
               int64_t ldc, int mask) {
    sim::push(s, sim::Op{sim::RUN, true, [=] { ob::gemm_masked(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, C, ldc, mask); }, nullptr, nullptr, 0, "dgemm"});
}
template <> void gemm_dev<double>(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                                  double* C, int64_t ldc, int mask) { dgemm_dev(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask); }
static void trxm(bool solve, cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, double alpha, const double* A, int64_t lda, double* B, int64_t ldb) {
    sim::push(s, sim::Op{sim::RUN, true, [=] {
        (solve ? ob::dtrsm : ob::dtrmm)(ob::ColMajor, side == 'L' ? ob::Left : ob::Right, uplo == 'U' ? ob::Upper : ob::Lower, ob::tr(trans), diag == 'U' ? ob::Unit : ob::NonUnit,
                                        m, n, alpha, A, (int)lda, B, (int)ldb);
    }, nullptr, nullptr, 0, solve ? "dtrsm" : "dtrmm"});
}
template <> void trsm_dev<double>(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, double alpha, const double* A, int64_t lda, double* B, int64_t ldb) {
    trxm(true, s, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb);
}
template <> void trmm_dev<double>(cudaStream_t s, char side, char uplo, char trans, char diag, int m, int n, double alpha, const double* A, int64_t lda, double* B, int64_t ldb) {
    trxm(false, s, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb);
}
void potrf_lower_dev(cudaStream_t s, int n, double* A, int64_t lda, int* info_dev, int base) {
    sim::push(s, sim::Op{sim::RUN, true, [=] { int info = 0, ld = (int)lda; ob::dpotrf("L", &n, A, &ld, &info); if (info > 0 && *info_dev == 0) *info_dev = base + info; }, nullptr, nullptr, 0, "potrf"});
}
// the other precisions are not simulated (the orchestration is type-independent; only their symbols are needed to link)
#define MGSIM_UNSIMULATED(T)                                                                                                                              \
    template <> void gemm_dev<T>(cudaStream_t, char, char, int, int, int, T, const T*, int64_t, const T*, int64_t, T, T*, int64_t, int) { abort(); }      \
    template <> void trsm_dev<T>(cudaStream_t, char, char, char, char, int, int, T, const T*, int64_t, T*, int64_t) { abort(); }                          \
    template <> void trmm_dev<T>(cudaStream_t, char, char, char, char, int, int, T, const T*, int64_t, T*, int64_t) { abort(); }
MGSIM_UNSIMULATED(float)
MGSIM_UNSIMULATED(cuFloatComplex)
MGSIM_UNSIMULATED(cuDoubleComplex)
void sgemm_dev(cudaStream_t, char, char, int, int, int, float, const float*, int64_t, const float*, int64_t, float, float*, int64_t, int) { abort(); }
void cgemm_dev(cudaStream_t, char, char, int, int, int, cuFloatComplex, const cuFloatComplex*, int64_t, const cuFloatComplex*, int64_t, cuFloatComplex, cuFloatComplex*, int64_t, int) { abort(); }
void zgemm_dev(cudaStream_t, char, char, int, int, int, cuDoubleComplex, const cuDoubleComplex*, int64_t, const cuDoubleComplex*, int64_t, cuDoubleComplex, cuDoubleComplex*, int64_t, int) { abort(); }
}  // namespace b200

// ------------------------------------------------------------------------------------------------ the driver
static unsigned long long g_data = 88172645463325252ull;
static void fill(std::vector<double>& v, double scale = 1.0) {      // xorshift64: the matrices are large, the generator must be cheap
    for (auto& x : v) {
        g_data ^= g_data << 13; g_data ^= g_data >> 7; g_data ^= g_data << 17;
        x = scale * ((double)(g_data >> 11) * (2.0 / 9007199254740992.0) - 1.0);
    }
}
// a buffer in "device" memory (cudaMalloc: classified as device-resident) or host memory declared pinned / left pageable
struct Buf {
    double* p = nullptr; size_t n = 0; int where = 0;      // 0 device, 1 pinned host, 2 pageable host
    Buf(const std::vector<double>& src, int where_) : n(src.size()), where(where_) {
        if (where == 0) { cudaMalloc((void**)&p, n * 8); } else { p = (double*)malloc(n * 8); if (where == 1) sim::registered[(const char*)p] = {n * 8, (int)b200::RES_HOST_PINNED}; }
        memcpy(p, src.data(), n * 8);
    }
    ~Buf() { sim::run_all(); if (where == 0) cudaFree(p); else { sim::registered.erase((const char*)p); free(p); } }
};
static double max_abs_diff(const double* a, const double* b, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; i++) { const double d = std::fabs(a[i] - b[i]); if (!(d <= m)) m = d; }     // NaN propagates into m
    return m;
}
static int g_fail = 0, g_cases = 0;
static void verdict(const char* what, int ndev, int where, double err, double tol, bool partitioned) {
    g_cases++;
    const bool ok = partitioned && err <= tol && sim::deadlocks == 0;
    if (!ok) g_fail++;
    printf("%s %-28s devices=%d operands=%s policy=%d err=%.3e tol=%.1e%s\n", ok ? "ok  " : "FAIL", what, ndev, where == 0 ? "device" : (where == 1 ? "pinned" : "pageable"), sim::policy, err,
           tol, partitioned ? "" : " (NOT PARTITIONED)");
    sim::deadlocks = 0;
}

static void case_gemm(int ndev, char ta, char tb, int where) {
    int P, Q;
    b200::mg_grid(ndev, &P, &Q);
    const int m = 1024 * P + 136, n = 1024 * Q + 264, k = 1024 + 72;
    const int ra = ta == 'N' ? m : k, ca = ta == 'N' ? k : m, rb = tb == 'N' ? k : n, cb = tb == 'N' ? n : k;
    const int lda = ra + 2, ldb = rb + 4, ldc = m + 6;
    std::vector<double> A((size_t)lda * ca), B((size_t)ldb * cb), C((size_t)ldc * n);
    fill(A); fill(B); fill(C);
    std::vector<double> R = C;
    ob::dgemm(ob::ColMajor, ob::tr(ta), ob::tr(tb), m, n, k, 0.7, A.data(), lda, B.data(), ldb, 1.3, R.data(), ldc);
    Buf a(A, where), b(B, where), c(C, where);
    const unsigned long long calls0 = b200::g_mg_stats.calls;
    const bool part = b200::multi_gemm<double>(ta, tb, m, n, k, 0.7, a.p, lda, b.p, ldb, 1.3, c.p, ldc);
    sim::run_all();
    char name[64]; snprintf(name, sizeof name, "dgemm %c%c %dx%dx%d", ta, tb, m, n, k);
    verdict(name, ndev, where, max_abs_diff(c.p, R.data(), R.size()), 1e-10 * k, part && b200::g_mg_stats.calls == calls0 + 1);
}
static void case_syrk(int ndev, char uplo, char trans, double beta, int where) {
    const int n = 1024 * ndev + 200, k = 300;
    const int ra = trans == 'N' ? n : k, ca = trans == 'N' ? k : n, lda = ra + 2, ldc = n + 6;
    std::vector<double> A((size_t)lda * ca), C((size_t)ldc * n);
    fill(A); fill(C);
    std::vector<double> R = C, F((size_t)n * n);
    ob::dgemm(ob::ColMajor, trans == 'N' ? ob::NoTrans : ob::Trans, trans == 'N' ? ob::Trans : ob::NoTrans, n, n, k, 0.7, A.data(), lda, A.data(), lda, 0.0, F.data(), n);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
            if (uplo == 'L' ? i >= j : i <= j) R[i + (size_t)j * ldc] = F[i + (size_t)j * n] + beta * C[i + (size_t)j * ldc];
    Buf a(A, where), c(C, where);
    const bool part = b200::multi_syrk<double>(uplo, trans, n, k, 0.7, a.p, lda, beta, c.p, ldc);
    sim::run_all();
    char name[64]; snprintf(name, sizeof name, "dsyrk %c%c n=%d k=%d beta=%g", uplo, trans, n, k, beta);
    verdict(name, ndev, where, max_abs_diff(c.p, R.data(), R.size()), 1e-10 * k, part);      // whole array: the other triangle and the padding must be untouched
}
static void case_trxm(int ndev, bool solve, char side, char uplo, char trans, int where) {
    const int na = 2048 + 40, nfree = 512 * ndev + 70;
    const int m = side == 'L' ? na : nfree, n = side == 'L' ? nfree : na, lda = na + 2, ldb = m + 6;
    std::vector<double> A((size_t)lda * na), B((size_t)ldb * n);
    fill(A, 1.0 / na); fill(B);
    for (int i = 0; i < na; i++) A[i + (size_t)i * lda] = 2.0;
    for (int j = 0; j < na; j++)            // the unreferenced triangle must never be read
        for (int i = 0; i < na; i++)
            if (uplo == 'L' ? i < j : i > j) A[i + (size_t)j * lda] = NAN;
    std::vector<double> R = B;
    (solve ? ob::dtrsm : ob::dtrmm)(ob::ColMajor, side == 'L' ? ob::Left : ob::Right, uplo == 'U' ? ob::Upper : ob::Lower, ob::tr(trans), ob::NonUnit, m, n, 0.7, A.data(), lda, R.data(), ldb);
    Buf a(A, where), b(B, where);
    const bool part = b200::multi_trxm<double>(solve, side, uplo, trans, 'N', m, n, 0.7, a.p, lda, b.p, ldb);
    sim::run_all();
    char name[64]; snprintf(name, sizeof name, "%s %c%c%cN %dx%d", solve ? "dtrsm" : "dtrmm", side, uplo, trans, m, n);
    verdict(name, ndev, where, max_abs_diff(b.p, R.data(), R.size()), 1e-9, part);
}
static void case_cholesky(int ndev, int nb) {
    const int n = 8 * nb + 136, lda = n + 2;
    std::vector<double> G((size_t)n * n), A((size_t)lda * n, -7e9);
    fill(G);
    std::vector<double> S((size_t)n * n);
    ob::dgemm(ob::ColMajor, ob::NoTrans, ob::Trans, n, n, n, 1.0 / n, G.data(), n, G.data(), n, 0.0, S.data(), n);
    for (int j = 0; j < n; j++)
        for (int i = j; i < n; i++) A[i + (size_t)j * lda] = S[i + (size_t)j * n] + (i == j ? 1.0 : 0.0);     // lower triangle; the upper holds a rogue value
    std::vector<double> R = A;
    int info = 0;
    ob::dpotrf("L", &n, R.data(), &lda, &info);
    Buf a(A, 0);
    const int got = b200::multi_cholesky_lower(n, a.p, lda, nb, ndev);
    sim::run_all();
    char name[64]; snprintf(name, sizeof name, "cholesky n=%d nb=%d info=%d/%d", n, nb, got, info);
    verdict(name, ndev, 0, max_abs_diff(a.p, R.data(), R.size()), 1e-9, got == info);
}

int main(int argc, char** argv) {
    if (!ob::load()) { printf("RESULT skipped (no OpenBLAS)\n"); return 0; }
    b200::g_opts.multi_gpu_min_dim = 1024;
    int only = argc > 1 ? atoi(argv[1]) : 0;
    const int counts[] = {2, 3, 4, 6, 8};
    for (int nd : counts) {
        if (only && nd != only) continue;
        b200::g_opts.devices = nd;
        sim::ndev = nd > 8 ? nd : 8;
        const int npol = nd >= 6 ? 2 : 4;             // (the larger grids take longest: random and kernels-first there)
        for (int pol = 0; pol < npol; pol++) {
            sim::policy = pol == 3 ? 0 : pol;
            sim::rng.seed(1000 + 17 * nd + pol);
            case_gemm(nd, 'N', 'N', 0);
            case_gemm(nd, 'T', 'N', pol & 1 ? 1 : 2);
            if (pol == 0) case_gemm(nd, 'N', 'T', 1);
            case_syrk(nd, 'L', 'N', 1.3, 0);
            case_syrk(nd, 'U', 'T', 0.0, pol & 1 ? 2 : 1);
            if (pol == 0) { case_syrk(nd, 'U', 'N', 1.3, 0); case_syrk(nd, 'L', 'T', 0.0, 1); }
            case_trxm(nd, true, 'L', 'L', 'N', 0);
            case_trxm(nd, true, 'R', 'U', 'T', pol & 1 ? 1 : 2);
            case_trxm(nd, false, 'L', 'U', 'N', pol & 1 ? 0 : 1);
            if (pol == 0) case_trxm(nd, false, 'R', 'L', 'T', 0);
            case_cholesky(nd, 128);
            if (pol == 0) case_cholesky(nd, 256);
        }
    }
    printf("RESULT cases=%d failed=%d operations=%ld\n", g_cases, g_fail, sim::executed);
    return g_fail != 0;
}
