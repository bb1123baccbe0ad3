// stagesim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The host-staging pipelines -- staged_gemm.cuh (gemm_pipelined, gemm_first_touch), staged_level3.cuh (syrk_pipelined,
// trxm_pipelined) and the pageable bounce ring host_stager.cu -- under the CUDA stream / event simulator of simcuda.inc (see
// mgsim.cpp): three streams per call (compute, H2D, D2H) ordered only by events, executed here in adversarial orders (random,
// kernels first, copies first) with OpenBLAS as the kernels.  A chunk multiplied before its copy has landed, a panel returned
// before its multiply, a bounce slot re-packed while in flight: wrong result on the CPU.  Staged device copies live in memory that
// starts as NaN.  tests/test_host_stager_cpu.py builds and runs it; libb200blas.so never sees this file.
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

#include "simcuda.inc"

// ---- what the staging headers need beyond simcuda.inc ----
static std::map<const char*, std::pair<size_t, int>> g_managed;      // tracked managed blocks: size, resident flag
extern "C" {
cudaError_t cudaEventSynchronize(cudaEvent_t) { sim::run_all(); return cudaSuccess; }
cudaError_t cudaMemPrefetchAsync(const void*, size_t, int, cudaStream_t s) {
    sim::push(s, sim::Op{sim::RUN, false, nullptr, nullptr, nullptr, 0, "prefetch"});     // residency is a performance matter: nothing to model
    return cudaSuccess;
}
cudaError_t cudaMemAdvise(const void*, size_t, cudaMemoryAdvise, int) { return cudaSuccess; }
int tracker_lookup(const void* p, void** base, size_t* size) {
    const char* c = (const char*)p;
    auto it = g_managed.upper_bound(c);
    if (it == g_managed.begin()) return 0;
    --it;
    if (c >= it->first + it->second.first) return 0;
    if (base) *base = (void*)it->first;
    if (size) *size = it->second.first;
    return 1;
}
int tracker_peek_resident(const void* p) {
    const char* c = (const char*)p;
    auto it = g_managed.upper_bound(c);
    if (it == g_managed.begin()) return -1;
    --it;
    return c < it->first + it->second.first ? it->second.second : -1;
}
int tracker_test_and_set_resident(const void* p) {
    const char* c = (const char*)p;
    auto it = g_managed.upper_bound(c);
    if (it == g_managed.begin()) return -1;
    --it;
    if (c >= it->first + it->second.first) return -1;
    const int prev = it->second.second;
    it->second.second = 1;
    return prev;
}
}
namespace b200 {
static std::vector<void*> g_ws;
void* ws_alloc(size_t bytes) { void* p; cudaMalloc(&p, bytes); g_ws.push_back(p); return p; }
static void ws_release() { for (void* p : g_ws) cudaFree(p); g_ws.clear(); }
static cudaStream_t g_aux[2];
cudaStream_t aux_stream(int which) { if (!g_aux[which]) cudaStreamCreateWithFlags(&g_aux[which], 0); return g_aux[which]; }
static cudaEvent_t g_pool[64];
cudaEvent_t pooled_event(int idx) { if (idx >= 64) fatal("pooled_event", __FILE__, __LINE__, "pool exhausted"); if (!g_pool[idx]) cudaEventCreateWithFlags(&g_pool[idx], 0); return g_pool[idx]; }
void call_xerbla(const char*, int) {}
void finish_call() { sim::run_all(); }
void log_exec(const char*, const char*, ...) {}
}  // namespace b200

#include "../../libgpublas_b200/csrc/staged_gemm.cuh"
#include "../../libgpublas_b200/csrc/staged_level3.cuh"

// (the residency classifier of simcuda.inc knows cudaMalloc'd and registered buffers; managed blocks are registered as RES_MANAGED)
static unsigned long long g_data = 88172645463325252ull;
static void fill(std::vector<double>& v, double scale = 1.0) {
    for (auto& x : v) { g_data ^= g_data << 13; g_data ^= g_data >> 7; g_data ^= g_data << 17; x = scale * ((double)(g_data >> 11) * (2.0 / 9007199254740992.0) - 1.0); }
}
struct HostBuf {        // where: 1 pinned, 2 pageable, 3 tracked managed (fresh: not yet resident)
    double* p; size_t n; int where;
    HostBuf(const std::vector<double>& src, int w) : n(src.size()), where(w) {
        p = (double*)malloc(n * 8); memcpy(p, src.data(), n * 8);
        if (w == 1) sim::registered[(const char*)p] = {n * 8, (int)b200::RES_HOST_PINNED};
        if (w == 3) { sim::registered[(const char*)p] = {n * 8, (int)b200::RES_MANAGED}; g_managed[(const char*)p] = {n * 8, 0}; }
    }
    ~HostBuf() { sim::run_all(); sim::registered.erase((const char*)p); g_managed.erase((const char*)p); free(p); }
};
static double max_abs_diff(const double* a, const double* b, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; i++) { const double d = std::fabs(a[i] - b[i]); if (!(d <= m)) m = d; }
    return m;
}
static int g_fail = 0, g_cases = 0;
static void verdict(const char* what, int where, double err, double tol, bool taken) {
    g_cases++;
    const bool ok = taken && err <= tol && sim::deadlocks == 0;
    if (!ok) g_fail++;
    printf("%s %-34s operands=%s policy=%d err=%.3e tol=%.1e%s\n", ok ? "ok  " : "FAIL", what, where == 1 ? "pinned" : (where == 2 ? "pageable" : "managed"), sim::policy, err, tol,
           taken ? "" : " (PATH NOT TAKEN)");
    sim::deadlocks = 0;
    b200::ws_release();
}
static void dgemm_mock(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb, double beta, double* C,
                       int64_t ldc, int mask) { b200::dgemm_dev(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask); }

static void case_gemm(char ta, char tb, double beta, int where, bool first_touch) {
    const int m = 300, n = 2048 + 260, k = 2048 + 90;
    const int ra = ta == 'N' ? m : k, ca = ta == 'N' ? k : m, rb = tb == 'N' ? k : n, cb = tb == 'N' ? n : k;
    const int lda = ra + 2, ldb = rb + 4, ldc = m + 6;
    std::vector<double> A((size_t)lda * ca), B((size_t)ldb * cb), C((size_t)ldc * n);
    fill(A); fill(B); fill(C);
    std::vector<double> R = C;
    ob::dgemm(ob::ColMajor, ob::tr(ta), ob::tr(tb), m, n, k, 0.7, A.data(), lda, B.data(), ldb, beta, R.data(), ldc);
    HostBuf a(A, where), b(B, where), c(C, where);
    const bool taken = first_touch ? b200::gemm_first_touch<double>(dgemm_mock, ta, tb, m, n, k, 0.7, a.p, (int64_t)lda, b.p, (int64_t)ldb, beta, c.p, (int64_t)ldc)
                                   : b200::gemm_pipelined<double>(dgemm_mock, ta, tb, m, n, k, 0.7, a.p, (int64_t)lda, b.p, (int64_t)ldb, beta, c.p, (int64_t)ldc);
    sim::run_all();
    char name[80]; snprintf(name, sizeof name, "%s dgemm %c%c beta=%g", first_touch ? "first-touch" : "pipelined", ta, tb, beta);
    verdict(name, where, max_abs_diff(c.p, R.data(), R.size()), 1e-9, taken);      // (k is consumed in chunks: not bit-identical to one pass)
}
static void case_syrk(char uplo, char trans, double beta, int where) {
    const int n = 1500, k = 2300;
    const int ra = trans == 'N' ? n : k, ca = trans == 'N' ? k : n, lda = ra + 1, ldc = n + 3;
    std::vector<double> A((size_t)lda * ca), C((size_t)ldc * n);
    fill(A); fill(C);
    std::vector<double> R = C, F((size_t)n * n);
    ob::dgemm(ob::ColMajor, trans == 'N' ? ob::NoTrans : ob::Trans, trans == 'N' ? ob::Trans : ob::NoTrans, n, n, k, 0.7, A.data(), lda, A.data(), lda, 0.0, F.data(), n);
    for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
            if (uplo == 'L' ? i >= j : i <= j) R[i + (size_t)j * ldc] = F[i + (size_t)j * n] + beta * C[i + (size_t)j * ldc];
    HostBuf a(A, where), c(C, where);
    const bool taken = b200::syrk_pipelined<double>(uplo, trans, n, k, 0.7, a.p, (int64_t)lda, beta, c.p, (int64_t)ldc);
    sim::run_all();
    char name[80]; snprintf(name, sizeof name, "pipelined dsyrk %c%c beta=%g", uplo, trans, beta);
    verdict(name, where, max_abs_diff(c.p, R.data(), R.size()), 1e-9, taken);
}
static void case_trxm(bool solve, char side, char uplo, char trans, int where) {
    const int na = side == 'L' ? 300 : 260, nfree = side == 'L' ? 2300 : 2500;
    const int m = side == 'L' ? na : nfree, n = side == 'L' ? nfree : na, lda = na + 1, ldb = m + 2;
    std::vector<double> A((size_t)lda * na), B((size_t)ldb * n);
    fill(A, 1.0 / na); fill(B);
    for (int i = 0; i < na; i++) A[i + (size_t)i * lda] = 2.0;
    std::vector<double> R = B;
    (solve ? ob::dtrsm : ob::dtrmm)(ob::ColMajor, side == 'L' ? ob::Left : ob::Right, uplo == 'U' ? ob::Upper : ob::Lower, ob::tr(trans), ob::NonUnit, m, n, 0.7, A.data(), lda, R.data(), ldb);
    for (int j = 0; j < na; j++)            // the unreferenced triangle never crosses to the device: poison it on the host side
        for (int i = 0; i < na; i++)
            if (uplo == 'L' ? i < j : i > j) A[i + (size_t)j * lda] = NAN;
    HostBuf a(A, where), b(B, where);
    const bool taken = b200::trxm_pipelined<double>(solve, side, uplo, trans, 'N', m, n, 0.7, a.p, (int64_t)lda, b.p, (int64_t)ldb);
    sim::run_all();
    char name[80]; snprintf(name, sizeof name, "pipelined %s %c%c%cN %dx%d", solve ? "dtrsm" : "dtrmm", side, uplo, trans, m, n);
    verdict(name, where, max_abs_diff(b.p, R.data(), R.size()), 1e-9, taken);
}

int main() {
    if (!ob::load()) { printf("RESULT skipped (no OpenBLAS)\n"); return 0; }
    b200::g_opts.pipeline_min_bytes = 1000;
    sim::ndev = 1;
    for (int pol = 0; pol < 5; pol++) {
        sim::policy = pol >= 3 ? 0 : pol;
        sim::rng.seed(2000 + pol);
        const int w = pol & 1 ? 2 : 1;
        case_gemm('N', 'N', 1.3, w, false);
        case_gemm('T', 'N', 0.0, 3 - w, false);
        case_gemm('N', 'T', 1.3, w, false);
        case_gemm('N', 'N', 1.3, 3, true);
        case_gemm('T', 'T', 0.0, 3, true);
        case_syrk('L', 'N', 1.3, w);
        case_syrk('U', 'T', 0.0, 3 - w);
        case_syrk('U', 'N', 0.0, w);
        case_syrk('L', 'T', 1.3, 3 - w);
        case_trxm(true, 'L', 'L', 'N', w);
        case_trxm(true, 'R', 'U', 'T', 3 - w);
        case_trxm(false, 'L', 'U', 'T', w);
        case_trxm(false, 'R', 'L', 'N', 3 - w);
    }
    printf("RESULT cases=%d failed=%d operations=%ld\n", g_cases, g_fail, sim::executed);
    return g_fail != 0;
}
