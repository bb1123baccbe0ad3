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

#include "simcuda.inc"


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
    double* p = nullptr; size_t n = 0; int where = 0;      // 0 device, 1 pinned host, 2 pageable host, 3 tracked managed
    Buf(const std::vector<double>& src, int where_) : n(src.size()), where(where_) {
        if (where == 0) { cudaMalloc((void**)&p, n * 8); }
        else {
            p = (double*)malloc(n * 8);
            if (where == 1) sim::registered[(const char*)p] = {n * 8, (int)b200::RES_HOST_PINNED};
            if (where == 3) sim::registered[(const char*)p] = {n * 8, (int)b200::RES_MANAGED};
        }
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
    printf("%s %-28s devices=%d operands=%s policy=%d err=%.3e tol=%.1e%s\n", ok ? "ok  " : "FAIL", what, ndev, where == 0 ? "device" : (where == 1 ? "pinned" : (where == 2 ? "pageable" : "managed")), sim::policy, err,
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
            if (pol == 0) { case_gemm(nd, 'N', 'T', 1); case_gemm(nd, 'T', 'T', 3); }
            case_syrk(nd, 'L', 'N', 1.3, 0);
            case_syrk(nd, 'U', 'T', 0.0, pol & 1 ? 2 : 1);
            if (pol == 0) { case_syrk(nd, 'U', 'N', 1.3, 0); case_syrk(nd, 'L', 'T', 0.0, 1); case_syrk(nd, 'L', 'N', 0.0, 3); }
            case_trxm(nd, true, 'L', 'L', 'N', 0);
            case_trxm(nd, true, 'R', 'U', 'T', pol & 1 ? 1 : 2);
            case_trxm(nd, false, 'L', 'U', 'N', pol & 1 ? 0 : 1);
            if (pol == 0) { case_trxm(nd, false, 'R', 'L', 'T', 0); case_trxm(nd, true, 'L', 'U', 'T', 3); }
            case_cholesky(nd, 128);
            if (pol == 0) case_cholesky(nd, 256);
        }
    }
    printf("RESULT cases=%d failed=%d operations=%ld\n", g_cases, g_fail, sim::executed);
    return g_fail != 0;
}
