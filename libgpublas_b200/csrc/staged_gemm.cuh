// staged_gemm.cuh -- GEMM on operands that live in ordinary host memory, with the host<->device
// traffic overlapped with the tensor-pipe work.
//
// The reference's only answer to an untracked operand is a whole-array blocking cudaMemcpy before the
// cuBLAS call and another after it (runtime-mem.hpp:84-112, :116-135); SURVEY.md section 8(f) rank 2
// names this path as the next row because real applications hit it for most calls (b2c_misses).
// At BASELINE config 2 (16384^3 f64) that serialisation costs as much as the GEMM itself
// (6.4 GB over PCIe vs 245 ms of DMMA work).
//
// Schedule (three streams, events only, no host synchronisation until the end):
//   * C is cut into column panels; A stays on the device once it has arrived.
//   * panel 0 is computed in k-chunks: chunk i of A (and of B's panel-0 slice) is copied on the H2D
//     stream while chunk i-1 is being multiplied (C_panel += A_chunk * B_chunk), so only the first
//     chunk's copy is exposed;
//   * panel p >= 1 needs only B's panel-p slice (and C's, when beta != 0): copied during panel p-1;
//   * the finished panel p-1 of C goes back on the D2H stream while panel p is computed (PCIe is
//     full duplex), so only the last panel's return is exposed.
// Operands that are already device-accessible (tracked managed blocks, device pointers) take part
// without copies, so any mix of resident and host operands uses the same code.
#pragma once
#include "abi_common.h"
#include "gemm_generic.cuh"
#include <algorithm>
#include <unistd.h>

namespace b200 {

struct StagedMat {
    const char* host; char* dev; int64_t ld, dld, rows, cols; size_t es; bool staged;
};

static inline StagedMat stage_matrix(const void* p, int64_t rows, int64_t cols, int64_t ld, size_t es, cudaStream_t s) {
    StagedMat m;
    m.host = (const char*)p; m.rows = rows; m.cols = cols; m.ld = ld; m.es = es;
    const Residency r = classify(p);
    m.staged = !(r == RES_DEVICE || r == RES_MANAGED);
    if (m.staged) {
        int64_t per16 = 16 / (int64_t)es; if (per16 < 1) per16 = 1;
        m.dld = (rows + per16 - 1) / per16 * per16;          // TMA-addressable pitch
        m.dev = (char*)ws_alloc((size_t)m.dld * cols * es);
        __atomic_fetch_add(&g_stats.misses, 1ull, __ATOMIC_RELAXED);
    } else {
        m.dld = ld; m.dev = (char*)p;
        __atomic_fetch_add(&g_stats.hits, 1ull, __ATOMIC_RELAXED);
        if (r == RES_MANAGED) make_resident(p, (size_t)((cols - 1) * ld + rows) * es, s);
    }
    return m;
}

// rows [r0, r0+nr) x columns [c0, c0+nc) between the host matrix and its device mirror
static inline void copy_region(const StagedMat& m, int64_t r0, int64_t c0, int64_t nr, int64_t nc, cudaStream_t st, bool to_device) {
    if (!m.staged || nr <= 0 || nc <= 0) return;
    TrackerGuard guard;
    const char* h = m.host + (size_t)(r0 + c0 * m.ld) * m.es;
    char* d = m.dev + (size_t)(r0 + c0 * m.dld) * m.es;
    if (to_device) {
        B200_CUDA(cudaMemcpy2DAsync(d, (size_t)m.dld * m.es, h, (size_t)m.ld * m.es, (size_t)nr * m.es, (size_t)nc, cudaMemcpyHostToDevice, st));
        __atomic_fetch_add(&g_stats.h2d_bytes, (unsigned long long)(nr * nc * m.es), __ATOMIC_RELAXED);
    } else {
        B200_CUDA(cudaMemcpy2DAsync((void*)h, (size_t)m.ld * m.es, d, (size_t)m.dld * m.es, (size_t)nr * m.es, (size_t)nc, cudaMemcpyDeviceToHost, st));
        __atomic_fetch_add(&g_stats.d2h_bytes, (unsigned long long)(nr * nc * m.es), __ATOMIC_RELAXED);
    }
    if (g_opts.trace_copy)
        b200_writef(STDOUT_FILENO, "b200blas: copy %zu B %s (%lld x %lld region)\n", (size_t)(nr * nc * m.es),
                    to_device ? "CPU ---> GPU" : "GPU ---> CPU", (long long)nr, (long long)nc);
}

static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Returns false when the call should take the plain path (small, or nothing to stage).
// GEMM: void(cudaStream_t, char, char, int, int, int, T, const T*, int64_t, const T*, int64_t, T, T*, int64_t, int)
template <typename T, typename GEMM>
bool gemm_pipelined(GEMM gemm, char ta, char tb, int m, int n, int k, T alpha, const T* a, int64_t lda, const T* b, int64_t ldb,
                    T beta, T* c, int64_t ldc) {
    const size_t es = sizeof(T);
    const bool nota = ta == 'N', notb = tb == 'N';
    const size_t total = ((size_t)m * k + (size_t)k * n + (size_t)m * n) * es;
    if (total < g_opts.pipeline_min_bytes || n < 1024 || k < 1024 || m < 128) return false;
    {
        const Residency ra = classify(a), rb = classify(b), rc = classify(c);
        auto host = [](Residency r) { return r == RES_HOST_PINNED || r == RES_HOST_PAGEABLE; };
        if (!host(ra) && !host(rb) && !host(rc)) return false;
    }
    cudaStream_t s = current_stream(), h2d = aux_stream(0), d2h = aux_stream(1);
    StagedMat A = stage_matrix(a, nota ? m : k, nota ? k : m, lda, es, s);
    StagedMat B = stage_matrix(b, notb ? k : n, notb ? n : k, ldb, es, s);
    StagedMat C = stage_matrix(c, m, n, ldc, es, s);
    const bool beta0 = is0(beta);
    const T one = num<T>::real(1.0);

    const int np = (int)std::max<int64_t>(512, round_up((n + 7) / 8, 128));      // <= 8 panels
    const int kc = (int)std::max<int64_t>(512, round_up((k + 7) / 8, 256));      // <= 8 chunks
    const int P = (n + np - 1) / np, NCH = (k + kc - 1) / kc;
    enum { EV_START = 0, EV_CHUNK = 1, EV_READY = 16, EV_DONE = 32, EV_END = 48 };

    {   // the copy streams must not overwrite workspace a previous (asynchronous) call may still be reading
        TrackerGuard guard;
        B200_CUDA(cudaEventRecord(pooled_event(EV_START), s));
        B200_CUDA(cudaStreamWaitEvent(h2d, pooled_event(EV_START), 0));
    }
    auto a_chunk = [&](int64_t k0) { return (const T*)(A.dev + (size_t)(nota ? k0 * A.dld : k0) * es); };
    auto b_block = [&](int64_t k0, int64_t n0) { return (const T*)(B.dev + (size_t)(notb ? k0 + n0 * B.dld : n0 + k0 * B.dld) * es); };
    auto c_panel = [&](int64_t n0) { return (T*)(C.dev + (size_t)(n0 * C.dld) * es); };
    auto copy_a = [&](int64_t k0, int64_t kk) { if (nota) copy_region(A, 0, k0, m, kk, h2d, true); else copy_region(A, k0, 0, kk, m, h2d, true); };
    auto copy_b = [&](int64_t k0, int64_t kk, int64_t n0, int64_t nn) {
        if (notb) copy_region(B, k0, n0, kk, nn, h2d, true); else copy_region(B, n0, k0, nn, kk, h2d, true);
    };
    auto record = [&](int ev, cudaStream_t on) { TrackerGuard guard; B200_CUDA(cudaEventRecord(pooled_event(ev), on)); };
    auto wait = [&](cudaStream_t who, int ev) { TrackerGuard guard; B200_CUDA(cudaStreamWaitEvent(who, pooled_event(ev), 0)); };

    // ---- panel 0: k-chunked so that the multiply starts after the first chunk has landed ----
    const int nn0 = std::min(np, n);
    for (int i = 0; i < NCH; i++) {
        const int64_t k0 = (int64_t)i * kc; const int kk = (int)std::min<int64_t>(kc, k - k0);
        if (i == 0 && !beta0) copy_region(C, 0, 0, m, nn0, h2d, true);
        copy_a(k0, kk);
        copy_b(k0, kk, 0, nn0);
        record(EV_CHUNK + i, h2d);
        wait(s, EV_CHUNK + i);
        gemm(s, ta, tb, m, nn0, kk, alpha, a_chunk(k0), A.dld, b_block(k0, 0), B.dld, i == 0 ? beta : one, c_panel(0), C.dld, MASK_FULL);
    }
    record(EV_DONE + 0, s);
    // ---- panels 1..P-1: B slice in, previous C panel out, both under the multiply ----
    for (int p = 1; p < P; p++) {
        const int64_t n0 = (int64_t)p * np; const int nn = (int)std::min<int64_t>(np, n - n0);
        if (!beta0) copy_region(C, 0, n0, m, nn, h2d, true);
        copy_b(0, k, n0, nn);
        record(EV_READY + p, h2d);
        wait(s, EV_READY + p);
        gemm(s, ta, tb, m, nn, k, alpha, a_chunk(0), A.dld, b_block(0, n0), B.dld, beta, c_panel(n0), C.dld, MASK_FULL);
        record(EV_DONE + p, s);
        wait(d2h, EV_DONE + p - 1);
        copy_region(C, 0, n0 - np, m, np, d2h, false);
    }
    wait(d2h, EV_DONE + P - 1);
    {
        const int64_t n0 = (int64_t)(P - 1) * np;
        copy_region(C, 0, n0, m, n - n0, d2h, false);
    }
    record(EV_END, d2h);
    wait(s, EV_END);      // the call's own stream completes only when C is back: finish_call() then covers everything
    return true;
}

}  // namespace b200
