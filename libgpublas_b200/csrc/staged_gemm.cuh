// staged_gemm.cuh -- GEMM on operands that live in ordinary host memory, with the host<->device
// traffic overlapped with the tensor-pipe work.
//
// The reference's only answer to an untracked operand is a whole-array blocking cudaMemcpy before the
// cuBLAS call and another after it (runtime-mem.hpp:84-112, :116-135); SURVEY.md section 8(f) rank 2
// names this path as the next row because real applications hit it for most calls (b2c_misses).
// At BASELINE config 2 (16384^3 f64) that serialisation costs as much as the GEMM itself
// (6.4 GB over PCIe vs 245 ms of DMMA work).
//
// Schedule (three streams, events only, no host synchronisation until the end; details at the boundaries below):
//   * the first ~3/4 of k in full-width k-chunks, chunk i+1's slices of A and B copied while chunk i is multiplied;
//   * the last ~1/4 of k column panel by column panel, each finished panel of C returning on the D2H stream under
//     the next panel's multiply (PCIe is full duplex).
// Operands that are already device-accessible (tracked managed blocks, device pointers) take part
// without copies, so any mix of resident and host operands uses the same code.
#pragma once
#include "abi_common.h"
#include "gemm_generic.cuh"
#include <algorithm>
#include <unistd.h>

namespace b200 {

struct StagedMat {
    const char* host; char* dev; int64_t ld, dld, rows, cols; size_t es; bool staged; bool pageable;
};

static inline StagedMat stage_matrix(const void* p, int64_t rows, int64_t cols, int64_t ld, size_t es, cudaStream_t s) {
    StagedMat m;
    m.host = (const char*)p; m.rows = rows; m.cols = cols; m.ld = ld; m.es = es;
    const Residency r = classify(p);
    m.staged = !(r == RES_DEVICE || r == RES_MANAGED);
    m.pageable = r == RES_HOST_PAGEABLE;
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
    // pageable memory goes through the pinned bounce ring packed by the host-thread pool (host_stager.cu): the driver's own
    // staging of a pageable cudaMemcpy2DAsync is one thread at ~11 GB/s
    const bool bounce = m.pageable && staged_copy_worthwhile((size_t)nr * nc * m.es);
    if (to_device) {
        if (bounce) staged_copy_2d(d, (size_t)m.dld * m.es, h, (size_t)m.ld * m.es, (size_t)nr * m.es, (size_t)nc, true, st);
        else B200_CUDA(cudaMemcpy2DAsync(d, (size_t)m.dld * m.es, h, (size_t)m.ld * m.es, (size_t)nr * m.es, (size_t)nc, cudaMemcpyHostToDevice, st));
        __atomic_fetch_add(&g_stats.h2d_bytes, (unsigned long long)(nr * nc * m.es), __ATOMIC_RELAXED);
    } else {
        if (bounce) staged_copy_2d((void*)h, (size_t)m.ld * m.es, d, (size_t)m.dld * m.es, (size_t)nr * m.es, (size_t)nc, false, st);
        else B200_CUDA(cudaMemcpy2DAsync((void*)h, (size_t)m.ld * m.es, d, (size_t)m.dld * m.es, (size_t)nr * m.es, (size_t)nc, cudaMemcpyDeviceToHost, st));
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

    // Schedule.  The first ~3/4 of k is consumed in full-width k-chunks: C += A[:, chunk] * B[chunk, :], each chunk's
    // operand slices (m*kc + kc*n elements) copied on the H2D stream while the previous chunk is multiplied -- per chunk
    // the copy is ~1/3 of the multiply, so after the first chunk (split 1/4 + 1/4 + 1/2 to shrink the exposed head) the
    // tensor pipe never waits for PCIe.  The last ~1/4 of k is consumed column panel by column panel, each finished panel
    // of C returning on the D2H stream under the next panel's multiply (last panel split 1/2 + 1/4 + 1/4 to shrink the
    // exposed tail).  (An earlier schedule -- column panels outermost, A resident after panel 0 -- made panel 0 wait for
    // all of A: 44 ms of copies against 31 ms of work.)
    const int np = (int)std::max<int64_t>(512, round_up((n + 7) / 8, 128));
    const int kc = (int)std::max<int64_t>(512, round_up((k + 7) / 8, 256));
    const int64_t kL = k >= 4 * kc ? (int64_t)(3 * (int64_t)k / 4) / kc * kc : 0;      // [0, kL) by chunks, [kL, k) by panels
    int pb[24], kb[24], P = 0, NCH = 0;          // boundaries: panel p = [pb[p], pb[p+1]), chunk i = [kb[i], kb[i+1])
    pb[0] = 0;
    for (int64_t c = 0; c < n; c += np) {
        const int64_t w = std::min<int64_t>(np, n - c);
        if (c + np >= n && w >= 512) {           // last panel
            const int64_t h = round_up(w / 2, 128), q = round_up(w / 4, 128);
            pb[++P] = (int)(c + h); if (c + h + q < c + w) pb[++P] = (int)(c + h + q); pb[++P] = (int)(c + w);
        } else pb[++P] = (int)(c + w);
    }
    kb[0] = 0;
    for (int64_t c = 0; c < kL; c += kc) {
        const int64_t d = std::min<int64_t>(kc, kL - c);
        if (c == 0 && d >= 1024) {               // first chunk
            const int64_t q = round_up(d / 4, 128);
            kb[++NCH] = (int)q; kb[++NCH] = (int)(2 * q); kb[++NCH] = (int)d;
        } else kb[++NCH] = (int)(c + d);
    }
    enum { EV_START = 0, EV_CHUNK = 1, EV_TAIL = 15, EV_DONE = 32, EV_END = 60 };

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

    // ---- H2D stream: every copy is queued up front (asynchronous for pinned memory; for pageable memory the calls
    // block the host, which is why they are interleaved with the launches below instead) ----
    if (!beta0) copy_region(C, 0, 0, m, n, h2d, true);
    // ---- full-width k-chunks ----
    for (int i = 0; i < NCH; i++) {
        const int64_t k0 = kb[i]; const int kk = kb[i + 1] - kb[i];
        copy_a(k0, kk);
        copy_b(k0, kk, 0, n);
        record(EV_CHUNK + i, h2d);
        wait(s, EV_CHUNK + i);
        gemm(s, ta, tb, m, n, kk, alpha, a_chunk(k0), A.dld, b_block(k0, 0), B.dld, i == 0 ? beta : one, (T*)C.dev, C.dld, MASK_FULL);
    }
    // ---- the rest of k, panel by panel, with C flowing back ----
    copy_a(kL, k - kL);
    copy_b(kL, k - kL, 0, n);
    record(EV_TAIL, h2d);
    wait(s, EV_TAIL);
    // (panel p's copy back is issued AFTER panel p+1's multiply has been launched: with pageable memory the copy blocks the host
    // until the panel is in host memory, and the next multiply must already be running underneath it)
    for (int p = 0; p <= P; p++) {
        if (p < P) {
            const int64_t n0 = pb[p]; const int nn = pb[p + 1] - pb[p];
            gemm(s, ta, tb, m, nn, (int)(k - kL), alpha, a_chunk(kL), A.dld, b_block(kL, n0), B.dld, NCH ? one : beta, c_panel(n0), C.dld, MASK_FULL);
            record(EV_DONE + p, s);
        }
        if (p > 0) {
            const int64_t n0 = pb[p - 1]; const int nn = pb[p] - pb[p - 1];
            wait(d2h, EV_DONE + p - 1);
            copy_region(C, 0, n0, m, nn, d2h, false);
        }
    }
    record(EV_END, d2h);
    wait(s, EV_END);      // the call's own stream completes only when C is back: finish_call() then covers everything
    return true;
}

// ---------------------------------------------------------------------------------------------
// First use of TRACKED MANAGED operands the CPU has just filled -- the reference's hit path (an unmodified program callocs its
// matrices under the interposer and calls dgemm_): the pages live in host memory.  make_resident() migrates whole blocks before the
// kernel starts: 6.4 GB at ~43 GB/s = 150 ms in front of 245 ms of DMMA for 16384^3 (profiles/r02a_bench_n1.json:
// e2e.managed_first_touch = 22 TFLOP/s).  Here the migration is issued as cudaMemPrefetchAsync of CONTIGUOUS ranges on a second
// stream in the order the multiply needs them, and the multiply follows range by range, in place (no staging buffers):
//   column panel 0 of B and C, then A in k-chunks (panel 0 is multiplied chunk by chunk as A arrives), then the panels 1.. of
//   B and C under the multiplies of the panels before them.
// Ranges are contiguous only for 'N' operands (columns of a column-major array); a transposed operand is migrated whole, first.
// Operands that are already resident (or not tracked: the application manages those) take part without prefetches.
template <typename T, typename GEMM>
bool gemm_first_touch(GEMM gemm, char ta, char tb, int m, int n, int k, T alpha, const T* a, int64_t lda, const T* b, int64_t ldb,
                      T beta, T* c, int64_t ldc) {
    const size_t es = sizeof(T);
    if (g_opts.prefetch != 1 || n < 2048 || k < 2048 || m < 128) return false;
    const size_t total = ((size_t)m * k + (size_t)k * n + (size_t)m * n) * es;
    if (total < g_opts.pipeline_min_bytes) return false;
    if (classify(a) != RES_MANAGED || classify(b) != RES_MANAGED || classify(c) != RES_MANAGED) return false;
    const bool fresh_a = tracker_peek_resident(a) == 0, fresh_b = tracker_peek_resident(b) == 0, fresh_c = tracker_peek_resident(c) == 0;
    if (!fresh_a && !fresh_b) return false;                       // nothing large to migrate: the plain path
    const bool nota = ta == 'N', notb = tb == 'N';
    cudaStream_t s = current_stream(), pf = aux_stream(0);
    const int dev = current_device();
    const T one = num<T>::real(1.0);
    auto record = [&](int ev, cudaStream_t on) { TrackerGuard guard; B200_CUDA(cudaEventRecord(pooled_event(ev), on)); };
    auto wait = [&](cudaStream_t who, int ev) { TrackerGuard guard; B200_CUDA(cudaStreamWaitEvent(who, pooled_event(ev), 0)); };
    auto prefetch = [&](const void* p, size_t bytes) {
        TrackerGuard guard;
        if (cudaMemPrefetchAsync(p, bytes, dev, pf) == cudaSuccess) __atomic_fetch_add(&g_stats.prefetch_bytes, (unsigned long long)bytes, __ATOMIC_RELAXED);
        else cudaGetLastError();
    };
    auto advise = [&](const void* p) {        // whole block: preferred location = this device, marked as migrated
        TrackerGuard guard;
        void* base = nullptr; size_t bsize = 0;
        if (tracker_lookup(p, &base, &bsize)) cudaMemAdvise(base, bsize, cudaMemAdviseSetPreferredLocation, dev);
        cudaGetLastError();
        tracker_test_and_set_resident(p);
    };
    enum { EV_START = 0, EV_CHUNK = 1, EV_PANEL = 20, EV_END = 60 };
    record(EV_START, s);
    wait(pf, EV_START);
    if (fresh_a) advise(a);
    if (fresh_b) advise(b);
    if (fresh_c) advise(c);
    __atomic_fetch_add(&g_stats.hits, 3ull, __ATOMIC_RELAXED);
    const int P = 8;
    const int64_t np = round_up((n + P - 1) / P, 128);
    const int NCH = 8;
    const int64_t kc = round_up((k + NCH - 1) / NCH, 256);
    const size_t a_bytes = (size_t)(((nota ? k : m) - 1) * lda + (nota ? m : k)) * es;
    const size_t b_bytes = (size_t)(((notb ? n : k) - 1) * ldb + (notb ? k : n)) * es;
    auto b_panel = [&](int64_t n0, int64_t nn) { if (fresh_b && notb) prefetch(b + n0 * ldb, (size_t)((nn - 1) * ldb + k) * es); };
    auto c_panel = [&](int64_t n0, int64_t nn) { if (fresh_c) prefetch(c + n0 * ldc, (size_t)((nn - 1) * ldc + m) * es); };
    if (fresh_b && !notb) prefetch(b, b_bytes);
    if (fresh_a && !nota) prefetch(a, a_bytes);
    // ---- panel 0: B and C first, then A chunk by chunk with the multiply behind it ----
    const int64_t n00 = std::min<int64_t>(np, n);
    b_panel(0, n00);
    c_panel(0, n00);
    int nev = 0;
    for (int64_t k0 = 0; k0 < k; k0 += kc, nev++) {
        const int64_t kk = std::min<int64_t>(kc, k - k0);
        if (fresh_a && nota) prefetch(a + k0 * lda, (size_t)((kk - 1) * lda + m) * es);
        record(EV_CHUNK + nev, pf);
        wait(s, EV_CHUNK + nev);
        gemm(s, ta, tb, m, (int)n00, (int)kk, alpha, nota ? a + k0 * lda : a + k0, lda, notb ? b + k0 : b + k0 * ldb, ldb, k0 == 0 ? beta : one, c, ldc, MASK_FULL);
    }
    // ---- the other panels: their B and C columns travel under the multiplies before them ----
    int pe = 0;
    for (int64_t n0 = n00; n0 < n; n0 += np, pe++) {
        const int64_t nn = std::min<int64_t>(np, n - n0);
        b_panel(n0, nn);
        c_panel(n0, nn);
        record(EV_PANEL + pe, pf);
        wait(s, EV_PANEL + pe);
        gemm(s, ta, tb, m, (int)nn, k, alpha, a, lda, notb ? b + n0 * ldb : b + n0, ldb, beta, c + n0 * ldc, ldc, MASK_FULL);
    }
    return true;
}

}  // namespace b200
