// staged_level3.cuh -- ?syrk_, ?trsm_ and ?trmm_ on operands in ordinary host memory, with the host<->device traffic
// overlapped with the compute (the same three-stream scheme as staged_gemm.cuh).
//
// The reference's miss path stages whole arrays with blocking copies around the cuBLAS call (runtime-mem.hpp:84-165), and until
// now so did these three routines (Operand): for DSYRK n = k = 16384 from host memory that is 2 GB in, 2 GB of C in, 2 GB of C
// out in series with 125 ms of DMMA work.
//
//  SYRK   A is staged once (it is both operands).  The first ~3/4 of k goes in k-chunks, chunk i+1's columns of A copied while
//         chunk i is multiplied (one masked GEMM per chunk over the whole triangle).  The last ~1/4 of k goes trapezoid by
//         trapezoid -- lower: rows [p0, n) of columns [p0, p1); upper: rows [p0, p1) of columns [p0, n) -- each finished
//         trapezoid returning on the D2H stream under the next one's multiply.  Inbound C: with beta == 0 only the diagonal
//         blocks of the trapezoids (their unreferenced halves must survive the round trip), else the trapezoids themselves;
//         the strictly unreferenced rectangles never cross PCIe in either direction (the old path moved the full square both ways).
//  TRSM / TRMM  right-hand sides are independent: B goes through in blocks of columns (side 'L') or rows (side 'R'); block
//         p+1 travels in and block p-1 travels out while block p is solved / multiplied.  The triangle A is staged once, up front.
#pragma once
#include "staged_gemm.cuh"

namespace b200 {

template <typename T>
bool syrk_pipelined(char uplo, char trans, int n, int k, T alpha, const T* a, int64_t lda, T beta, T* c, int64_t ldc) {
    const size_t es = sizeof(T);
    const bool nota = trans == 'N', upper = uplo == 'U';
    const size_t total = ((size_t)n * k + (size_t)n * n) * es;
    if (total < g_opts.pipeline_min_bytes || n < 1024 || k < 1024) return false;
    {
        const Residency ra = classify(a), rc = classify(c);
        auto host = [](Residency r) { return r == RES_HOST_PINNED || r == RES_HOST_PAGEABLE; };
        if (!host(ra) && !host(rc)) return false;
    }
    cudaStream_t s = current_stream(), h2d = aux_stream(0), d2h = aux_stream(1);
    StagedMat A = stage_matrix(a, nota ? n : k, nota ? k : n, lda, es, s);
    StagedMat C = stage_matrix(c, n, n, ldc, es, s);
    const bool beta0 = is0(beta);
    const T one = num<T>::real(1.0);
    const char ta = nota ? 'N' : 'T', tb = nota ? 'T' : 'N';
    const int mask = upper ? MASK_UPPER : MASK_LOWER;

    // trapezoids of roughly equal area (boundaries n(1 - sqrt(1 - p/P)) on multiples of 128), the last one split once more: the
    // exposed tail is the D2H of the last trapezoid
    int pb[24], P = 0;
    pb[0] = 0;
    {
        const int parts = 8;
        for (int p = 1; p < parts; p++) {
            int64_t v = (int64_t)((double)n * (1.0 - std::sqrt(1.0 - (double)p / parts)) / 128.0 + 0.5) * 128;
            if (v > pb[P] && v < n) pb[++P] = (int)v;
        }
        const int64_t last = n - pb[P];
        if (last >= 1024) { pb[P + 1] = (int)(pb[P] + round_up(last / 3, 128)); P++; }
        pb[++P] = n;
    }
    const int kc = (int)std::max<int64_t>(512, round_up((k + 7) / 8, 256));
    const int64_t kL = k >= 4 * kc ? (int64_t)(3 * (int64_t)k / 4) / kc * kc : 0;      // [0, kL) by chunks, [kL, k) by trapezoids
    int kb[24], NCH = 0;
    kb[0] = 0;
    for (int64_t c0 = 0; c0 < kL; c0 += kc) {
        const int64_t d = std::min<int64_t>(kc, kL - c0);
        if (c0 == 0 && d >= 1024) { const int64_t q = round_up(d / 4, 128); kb[++NCH] = (int)q; kb[++NCH] = (int)(2 * q); kb[++NCH] = (int)d; }
        else kb[++NCH] = (int)(c0 + d);
    }
    enum { EV_START = 0, EV_CHUNK = 1, EV_TAIL = 15, EV_DONE = 32, EV_END = 60 };
    auto record = [&](int ev, cudaStream_t on) { TrackerGuard guard; B200_CUDA(cudaEventRecord(pooled_event(ev), on)); };
    auto wait = [&](cudaStream_t who, int ev) { TrackerGuard guard; B200_CUDA(cudaStreamWaitEvent(who, pooled_event(ev), 0)); };
    record(EV_START, s);
    wait(h2d, EV_START);
    auto a_rows = [&](int64_t r0, int64_t k0) { return (const T*)(A.dev + (size_t)(nota ? r0 + k0 * A.dld : k0 + r0 * A.dld) * es); };   // op(A)[r0.., k0..]
    auto c_at = [&](int64_t p0) { return (T*)(C.dev + (size_t)(p0 + p0 * C.dld) * es); };
    auto copy_a = [&](int64_t k0, int64_t kk) { if (nota) copy_region(A, 0, k0, n, kk, h2d, true); else copy_region(A, k0, 0, kk, n, h2d, true); };
    // trapezoid p of C: its rectangle, or (diag_only) just the diagonal block
    auto copy_c = [&](int p, bool diag_only, cudaStream_t st, bool to_device) {
        const int64_t p0 = pb[p], w = pb[p + 1] - pb[p];
        const int64_t nr = diag_only ? w : (upper ? w : n - p0), nc = diag_only ? w : (upper ? n - p0 : w);
        copy_region(C, p0, p0, nr, nc, st, to_device);
    };
    for (int p = 0; p < P; p++) copy_c(p, beta0, h2d, true);
    for (int i = 0; i < NCH; i++) {
        const int64_t k0 = kb[i]; const int kk = kb[i + 1] - kb[i];
        copy_a(k0, kk);
        record(EV_CHUNK + i, h2d);
        wait(s, EV_CHUNK + i);
        gemm_dev<T>(s, ta, tb, n, n, kk, alpha, a_rows(0, k0), A.dld, a_rows(0, k0), A.dld, i == 0 ? beta : one, (T*)C.dev, C.dld, mask);
    }
    copy_a(kL, k - kL);
    record(EV_TAIL, h2d);
    wait(s, EV_TAIL);
    for (int p = 0; p <= P; p++) {
        if (p < P) {
            const int64_t p0 = pb[p], w = pb[p + 1] - pb[p];
            const int tm = (int)(upper ? w : n - p0), tn = (int)(upper ? n - p0 : w);
            gemm_dev<T>(s, ta, tb, tm, tn, (int)(k - kL), alpha, a_rows(p0, kL), A.dld, a_rows(p0, kL), A.dld, NCH ? one : beta, c_at(p0), C.dld, mask);
            record(EV_DONE + p, s);
        }
        if (p > 0) { wait(d2h, EV_DONE + p - 1); copy_c(p - 1, false, d2h, false); }
    }
    record(EV_END, d2h);
    wait(s, EV_END);
    return true;
}

// solve: ?trsm_, else ?trmm_ (same data flow)
template <typename T>
bool trxm_pipelined(bool solve, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* a, int64_t lda, T* b, int64_t ldb) {
    const size_t es = sizeof(T);
    const bool lside = side == 'L';
    const int64_t na = lside ? m : n, nfree = lside ? n : m;
    if ((size_t)m * n * es < g_opts.pipeline_min_bytes || nfree < 2048 || na < 256 || is0(alpha)) return false;
    {
        const Residency rb = classify(b);
        if (!(rb == RES_HOST_PINNED || rb == RES_HOST_PAGEABLE)) return false;
    }
    cudaStream_t s = current_stream(), h2d = aux_stream(0), d2h = aux_stream(1);
    StagedMat A = stage_matrix(a, na, na, lda, es, s);
    StagedMat B = stage_matrix(b, m, n, ldb, es, s);
    // blocks: a short first one (the exposed head is its H2D), equal ones after, short last ones (the exposed tail is the last D2H)
    int bb[24], P = 0;
    {
        // few, wide blocks: the substitution leaves of trsm_dev run one right-hand side per thread, so a block's leaf chain costs
        // the same whatever its width (8 blocks of 1024 columns: 68 ms for DTRSM 8192^2 against 20.5 ms resident, profiles/r02o_bench_n1.json)
        const int64_t body = std::max<int64_t>(1024, round_up((nfree + 3) / 4, 128));
        bb[0] = 0;
        int64_t c0 = std::min<int64_t>(nfree, round_up(body / 2, 128));
        bb[++P] = (int)c0;
        while (c0 < nfree) {
            int64_t w = std::min<int64_t>(body, nfree - c0);
            if (nfree - c0 - w < body / 2 && nfree - c0 > body / 2 + 128) w = round_up((nfree - c0) / 2, 128);     // split the remainder in two short blocks
            c0 += w; bb[++P] = (int)c0;
        }
    }
    enum { EV_START = 0, EV_IN = 1, EV_DONE = 24, EV_END = 60 };       // at most ~12 blocks (pooled_event holds 64)
    auto record = [&](int ev, cudaStream_t on) { TrackerGuard guard; B200_CUDA(cudaEventRecord(pooled_event(ev), on)); };
    auto wait = [&](cudaStream_t who, int ev) { TrackerGuard guard; B200_CUDA(cudaStreamWaitEvent(who, pooled_event(ev), 0)); };
    record(EV_START, s);
    wait(h2d, EV_START);
    auto copy_b = [&](int p, cudaStream_t st, bool to_device) {
        if (lside) copy_region(B, 0, bb[p], m, bb[p + 1] - bb[p], st, to_device); else copy_region(B, bb[p], 0, bb[p + 1] - bb[p], n, st, to_device);
    };
    {   // only the referenced trapezoids of A cross PCIe (column groups: rows [c, na) of a lower, [0, c + w) of an upper triangle)
        const int64_t cg = std::max<int64_t>(256, round_up((na + 7) / 8, 128));
        for (int64_t c0 = 0; c0 < na; c0 += cg) {
            const int64_t w = std::min(cg, na - c0);
            if (uplo == 'U') copy_region(A, 0, c0, c0 + w, w, h2d, true); else copy_region(A, c0, c0, na - c0, w, h2d, true);
        }
    }
    copy_b(0, h2d, true);
    record(EV_IN + 0, h2d);
    for (int p = 0; p <= P; p++) {
        if (p + 1 < P) { copy_b(p + 1, h2d, true); record(EV_IN + p + 1, h2d); }
        if (p < P) {
            wait(s, EV_IN + p);
            T* blk = (T*)(B.dev + (size_t)(lside ? (int64_t)bb[p] * B.dld : bb[p]) * es);
            const int bm = lside ? m : bb[p + 1] - bb[p], bn = lside ? bb[p + 1] - bb[p] : n;
            if (solve) trsm_dev<T>(s, side, uplo, trans, diag, bm, bn, alpha, (const T*)A.dev, A.dld, blk, B.dld);
            else trmm_dev<T>(s, side, uplo, trans, diag, bm, bn, alpha, (const T*)A.dev, A.dld, blk, B.dld);
            record(EV_DONE + p, s);
        }
        if (p > 0) { wait(d2h, EV_DONE + p - 1); copy_b(p - 1, d2h, false); }
    }
    record(EV_END, d2h);
    wait(s, EV_END);
    return true;
}

}  // namespace b200
