// multi_level3.cu -- ?syrk_, ?trsm_ and ?trmm_ partitioned over the GPUs of one box, inside the interposed symbol
// (north_star (4): "for partitioned GEMM/SYRK/TRSM, at 2, 4 and 8 GPUs").  The reference forwards
// these calls to one device (blas_level3/syrk.cc:43-76, trsm.cc:40-73, trmm.cc:42-79); with `devices=<n>` an unmodified program
// gets n devices from the same call.  State, streams and the copy-engine chains are those of multi_gemm.cu (multi_state.h).
//
// SYRK  C := alpha*op(A)*op(A)^T + beta*C on one triangle.  The triangle is cut into N strips of EQUAL REFERENCED AREA
//       (boundaries b_s = n(1 - sqrt(1 - s/N)), on multiples of the 128-wide CTA tile): lower -> column strips, device s computes the trapezoid
//       C[b_s:n, b_s:b_s+1) with ONE masked GEMM launch (the SYRK of its diagonal block and the GEMM below it are the same launch:
//       tiles above the diagonal exit, level3_blocked.cu: syrk_dev); upper -> row strips C[b_s:b_s+1, b_s:n).  Either way device s
//       needs rows [b_s, n) of op(A) and nothing else.  A row piece that lies in strip j is therefore consumed by devices 0..j: it
//       leaves its origin (home HBM, or host memory over the first receiver's own PCIe link) ONCE and is forwarded along a chain
//       through those devices, first receiver rotating with the piece index; pieces are queued bottom strip first, so the device
//       with the shortest panel starts first.  k is never split: every C element is produced by one device.
// TRSM / TRMM  the right-hand sides are independent: side 'L' -> column blocks of B, side 'R' -> row blocks.  The referenced
//       triangle of A travels to every device in column groups (trapezoids: only the referenced rows) along rotating chains; each
//       device pulls its block of B with its own copy engine, runs the single-GPU recursion (trsm_dev / trmm_dev) on it and sends
//       it back.  The home GPU works on its block in place.
// Both are "bulk": a device's kernel starts when its operands have landed (events), not tile by tile as the flag-polling DGEMM.
#include "abi_common.h"
#include "kernels.h"
#include "multi_gemm.h"
#include "multi_state.h"
#include <algorithm>
#include <cmath>
#include <vector>

namespace b200 {

// ---------------------------------------------------------------------------------------------
// pure host logic (exported for the CPU tests)
void ml3_strips(int64_t n, int ndev, int64_t* b) {
    b[0] = 0;
    for (int s = 1; s < ndev; s++) {
        const double x = (double)n * (1.0 - std::sqrt(1.0 - (double)s / ndev));
        int64_t v = (int64_t)(x / 128.0 + 0.5) * 128;              // the CTA tile
        v = std::max(v, b[s - 1] + 128);
        b[s] = std::min(v, n);
    }
    b[ndev] = n;
}
int64_t ml3_piece_rows(int64_t n) { return std::max<int64_t>(256, ((n / 32 + 127) / 128) * 128); }

// Chains of one piece through `members` (slots that consume it), first receiver rotating with `rot`.  Device-resident operands:
// the home GPU (slot 0) is the origin and never a receiver.
static void ml3_chain(std::vector<MgHop>& plan, std::vector<int> members, bool host_source, int rot, int kind, int gidx, int piece, int64_t off, int64_t len) {
    if (!host_source) members.erase(std::remove(members.begin(), members.end(), 0), members.end());
    const int cnt = (int)members.size();
    int prev = -1;
    for (int i = 0; i < cnt; i++) {
        const int s = members[(rot + i) % cnt];
        plan.push_back(MgHop{kind, gidx, piece, off, len, prev, s});
        prev = s;
    }
}
// SYRK: row pieces of op(A); gidx = the strip the piece lies in (consumers: slots 0..gidx), off = global row
std::vector<MgHop> ml3_syrk_plan(int ndev, int64_t n, bool host_source) {
    std::vector<int64_t> b(ndev + 1);
    ml3_strips(n, ndev, b.data());
    const int64_t pg = ml3_piece_rows(n);
    std::vector<MgHop> plan;
    int piece = 0;
    for (int j = ndev - 1; j >= 0; j--)
        for (int64_t off = b[j]; off < b[j + 1]; off += pg, piece++) {
            std::vector<int> members;
            for (int s = 0; s <= j; s++) members.push_back(s);
            ml3_chain(plan, members, host_source, piece, 0, j, piece, off, std::min(pg, b[j + 1] - off));
        }
    return plan;
}
// TRSM/TRMM: column groups of the triangle A (na x na); every slot consumes every group; off = first column
std::vector<MgHop> ml3_tri_plan(int ndev, int64_t na, bool host_source) {
    const int64_t cg = std::max<int64_t>(256, ((na / 16 + 127) / 128) * 128);
    std::vector<MgHop> plan;
    std::vector<int> members;
    for (int s = 0; s < ndev; s++) members.push_back(s);
    int piece = 0;
    for (int64_t off = 0; off < na; off += cg, piece++) ml3_chain(plan, members, host_source, piece, 0, 0, piece, off, std::min(cg, na - off));
    return plan;
}

namespace {

inline bool on_host(Residency r) { return r == RES_HOST_PINNED || r == RES_HOST_PAGEABLE; }
template <typename T> inline int64_t even_rows(int64_t v) { const int64_t per16 = std::max<int64_t>(1, 16 / (int64_t)sizeof(T)); return (v + per16 - 1) / per16 * per16; }

// Common prologue / epilogue of a bulk partitioned call: every stream of every device waits for the caller's stream; at the end
// the caller's stream waits for every device.
struct BulkCall {
    MgState& st; int ndev; cudaStream_t home_stream;
    std::vector<std::vector<cudaEvent_t>> landed;       // per slot: events its compute stream must wait for
    BulkCall(MgState& s, int n) : st(s), ndev(n), home_stream(current_stream()), landed(n) {
        cudaEvent_t start;
        { MgDev& h = st.dev[0]; h.next_event = 0; start = next_event(h); B200_CUDA(cudaEventRecord(start, home_stream)); }
        for (int i = 0; i < ndev; i++) {
            MgDev& d = st.dev[i];
            if (i) d.next_event = 0;
            DeviceScope scope(d.id);
            if (i) ws_reset();           // (the home context was reset by the entry point)
            B200_CUDA(cudaStreamWaitEvent(d.comp, start, 0));
            B200_CUDA(cudaStreamWaitEvent(d.in, start, 0));
            B200_CUDA(cudaStreamWaitEvent(d.out, start, 0));
            B200_CUDA(cudaStreamWaitEvent(push_stream(d), start, 0));
        }
    }
    cudaStream_t comp(int s) const { return s == 0 ? home_stream : st.dev[s].comp; }
    // one hop of a chain: a 2-D copy on the stream of whoever holds the piece (origin in host memory: the receiver's own H2D stream)
    cudaEvent_t hop(const MgHop& hp, bool host_source, char* dstp, size_t dpitch, const char* src, size_t spitch, size_t width, size_t height,
                    cudaEvent_t after, unsigned long long* origin_bytes, unsigned long long* forward_bytes) {
        const bool h2d = hp.src < 0 && host_source;
        const int exec_slot = h2d ? hp.dst : (hp.src < 0 ? 0 : hp.src);
        MgDev& ex = st.dev[exec_slot];
        cudaStream_t stream = h2d ? ex.in : push_stream(ex);
        DeviceScope scope(ex.id);
        if (after) B200_CUDA(cudaStreamWaitEvent(stream, after, 0));
        B200_CUDA(cudaMemcpy2DAsync(dstp, dpitch, src, spitch, width, height, cudaMemcpyDefault, stream));
        cudaEvent_t ev = next_event(ex);
        B200_CUDA(cudaEventRecord(ev, stream));
        landed[hp.dst].push_back(ev);
        if (hp.src < 0) *origin_bytes += (unsigned long long)width * height; else *forward_bytes += (unsigned long long)width * height;
        if (h2d) __atomic_fetch_add(&g_stats.h2d_bytes, (unsigned long long)(width * height), __ATOMIC_RELAXED);
        return ev;
    }
    void wait_landed(int s) {
        DeviceScope scope(st.dev[s].id);
        for (cudaEvent_t e : landed[s]) B200_CUDA(cudaStreamWaitEvent(comp(s), e, 0));
    }
    void finish() {
        for (int s = 0; s < ndev; s++) {
            MgDev& d = st.dev[s];
            DeviceScope scope(d.id);
            cudaStream_t stream = comp(s);
            // the copy streams must drain before the call is over (the next call reuses the panels)
            for (cudaStream_t other : {push_stream(d), d.in, d.out}) { cudaEvent_t e = next_event(d); B200_CUDA(cudaEventRecord(e, other)); B200_CUDA(cudaStreamWaitEvent(stream, e, 0)); }
            if (s) B200_CUDA(cudaEventRecord(d.done, stream));
        }
        for (int s = 1; s < ndev; s++) B200_CUDA(cudaStreamWaitEvent(home_stream, st.dev[s].done, 0));
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
template <typename T>
bool multi_syrk(char uplo, char trans, int n, int k, T alpha, const T* a, int64_t lda, T beta, T* c, int64_t ldc) {
    const int ndev = g_opts.devices;
    if (ndev < 2) return false;
    if ((int64_t)n < (int64_t)1024 * ndev || k < 256 ||
        (double)n * n * k < (double)g_opts.multi_gpu_min_dim * g_opts.multi_gpu_min_dim * g_opts.multi_gpu_min_dim)
        return false;
    const Residency ra = classify(a), rc = classify(c);
    const bool host_source = on_host(ra);
    if (host_source != on_host(rc)) return false;           // mixed residency: single-GPU path
    const size_t es = sizeof(T);
    const bool nota = trans == 'N', upper = uplo == 'U';
    const bool beta0 = is0(beta);

    std::lock_guard<std::mutex> lock(g_mg.mu);
    TrackerGuard guard;
    if (!mg_init(ndev)) return false;
    MgState& st = g_mg;
    if (ra == RES_MANAGED) make_resident(a, (size_t)(((nota ? k : n) - 1) * lda + (nota ? n : k)) * es, current_stream());
    if (rc == RES_MANAGED) make_resident(c, (size_t)((int64_t)(n - 1) * ldc + n) * es, current_stream());
    BulkCall call(st, ndev);

    int64_t b[kMaxDevices + 1];
    ml3_strips(n, ndev, b);
    struct Geo { int64_t rows, w, ldp, ldt; bool in_place; };
    Geo geo[kMaxDevices];
    for (int s = 0; s < ndev; s++) {
        Geo& g = geo[s];
        g.rows = n - b[s]; g.w = b[s + 1] - b[s];
        g.in_place = s == 0 && !host_source;
        g.ldp = even_rows<T>(nota ? g.rows : k);                      // panel keeps the operand's orientation: rows x k ('N') or k x rows
        g.ldt = even_rows<T>(upper ? g.w : g.rows);                   // trapezoid: rows x w (lower) or w x rows (upper)
        if (g.in_place || g.w <= 0) continue;
        MgDev& d = st.dev[s];
        DeviceScope scope(d.id);
        ensure_cap(&d.panelA, &d.capA, (size_t)g.ldp * (nota ? k : g.rows) * es);
        ensure_cap(&d.ctile, &d.capC, (size_t)g.ldt * (upper ? g.rows : g.w) * es);
    }
    // ---- the row pieces of op(A) ----
    unsigned long long origin_bytes = 0, forward_bytes = 0;
    const std::vector<MgHop> plan = ml3_syrk_plan(ndev, n, host_source);
    std::vector<cudaEvent_t> arrived((size_t)ndev * 4096, nullptr);
    for (const MgHop& hp : plan) {
        const Geo& gd = geo[hp.dst];
        const int64_t loc = hp.off - b[hp.dst];                        // first row of the piece inside the receiver's panel
        size_t width, height, spitch;
        const char* src;
        char* dstp;
        if (nota) { width = (size_t)hp.len * es; height = (size_t)k; dstp = st.dev[hp.dst].panelA + (size_t)loc * es; }
        else      { width = (size_t)k * es; height = (size_t)hp.len; dstp = st.dev[hp.dst].panelA + (size_t)loc * gd.ldp * es; }
        if (hp.src < 0) {
            src = nota ? (const char*)(a + hp.off) : (const char*)(a + hp.off * lda);
            spitch = (size_t)lda * es;
        } else {
            const int64_t sloc = hp.off - b[hp.src];
            src = nota ? st.dev[hp.src].panelA + (size_t)sloc * es : st.dev[hp.src].panelA + (size_t)sloc * geo[hp.src].ldp * es;
            spitch = (size_t)geo[hp.src].ldp * es;
        }
        cudaEvent_t after = hp.src >= 0 ? arrived[(size_t)hp.src * 4096 + hp.piece] : nullptr;
        arrived[(size_t)hp.dst * 4096 + hp.piece] = call.hop(hp, host_source, dstp, (size_t)gd.ldp * es, src, spitch, width, height, after, &origin_bytes, &forward_bytes);
    }
    // ---- one masked GEMM per device ----
    const char ta = nota ? 'N' : 'T', tb = nota ? 'T' : 'N';
    for (int s = 0; s < ndev; s++) {
        const Geo& g = geo[s];
        if (g.w <= 0) continue;
        MgDev& d = st.dev[s];
        cudaStream_t stream = call.comp(s);
        call.wait_landed(s);
        DeviceScope scope(d.id);
        const T* P = g.in_place ? a : (const T*)d.panelA;
        const int64_t ldp = g.in_place ? lda : g.ldp;
        T* home = c + b[s] + b[s] * ldc;                               // the trapezoid's corner: its diagonal block comes first either way
        T* out = g.in_place ? home : (T*)d.ctile;
        const int64_t ldo = g.in_place ? ldc : g.ldt;
        const int64_t tm = upper ? g.w : g.rows, tn = upper ? g.rows : g.w;
        if (!g.in_place) {
            // the diagonal block is always read (its unreferenced half must survive the round trip), the rest only when beta != 0
            const int64_t lm = beta0 ? g.w : tm, ln = beta0 ? g.w : tn;
            B200_CUDA(cudaMemcpy2DAsync(out, (size_t)ldo * es, home, (size_t)ldc * es, (size_t)lm * es, (size_t)ln, cudaMemcpyDefault, stream));
            if (host_source) __atomic_fetch_add(&g_stats.h2d_bytes, (unsigned long long)(lm * ln * es), __ATOMIC_RELAXED);
        }
        gemm_dev<T>(stream, ta, tb, (int)tm, (int)tn, k, alpha, P, ldp, P, ldp, beta, out, ldo, upper ? MASK_UPPER : MASK_LOWER);
        if (!g.in_place) {
            B200_CUDA(cudaMemcpy2DAsync(home, (size_t)ldc * es, out, (size_t)ldo * es, (size_t)tm * es, (size_t)tn, cudaMemcpyDefault, stream));
            if (host_source) __atomic_fetch_add(&g_stats.d2h_bytes, (unsigned long long)(tm * tn * es), __ATOMIC_RELAXED);
        }
    }
    call.finish();
    g_mg_stats.calls++; g_mg_stats.devices = ndev; g_mg_stats.origin_bytes += origin_bytes; g_mg_stats.forward_bytes += forward_bytes;
    g_mg_stats.hops += plan.size();
    __atomic_fetch_add(&g_stats.hits, host_source ? 0ull : 2ull, __ATOMIC_RELAXED);
    __atomic_fetch_add(&g_stats.misses, host_source ? 2ull : 0ull, __ATOMIC_RELAXED);
    return true;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
bool multi_trxm(bool solve, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* a, int64_t lda, T* bmat, int64_t ldb) {
    const int ndev = g_opts.devices;
    if (ndev < 2 || is0(alpha)) return false;
    const bool lside = side == 'L', upper = uplo == 'U';
    const int64_t na = lside ? m : n, nfree = lside ? n : m;
    if (nfree < (int64_t)512 * ndev || na < 2048 ||
        (double)na * na * nfree < (double)g_opts.multi_gpu_min_dim * g_opts.multi_gpu_min_dim * g_opts.multi_gpu_min_dim)
        return false;
    const Residency ra = classify(a), rb = classify(bmat);
    const bool host_source = on_host(ra);
    if (host_source != on_host(rb)) return false;
    const size_t es = sizeof(T);

    std::lock_guard<std::mutex> lock(g_mg.mu);
    TrackerGuard guard;
    if (!mg_init(ndev)) return false;
    MgState& st = g_mg;
    if (ra == RES_MANAGED) make_resident(a, (size_t)((na - 1) * lda + na) * es, current_stream());
    if (rb == RES_MANAGED) make_resident(bmat, (size_t)((int64_t)(n - 1) * ldb + m) * es, current_stream());
    BulkCall call(st, ndev);

    struct Geo { int64_t lo, hi, bm, bn, ldt; bool in_place; };
    Geo geo[kMaxDevices];
    const int64_t ldp = even_rows<T>(na);
    for (int s = 0; s < ndev; s++) {
        Geo& g = geo[s];
        mg_block_range(nfree, ndev, s, &g.lo, &g.hi);
        g.bm = lside ? m : g.hi - g.lo; g.bn = lside ? g.hi - g.lo : n;
        g.ldt = even_rows<T>(g.bm);
        g.in_place = s == 0 && !host_source;
        if (g.in_place) continue;
        MgDev& d = st.dev[s];
        DeviceScope scope(d.id);
        ensure_cap(&d.panelA, &d.capA, (size_t)ldp * na * es);       // (every device is a link of the chains, block or not)
        if (g.hi > g.lo) ensure_cap(&d.ctile, &d.capC, (size_t)g.ldt * g.bn * es);
    }
    // ---- each device pulls its block of B with its own copy engine (device-resident: a peer read over NVLink; host: H2D) ----
    for (int s = 0; s < ndev; s++) {
        const Geo& g = geo[s];
        if (g.in_place || g.hi <= g.lo) continue;
        MgDev& d = st.dev[s];
        DeviceScope scope(d.id);
        const T* home = lside ? bmat + g.lo * ldb : bmat + g.lo;
        B200_CUDA(cudaMemcpy2DAsync(d.ctile, (size_t)g.ldt * es, home, (size_t)ldb * es, (size_t)g.bm * es, (size_t)g.bn, cudaMemcpyDefault, d.in));
        cudaEvent_t ev = next_event(d);
        B200_CUDA(cudaEventRecord(ev, d.in));
        call.landed[s].push_back(ev);
        if (host_source) __atomic_fetch_add(&g_stats.h2d_bytes, (unsigned long long)(g.bm * g.bn * es), __ATOMIC_RELAXED);
    }
    // ---- the referenced triangle of A, in column groups ----
    unsigned long long origin_bytes = 0, forward_bytes = 0;
    const std::vector<MgHop> plan = ml3_tri_plan(ndev, na, host_source);
    std::vector<cudaEvent_t> arrived((size_t)ndev * 4096, nullptr);
    for (const MgHop& hp : plan) {
        const int64_t r0 = upper ? 0 : hp.off, r1 = upper ? hp.off + hp.len : na;     // referenced rows of these columns
        const size_t off_p = (size_t)(r0 + hp.off * ldp) * es;
        const char* src = hp.src < 0 ? (const char*)(a + r0 + hp.off * lda) : st.dev[hp.src].panelA + off_p;
        const size_t spitch = (size_t)(hp.src < 0 ? lda : ldp) * es;
        MgDev& dd = st.dev[hp.dst];
        cudaEvent_t after = hp.src >= 0 ? arrived[(size_t)hp.src * 4096 + hp.piece] : nullptr;
        arrived[(size_t)hp.dst * 4096 + hp.piece] = call.hop(hp, host_source, dd.panelA + off_p, (size_t)ldp * es, src, spitch, (size_t)(r1 - r0) * es, (size_t)hp.len, after,
                                                             &origin_bytes, &forward_bytes);
    }
    // ---- the single-GPU recursion on every block ----
    for (int s = 0; s < ndev; s++) {
        const Geo& g = geo[s];
        if (g.hi <= g.lo) continue;
        MgDev& d = st.dev[s];
        cudaStream_t stream = call.comp(s);
        call.wait_landed(s);
        DeviceScope scope(d.id);
        const T* A = g.in_place ? a : (const T*)d.panelA;
        const int64_t la = g.in_place ? lda : ldp;
        T* home = lside ? bmat + g.lo * ldb : bmat + g.lo;
        T* blk = g.in_place ? home : (T*)d.ctile;
        const int64_t lb = g.in_place ? ldb : g.ldt;
        if (solve) trsm_dev<T>(stream, side, uplo, trans, diag, (int)g.bm, (int)g.bn, alpha, A, la, blk, lb);
        else trmm_dev<T>(stream, side, uplo, trans, diag, (int)g.bm, (int)g.bn, alpha, A, la, blk, lb);
        if (!g.in_place) {
            B200_CUDA(cudaMemcpy2DAsync(home, (size_t)ldb * es, blk, (size_t)lb * es, (size_t)g.bm * es, (size_t)g.bn, cudaMemcpyDefault, stream));
            if (host_source) __atomic_fetch_add(&g_stats.d2h_bytes, (unsigned long long)(g.bm * g.bn * es), __ATOMIC_RELAXED);
        }
    }
    call.finish();
    g_mg_stats.calls++; g_mg_stats.devices = ndev; g_mg_stats.origin_bytes += origin_bytes; g_mg_stats.forward_bytes += forward_bytes;
    g_mg_stats.hops += plan.size();
    __atomic_fetch_add(&g_stats.hits, host_source ? 0ull : 2ull, __ATOMIC_RELAXED);
    __atomic_fetch_add(&g_stats.misses, host_source ? 2ull : 0ull, __ATOMIC_RELAXED);
    return true;
}

#define B200_ML3_INST(T) \
    template bool multi_syrk<T>(char, char, int, int, T, const T*, int64_t, T, T*, int64_t); \
    template bool multi_trxm<T>(bool, char, char, char, char, int, int, T, const T*, int64_t, T*, int64_t);
B200_ML3_INST(float)
B200_ML3_INST(double)
B200_ML3_INST(cuFloatComplex)
B200_ML3_INST(cuDoubleComplex)

}  // namespace b200
