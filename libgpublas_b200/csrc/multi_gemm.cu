// multi_gemm.cu -- Level-3 calls partitioned over the GPUs of one box, INSIDE the interposed symbol
// (north_star (4); SURVEY.md section 5 "single process, 8 devices" and section 8(e)).
//
// The reference's whole point is that an unmodified program calls dgemm_ (blas_level3/gemm.cc:162-179) and gets the GPU; its
// own multi-GPU comparator is a preload too (tests/c/nvblas.conf:6-9, NVBLAS_GPU_LIST ALL0).  So `BLAS2CUDA_OPTIONS=devices=8`
// (or b200blas_set_options("devices=8")) makes ?gemm_ itself drive N devices from the calling thread -- no torchrun, no
// second process, nothing for the application to change.
//
// C := alpha*op(A)*op(B) + beta*C on a P x Q device grid (2 -> 1x2, 4 -> 2x2, 8 -> 2x4): device (p,q) owns
// C[row block p, column block q] and needs the row panel p of op(A) (all of k) and the column panel q of op(B).  k is never
// split, so there is no reduction and every C element is produced by one device with the same kernel and tile shape as the
// 1-GPU path: results are bit-identical to one GPU.
//
// Panel distribution -- the part round 1 got wrong.  Round 1 had the home GPU push every device's panels itself: 11.3 GB of
// home egress at N = 8 for 6.4 GB of operands, ~16 ms of a 35 ms step and the reason the 1->8 curve stopped at 0.876.
// Here a panel is cut into PIECES (A: row-groups of ~1/16 of the panel, B: bands of 2048 columns = 16 CTA tiles) and
// every piece travels along a CHAIN through the devices that need it (A row-group: the Q devices of its grid row; B band:
// the P devices of its grid column): the origin (home HBM, or host memory) is read ONCE per piece, every later hop is a
// forward by the device that has just received it, on that device's own copy engines and NVLink ports.  The first receiver
// rotates with the piece index, so all links carry equal shares:
//     device-resident operands : home egress 4.3 GB instead of 11.3 GB at N = 8 (each A/B byte leaves the home GPU once,
//                                except what the home GPU's own tile uses in place);
//     host-resident operands   : every GPU pulls 1/Q of its row's A pieces and 1/P of its column's B pieces over ITS OWN PCIe
//                                link (N x the H2D bandwidth of one GPU, total H2D bytes = operand bytes), the rest arrives
//                                over NVLink from the sibling that pulled it.
// Pieces are queued in the order the GEMM kernel's tile schedule consumes them; a 4-byte flag follows each piece on the same
// copy stream.  Every device runs ONE DGEMM launch whose TMA producer threads poll the flags of their tile's row-group and
// column band (gemm_f64.cu: DgemmParams::aflags), so transfer and compute overlap tile by tile inside one kernel; S/C/Z GEMM
// wait for their panels and run the ordinary kernel ("bulk").  C tiles return by peer stores from the epilogue into the home
// allocation (device-resident C) or by a copy-engine transfer of the finished tile (managed / host C).
// Cross-device ordering uses CUDA events only (one process), completion is an event per device that the caller's stream
// waits on: no NCCL, no host synchronisation, no barrier on the data path.
#include "abi_common.h"
#include "gemm_generic.cuh"
#include "multi_gemm.h"
#include "multi_state.h"
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>
#include <unistd.h>
#include <time.h>
#include <cstdio>

namespace b200 {

// ---------------------------------------------------------------------------------------------
// pure host logic (also exported for the CPU tests): grid, block ranges, piece sizes, chains
void mg_grid(int ndev, int* P, int* Q) {
    int p = (int)std::sqrt((double)ndev);
    while (p > 1 && ndev % p) p--;
    if (p < 1) p = 1;
    *P = p; *Q = ndev / p;
}
void mg_block_range(int64_t total, int parts, int idx, int64_t* lo, int64_t* hi) {
    // nearly equal blocks, multiples of 128 (the CTA tile) when the dimension allows it
    const int64_t base = total >= (int64_t)128 * parts ? ((total / parts + 127) / 128) * 128 : (total + parts - 1) / parts;
    *lo = std::min<int64_t>(total, base * idx);
    *hi = idx < parts - 1 ? std::min<int64_t>(total, base * (idx + 1)) : total;
    if (*hi < *lo) *hi = *lo;
}
int64_t mg_a_group(int64_t tm, int pieces) { return std::max<int64_t>(128, ((tm / pieces + 127) / 128) * 128); }
int64_t mg_b_group() { return 1024; }      // half a band of the kernel's tile schedule (16 tile columns): first tiles start sooner

// Hops of the whole call in issue order.  Positions follow the tile schedule of the GEMM kernel (bands of 16 tile-columns
// walked down the rows): B piece 0, A row-group 0, B piece 1, then the A row-groups top to bottom, then the remaining B pieces;
// at every position the grid rows / columns are interleaved so all devices receive their first pieces at the same time.
std::vector<MgHop> mg_plan(int ndev, int64_t m, int64_t n, bool host_source, int a_pieces) {
    int P, Q;
    mg_grid(ndev, &P, &Q);
    std::vector<MgHop> plan;
    int64_t max_ag = 0, max_bb = 0;
    for (int p = 0; p < P; p++) { int64_t lo, hi; mg_block_range(m, P, p, &lo, &hi); if (hi > lo) max_ag = std::max(max_ag, (hi - lo + mg_a_group(hi - lo, a_pieces) - 1) / mg_a_group(hi - lo, a_pieces)); }
    for (int q = 0; q < Q; q++) { int64_t lo, hi; mg_block_range(n, Q, q, &lo, &hi); max_bb = std::max(max_bb, (hi - lo + mg_b_group() - 1) / mg_b_group()); }
    auto chain = [&](int kind, int gidx, int piece, int64_t off, int64_t len) {
        // members of the chain: the devices that consume this piece
        std::vector<int> members;
        if (kind == 0) { for (int q = 0; q < Q; q++) members.push_back(gidx * Q + q); }
        else           { for (int p = 0; p < P; p++) members.push_back(p * Q + gidx); }
        std::vector<int> order;
        const int cnt = (int)members.size();
        bool home_member = false;
        for (int s : members) home_member = home_member || s == 0;
        if (!host_source && home_member) {
            // the home GPU (slot 0) uses its operands in place and is the origin: it hands the piece to a rotating first sibling
            std::vector<int> others;
            for (int s : members) if (s != 0) others.push_back(s);
            const int oc = (int)others.size();
            for (int i = 0; i < oc; i++) order.push_back(others[(piece + gidx + i) % oc]);
        } else {
            for (int i = 0; i < cnt; i++) order.push_back(members[(piece + gidx + i) % cnt]);     // + gidx: panels of a single piece still spread
        }
        int prev = -1;     // -1: the origin (home HBM / host memory)
        for (int s : order) {
            // a device with an empty tile still forwards (it is a link of the chain) -- tiles are never empty for the sizes
            // the selector admits, so no special case is needed
            plan.push_back(MgHop{kind, gidx, piece, off, len, prev, s});
            prev = s;
        }
    };
    auto b_piece = [&](int h) {
        for (int q = 0; q < Q; q++) {
            int64_t lo, hi; mg_block_range(n, Q, q, &lo, &hi);
            const int64_t off = (int64_t)h * mg_b_group();
            if (off < hi - lo) chain(1, q, h, off, std::min<int64_t>(mg_b_group(), hi - lo - off));
        }
    };
    auto a_piece = [&](int64_t g) {
        for (int p = 0; p < P; p++) {
            int64_t lo, hi; mg_block_range(m, P, p, &lo, &hi);
            const int64_t ag = mg_a_group(hi - lo, a_pieces), off = g * ag;
            if (hi > lo && off < hi - lo) chain(0, p, (int)g, off, std::min<int64_t>(ag, hi - lo - off));
        }
    };
    // the first tiles (tile row 0, tile columns 0..7) need B piece 0 and A row-group 0; the rest of band 0 needs B piece 1;
    // then the schedule walks down the rows of band 0 (A row-groups in order) before it touches band 1 (B pieces 2, 3, ...)
    b_piece(0);
    if (max_ag > 0) a_piece(0);
    if (max_bb > 1) b_piece(1);
    for (int64_t g = 1; g < max_ag; g++) a_piece(g);
    for (int64_t h = 2; h < max_bb; h++) b_piece((int)h);
    return plan;
}

// ---------------------------------------------------------------------------------------------
// per-device streams / panels and the process-wide state: multi_state.h (shared with multi_level3.cu)
MgState g_mg;

cudaEvent_t next_event(MgDev& d) {
    if (d.next_event == d.events.size()) {
        cudaEvent_t e;
        B200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        d.events.push_back(e);
    }
    return d.events[d.next_event++];
}
void ensure_cap(char** p, size_t* cap, size_t need) {
    if (need <= *cap) return;
    if (*p) B200_CUDA(cudaFree(*p));
    size_t ncap = (need + ((size_t)64 << 20) - 1) & ~(((size_t)64 << 20) - 1);
    B200_CUDA(cudaMalloc((void**)p, ncap));
    *cap = ncap;
}

// brings up `ndev` devices (home first), streams, flag arrays, peer access between every pair; false if the box cannot
bool mg_init(int ndev) {
    MgState& st = g_mg;
    if (st.ready && st.ndev >= ndev) return true;       // st.ndev: devices brought up so far (callers use their own count)
    if (st.failed) return false;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < ndev || ndev > kMaxDevices) { cudaGetLastError(); st.failed = true; return false; }
    const int home = home_device();
    int ids[kMaxDevices], nid = 0;
    ids[nid++] = home;
    for (int d = 0; d < count && nid < ndev; d++) if (d != home) ids[nid++] = d;
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < ndev; j++) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ids[i], ids[j]) != cudaSuccess || !can) {
                cudaGetLastError();
                b200_writef(STDERR_FILENO, "b200blas: devices=%d ignored: GPU %d cannot access GPU %d (peer access is required)\n", ndev, ids[i], ids[j]);
                st.failed = true;
                return false;
            }
        }
    if (!st.host_consts) {
        B200_CUDA(cudaMallocHost((void**)&st.host_consts, 65536 * sizeof(uint32_t)));
        for (uint32_t v = 0; v < 65536; v++) st.host_consts[v] = v;
    }
    for (int i = st.ready ? st.ndev : 0; i < ndev; i++) {
        MgDev& d = st.dev[i];
        d.id = ids[i];
        DeviceScope scope(d.id);
        for (int j = 0; j < ndev; j++) {
            if (i == j) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(ids[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) fatal("cudaDeviceEnablePeerAccess", __FILE__, __LINE__, cudaGetErrorString(e));
            cudaGetLastError();
        }
        int lo = 0, hi = 0;
        B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        B200_CUDA(cudaStreamCreateWithFlags(&d.comp, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&d.in, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&d.out, cudaStreamNonBlocking));
        B200_CUDA(cudaMalloc((void**)&d.flags, 4096 * sizeof(uint32_t)));
        B200_CUDA(cudaMemset(d.flags, 0, 4096 * sizeof(uint32_t)));
        B200_CUDA(cudaMalloc((void**)&d.consts, 65536 * sizeof(uint32_t)));
        B200_CUDA(cudaMemcpy(d.consts, st.host_consts, 65536 * sizeof(uint32_t), cudaMemcpyHostToDevice));
        B200_CUDA(cudaEventCreateWithFlags(&d.done, cudaEventDisableTiming));
        B200_CUDA(cudaDeviceSynchronize());
    }
    // devices brought up by an earlier, smaller call still need peer access to the new ones
    if (st.ready && st.ndev < ndev)
        for (int i = 0; i < st.ndev; i++) {
            DeviceScope scope(st.dev[i].id);
            for (int j = st.ndev; j < ndev; j++) { cudaDeviceEnablePeerAccess(ids[j], 0); cudaGetLastError(); }
        }
    st.ndev = ndev; st.ready = true;
    if (g_opts.debug_exec) b200_writef(STDERR_FILENO, "b200blas: partitioned Level-3 over %d devices (home %d), peer access enabled\n", ndev, home);
    return true;
}

// ONE push stream per device for the GEMM's pieces: a single peer copy already saturates the device's NVLink egress
// (measured 770 GB/s for one stream), and with one stream the pieces leave in exactly the order they were queued -- the
// order the consumers need them.  (A stream per destination let the copy engines run whole queues one after the other:
// some devices had all their pieces after 2 ms, others saw their first one after 3.4 ms; profiles/r02e_mg_debug8.txt.)
cudaStream_t push_stream(MgDev& d) {
    if (!d.push) {
        DeviceScope scope(d.id);
        B200_CUDA(cudaStreamCreateWithFlags(&d.push, cudaStreamNonBlocking));
    }
    return d.push;
}
cudaStream_t fwd_stream(MgDev& d, int dst_slot) {
    if (!d.fwd[dst_slot]) {
        DeviceScope scope(d.id);
        B200_CUDA(cudaStreamCreateWithFlags(&d.fwd[dst_slot], cudaStreamNonBlocking));
    }
    return d.fwd[dst_slot];
}

namespace {
template <typename T> struct GemmFn;
template <> struct GemmFn<float> { static constexpr auto fn = sgemm_dev; };
template <> struct GemmFn<double> { static constexpr auto fn = dgemm_dev; };
template <> struct GemmFn<cuFloatComplex> { static constexpr auto fn = cgemm_dev; };
template <> struct GemmFn<cuDoubleComplex> { static constexpr auto fn = zgemm_dev; };
}  // namespace

MgStats g_mg_stats = {0, 0, 0, 0, 0};

// B200BLAS_MG_TRACE=1: host-side timeline of one partitioned call (when each hop landed, when each device's kernel could start
// and when it finished), from polling the events -- ~20 us resolution, debugging only.
namespace {
struct TraceItem { char label[80]; cudaEvent_t ev; double t_ms; };
double now_ms() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
void trace_poll(std::vector<TraceItem>& items, double t0) {
    size_t left = items.size();
    for (auto& it : items) it.t_ms = -1;
    while (left) {
        for (auto& it : items) {
            if (it.t_ms >= 0) continue;
            cudaError_t e = cudaEventQuery(it.ev);
            if (e == cudaSuccess) { it.t_ms = now_ms() - t0; left--; }
            else if (e != cudaErrorNotReady) { cudaGetLastError(); it.t_ms = 1e9; left--; }
        }
    }
    std::sort(items.begin(), items.end(), [](const TraceItem& a, const TraceItem& b) { return a.t_ms < b.t_ms; });
    for (auto& it : items) b200_writef(STDERR_FILENO, "mgtrace %9.3f ms  %s\n", it.t_ms, it.label);
}
}  // namespace

// Returns false when the call should take the single-GPU path (too small for the grid, mixed residency, no peer access).
template <typename T>
bool multi_gemm(char ta, char tb, int m, int n, int k, T alpha, const T* a, int64_t lda, const T* b, int64_t ldb, T beta, T* c, int64_t ldc) {
    const int ndev = g_opts.devices;
    if (ndev < 2) return false;
    int P, Q;
    mg_grid(ndev, &P, &Q);
    // every device needs a tile of at least 1024 x 1024 and enough work to hide the distribution
    if ((int64_t)m < (int64_t)1024 * P || (int64_t)n < (int64_t)1024 * Q || (int64_t)k < 1024 ||
        (double)m * n * k < (double)g_opts.multi_gpu_min_dim * g_opts.multi_gpu_min_dim * g_opts.multi_gpu_min_dim)
        return false;
    const Residency ra = classify(a), rb = classify(b), rc = classify(c);
    auto host = [](Residency r) { return r == RES_HOST_PINNED || r == RES_HOST_PAGEABLE; };
    const bool host_source = host(ra) && host(rb);
    if (!host_source && (host(ra) || host(rb))) return false;          // mixed operand residency: single-GPU path
    if (host_source != host(rc)) return false;
    constexpr bool fused = std::is_same<T, double>::value;               // DGEMM: flag-polling kernel; others: bulk
    const size_t es = sizeof(T);
    const bool nota = ta == 'N', notb = tb == 'N';
    const bool beta0 = is0(beta);

    std::lock_guard<std::mutex> lock(g_mg.mu);
    TrackerGuard guard;
    if (!mg_init(ndev)) return false;
    static const bool tracing = getenv("B200BLAS_MG_TRACE") != nullptr;
    std::vector<TraceItem> trace;
    const double trace_t0 = tracing ? now_ms() : 0.0;
    auto trace_mark = [&](MgDev& d, cudaStream_t stream, const char* fmt, int a0, int a1, int a2, int a3) {
        if (!tracing) return;
        TraceItem it;
        snprintf(it.label, sizeof it.label, fmt, a0, a1, a2, a3);
        it.ev = next_event(d);
        B200_CUDA(cudaEventRecord(it.ev, stream));
        it.t_ms = -1;
        trace.push_back(it);
    };
    MgState& st = g_mg;
    cudaStream_t home_stream = current_stream();
    if (ra == RES_MANAGED) make_resident(a, (size_t)(((nota ? k : m) - 1) * lda + (nota ? m : k)) * es, home_stream);
    if (rb == RES_MANAGED) make_resident(b, (size_t)(((notb ? n : k) - 1) * ldb + (notb ? k : n)) * es, home_stream);
    if (rc == RES_MANAGED) make_resident(c, (size_t)((int64_t)(n - 1) * ldc + m) * es, home_stream);
    // C returns through peer stores only when it is plain device memory; managed / host C gets a copy-engine transfer
    const bool peer_store = fused && rc == RES_DEVICE;

    if (st.epoch >= 65535) {      // flag values are 16-bit table entries: restart the epoch counter
        for (int i = 0; i < ndev; i++) { DeviceScope scope(st.dev[i].id); B200_CUDA(cudaDeviceSynchronize()); B200_CUDA(cudaMemset(st.dev[i].flags, 0, 4096 * sizeof(uint32_t))); }
        st.epoch = 0;
    }
    const uint32_t epoch = ++st.epoch;

    // ---- geometry per slot ----
    struct Geo { int64_t r0, r1, c0, c1, tm, tn, lda_p, ldb_p, ldc_t; bool in_place; };
    Geo geo[kMaxDevices];
    for (int s = 0; s < ndev; s++) {
        Geo& g = geo[s];
        mg_block_range(m, P, s / Q, &g.r0, &g.r1); mg_block_range(n, Q, s % Q, &g.c0, &g.c1);
        g.tm = g.r1 - g.r0; g.tn = g.c1 - g.c0;
        g.in_place = (s == 0 && !host_source);
        const int64_t per16 = std::max<int64_t>(1, 16 / (int64_t)es);
        auto even = [&](int64_t v) { return (v + per16 - 1) / per16 * per16; };
        g.lda_p = even(nota ? g.tm : k);          // panel A keeps the operand's storage orientation: tm x k ('N') or k x tm
        g.ldb_p = even(notb ? k : g.tn);          // panel B: k x tn ('N') or tn x k
        g.ldc_t = even(g.tm);
    }
    // ---- start: every stream of every device waits for the caller's stream (inputs are ready, previous call has finished) ----
    cudaEvent_t start;
    { MgDev& h = st.dev[0]; h.next_event = 0; start = next_event(h); B200_CUDA(cudaEventRecord(start, home_stream)); }
    for (int s = 0; s < ndev; s++) {
        MgDev& d = st.dev[s];
        if (s) d.next_event = 0;
        DeviceScope scope(d.id);
        if (s) ws_reset();               // (the home context was reset by the entry point; its workspace may already hold a staged operand)
        const Geo& g = geo[s];
        if (!g.in_place) {
            ensure_cap(&d.panelA, &d.capA, (size_t)g.lda_p * (nota ? k : g.tm) * es);
            ensure_cap(&d.panelB, &d.capB, (size_t)g.ldb_p * (notb ? g.tn : k) * es);
        }
        if (!(peer_store) && !(s == 0 && !host_source)) ensure_cap(&d.ctile, &d.capC, (size_t)g.ldc_t * g.tn * es);
        B200_CUDA(cudaStreamWaitEvent(d.comp, start, 0));
        B200_CUDA(cudaStreamWaitEvent(d.in, start, 0));
        B200_CUDA(cudaStreamWaitEvent(d.out, start, 0));
        B200_CUDA(cudaStreamWaitEvent(push_stream(d), start, 0));
    }

    // ---- compute launches first (DGEMM): the kernels spin on the flags while the copies below are being queued ----
    auto a_ptr = [&](int s) -> const T* { return geo[s].in_place ? (nota ? a + geo[s].r0 : a + geo[s].r0 * lda) : (const T*)st.dev[s].panelA; };
    auto b_ptr = [&](int s) -> const T* { return geo[s].in_place ? (notb ? b + geo[s].c0 * ldb : b + geo[s].c0) : (const T*)st.dev[s].panelB; };
    auto a_ld = [&](int s) { return geo[s].in_place ? lda : geo[s].lda_p; };
    auto b_ld = [&](int s) { return geo[s].in_place ? ldb : geo[s].ldb_p; };
    T* c_home[kMaxDevices];
    for (int s = 0; s < ndev; s++) c_home[s] = c + geo[s].r0 + geo[s].c0 * ldc;
    // where slot s computes: straight into the caller's C (peer store / the home tile of a device-resident call) or into its ctile
    auto direct = [&](int s) { return peer_store || (s == 0 && !host_source); };

    // Detached tiles of a fused (DGEMM) call with host-resident C are computed in two column halves: the first half travels
    // device->host on the device's `out` stream underneath the second half's compute (PCIe is full duplex and the H2D side
    // is busy with pieces anyway).  Two launches of 6.9 waves each cost 0.2 wave more than one of 13.8; four would cost 2.
    // returned[s]: how many columns of slot s's tile have already been sent home by launch().
    int64_t returned[kMaxDevices] = {};
    auto launch = [&](int s, cudaStream_t stream) {
        MgDev& d = st.dev[s];
        const Geo& g = geo[s];
        if (g.tm <= 0 || g.tn <= 0) return;
        DeviceScope scope(d.id);
        T* out = direct(s) ? c_home[s] : (T*)d.ctile;
        const int64_t ldo = direct(s) ? ldc : g.ldc_t;
        T eff_beta = beta;
        if (!direct(s) && !beta0) {
            // beta != 0 with a detached tile: bring the old C tile in first (same stream, ahead of the kernel)
            B200_CUDA(cudaMemcpy2DAsync(d.ctile, (size_t)g.ldc_t * es, c_home[s], (size_t)ldc * es, (size_t)g.tm * es, (size_t)g.tn, cudaMemcpyDefault, stream));
        }
        trace_mark(d, stream, "dev %d kernel may start", s, 0, 0, 0);
        if constexpr (fused) {
            const int64_t bg = mg_b_group();
            const int64_t half = (host_source && !direct(s) && g.tn >= 4 * bg) ? (g.tn / 2) / (2 * bg) * (2 * bg) : 0;     // whole bands of the tile schedule
            for (int part = 0; part < (half ? 2 : 1); part++) {
                const int64_t c0 = part ? half : 0, c1 = (half && !part) ? half : g.tn;
                if (!g.in_place) dgemm_set_panel_flags(d.flags, (int)mg_a_group(g.tm, 16), d.flags + 2048 + c0 / bg, (int)bg, epoch);
                const T* bp = notb ? b_ptr(s) + c0 * b_ld(s) : b_ptr(s) + c0;
                dgemm_out_dev(stream, ta, tb, (int)g.tm, (int)(c1 - c0), k, alpha, a_ptr(s), a_ld(s), bp, b_ld(s), eff_beta, out + c0 * ldo, ldo, out + c0 * ldo, ldo, MASK_FULL);
                if (half && !part) {
                    cudaEvent_t e = next_event(d);
                    B200_CUDA(cudaEventRecord(e, stream));
                    B200_CUDA(cudaStreamWaitEvent(d.out, e, 0));
                    B200_CUDA(cudaMemcpy2DAsync(c_home[s], (size_t)ldc * es, d.ctile, (size_t)g.ldc_t * es, (size_t)g.tm * es, (size_t)half, cudaMemcpyDefault, d.out));
                    returned[s] = half;
                }
            }
        } else {
            GemmFn<T>::fn(stream, ta, tb, (int)g.tm, (int)g.tn, k, alpha, a_ptr(s), a_ld(s), b_ptr(s), b_ld(s), eff_beta, out, ldo, MASK_FULL);
        }
        trace_mark(d, stream, "dev %d kernel done", s, 0, 0, 0);
    };
    if (fused)
        for (int s = 0; s < ndev; s++) launch(s, s == 0 ? home_stream : st.dev[s].comp);
    // bulk types: columns/rows [k0, k0 + kk) of the panels; the first chunk applies beta (after bringing a detached tile's old C in)
    auto launch_chunk = [&](int s, cudaStream_t stream, int64_t k0, int64_t kk, bool first) {
        MgDev& d = st.dev[s];
        const Geo& g = geo[s];
        if (g.tm <= 0 || g.tn <= 0) return;
        DeviceScope scope(d.id);
        T* out = direct(s) ? c_home[s] : (T*)d.ctile;
        const int64_t ldo = direct(s) ? ldc : g.ldc_t;
        if (first && !direct(s) && !beta0)
            B200_CUDA(cudaMemcpy2DAsync(d.ctile, (size_t)g.ldc_t * es, c_home[s], (size_t)ldc * es, (size_t)g.tm * es, (size_t)g.tn, cudaMemcpyDefault, stream));
        const T* ap = nota ? a_ptr(s) + k0 * a_ld(s) : a_ptr(s) + k0;
        const T* bp = notb ? b_ptr(s) + k0 : b_ptr(s) + k0 * b_ld(s);
        if (first) trace_mark(d, stream, "dev %d kernel may start", s, 0, 0, 0);
        GemmFn<T>::fn(stream, ta, tb, (int)g.tm, (int)g.tn, (int)kk, alpha, ap, a_ld(s), bp, b_ld(s), first ? beta : num<T>::real(1.0), out, ldo, MASK_FULL);
        if (k0 + kk >= k) trace_mark(d, stream, "dev %d kernel done", s, 0, 0, 0);
    };

    // ---- the piece chains ----
    // Bulk types (S/C/Z) CAN consume k in NCH chunks -- every piece travels chunk by chunk, a device multiplies chunk c (beta on the
    // first chunk, then accumulating into its tile) as soon as that chunk of its panels has landed, while chunk c+1 is still on
    // the wires.  (One chunk: at N = 8 the SGEMM 16384^3 call was 2.5 ms of transfers, then the split pass, then 4.7 ms of tensor
    // work in series -- 4.3x; profiles/r02j_bench_n8.json.)  Coarser A pieces keep the hop count per call the same.
    static const int nch_env = getenv("B200BLAS_MG_KCHUNKS") ? atoi(getenv("B200BLAS_MG_KCHUNKS")) : 0;      // experiments: force the chunk count
    // Measured, and OFF by default: each chunk costs a pass of the tile's epilogues (not overlapped with the next tile's main loop
    // in the tcgen05 kernel), a split pass launch and a round of hops.  SGEMM 16384^3 at N = 2: 19.6 ms in one chunk, 20.4 in two,
    // 24.1 in four (profiles/r02r_kchunks_n2.txt); at N = 8 four chunks took 22.2 ms against 9.5 ms in one (profiles/r02v_bench_n8.json
    // vs r02j_bench_n8.json).  B200BLAS_MG_KCHUNKS=<n> keeps the path reachable (tests run it with 3 chunks).
    const int NCH = fused ? 1 : std::min(nch_env > 0 ? nch_env : 1, std::max(1, k / 1024));
    const int64_t kstep = ((k + NCH - 1) / NCH + 255) / 256 * 256;
    const std::vector<MgHop> plan = mg_plan(ndev, m, n, host_source, NCH > 1 ? 4 : 16);
    // arrival events: [slot][kind][piece]
    std::vector<cudaEvent_t> arrived((size_t)ndev * 2 * 2048, nullptr);
    auto arr = [&](int s, int kind, int piece) -> cudaEvent_t& { return arrived[((size_t)s * 2 + kind) * 2048 + piece]; };
    unsigned long long origin_bytes = 0, forward_bytes = 0;
    for (int ch = 0; ch < NCH; ch++) {
    const int64_t k0 = (int64_t)ch * kstep, kk = std::min<int64_t>(kstep, k - k0);
    if (kk <= 0) break;
    std::fill(arrived.begin(), arrived.end(), nullptr);
    for (const MgHop& hp : plan) {
        MgDev& dst = st.dev[hp.dst];
        const Geo& gd = geo[hp.dst];
        // geometry of the piece inside a panel / inside the original operand
        size_t width, height, spitch, dpitch;
        const char* src; char* dstp;
        // (k range [k0, k0 + kk) of the piece: columns of an 'N' A / rows of a 'T' A, rows of an 'N' B / columns of a 'T' B)
        if (hp.kind == 0) {       // rows [off, off+len) of op(A)'s row block
            if (nota) { width = (size_t)hp.len * es; height = (size_t)kk; dstp = dst.panelA + ((size_t)hp.off + (size_t)k0 * gd.lda_p) * es; dpitch = (size_t)gd.lda_p * es; }
            else      { width = (size_t)kk * es; height = (size_t)hp.len; dstp = dst.panelA + ((size_t)hp.off * gd.lda_p + (size_t)k0) * es; dpitch = (size_t)gd.lda_p * es; }
            if (hp.src < 0) {
                src = nota ? (const char*)(a + gd.r0 + hp.off + k0 * lda) : (const char*)(a + (gd.r0 + hp.off) * lda + k0);
                spitch = (size_t)lda * es;
            } else {
                const Geo& gs = geo[hp.src];
                src = nota ? st.dev[hp.src].panelA + ((size_t)hp.off + (size_t)k0 * gs.lda_p) * es : st.dev[hp.src].panelA + ((size_t)hp.off * gs.lda_p + (size_t)k0) * es;
                spitch = (size_t)gs.lda_p * es;
            }
        } else {                  // columns [off, off+len) of op(B)'s column block
            if (notb) { width = (size_t)kk * es; height = (size_t)hp.len; dstp = dst.panelB + ((size_t)hp.off * gd.ldb_p + (size_t)k0) * es; dpitch = (size_t)gd.ldb_p * es; }
            else      { width = (size_t)hp.len * es; height = (size_t)kk; dstp = dst.panelB + ((size_t)hp.off + (size_t)k0 * gd.ldb_p) * es; dpitch = (size_t)gd.ldb_p * es; }
            if (hp.src < 0) {
                src = notb ? (const char*)(b + (gd.c0 + hp.off) * ldb + k0) : (const char*)(b + gd.c0 + hp.off + k0 * ldb);
                spitch = (size_t)ldb * es;
            } else {
                const Geo& gs = geo[hp.src];
                src = notb ? st.dev[hp.src].panelB + ((size_t)hp.off * gs.ldb_p + (size_t)k0) * es : st.dev[hp.src].panelB + ((size_t)hp.off + (size_t)k0 * gs.ldb_p) * es;
                spitch = (size_t)gs.ldb_p * es;
            }
        }
        uint32_t* flag = dst.flags + (hp.kind ? 2048 : 0) + hp.piece;
        cudaStream_t stream;
        int exec_slot;            // the device whose stream carries this hop
        if (hp.src < 0 && host_source) { exec_slot = hp.dst; stream = dst.in; }                              // H2D over the receiver's own PCIe link
        else { exec_slot = hp.src < 0 ? 0 : hp.src; stream = push_stream(st.dev[exec_slot]); }               // push over NVLink by whoever holds the piece
        MgDev& ex = st.dev[exec_slot];
        DeviceScope scope(ex.id);
        if (hp.src >= 0) B200_CUDA(cudaStreamWaitEvent(stream, arr(hp.src, hp.kind, hp.piece), 0));
        B200_CUDA(cudaMemcpy2DAsync(dstp, dpitch, src, spitch, width, height, cudaMemcpyDefault, stream));
        if (hp.src < 0 && host_source) B200_CUDA(cudaMemcpyAsync(flag, st.host_consts + epoch, 4, cudaMemcpyHostToDevice, stream));
        else B200_CUDA(cudaMemcpyAsync(flag, ex.consts + epoch, 4, cudaMemcpyDefault, stream));
        cudaEvent_t ev = next_event(ex);
        B200_CUDA(cudaEventRecord(ev, stream));
        arr(hp.dst, hp.kind, hp.piece) = ev;
        if (tracing) { TraceItem it; snprintf(it.label, sizeof it.label, "%s piece %d: %d -> %d landed", hp.kind ? "B" : "A", hp.piece, hp.src, hp.dst); it.ev = ev; it.t_ms = -1; trace.push_back(it); }
        if (hp.src < 0) origin_bytes += (unsigned long long)width * height; else forward_bytes += (unsigned long long)width * height;
        if (hp.src < 0 && host_source) __atomic_fetch_add(&g_stats.h2d_bytes, (unsigned long long)(width * height), __ATOMIC_RELAXED);
    }
    // ---- bulk types: the ordinary kernel on this chunk of k once the chunk's pieces of the device's panels have landed ----
    if (!fused)
        for (int s = 0; s < ndev; s++) {
            cudaStream_t stream = s == 0 ? home_stream : st.dev[s].comp;
            if (!geo[s].in_place) {
                DeviceScope scope(st.dev[s].id);
                for (int kind = 0; kind < 2; kind++)
                    for (int piece = 0; piece < 2048; piece++)
                        if (arr(s, kind, piece)) B200_CUDA(cudaStreamWaitEvent(stream, arr(s, kind, piece), 0));
            }
            launch_chunk(s, stream, k0, kk, ch == 0);
        }
    }   // chunks of k
    // ---- C return for detached tiles, completion ----
    for (int s = 0; s < ndev; s++) {
        MgDev& d = st.dev[s];
        const Geo& g = geo[s];
        cudaStream_t stream = s == 0 ? home_stream : d.comp;
        DeviceScope scope(d.id);
        if (!direct(s) && g.tm > 0 && g.tn > 0) {
            const int64_t c0 = returned[s];          // columns [0, c0) went home under the second half's compute
            B200_CUDA(cudaMemcpy2DAsync(c_home[s] + c0 * ldc, (size_t)ldc * es, (T*)d.ctile + c0 * g.ldc_t, (size_t)g.ldc_t * es, (size_t)g.tm * es, (size_t)(g.tn - c0), cudaMemcpyDefault, stream));
            if (host_source) __atomic_fetch_add(&g_stats.d2h_bytes, (unsigned long long)(g.tm * g.tn * es), __ATOMIC_RELAXED);
            if (c0) { cudaEvent_t e = next_event(d); B200_CUDA(cudaEventRecord(e, d.out)); B200_CUDA(cudaStreamWaitEvent(stream, e, 0)); }
        }
        // a device's push stream must drain before the call is over too (the next call reuses the panels)
        { cudaEvent_t e = next_event(d); B200_CUDA(cudaEventRecord(e, push_stream(d))); B200_CUDA(cudaStreamWaitEvent(stream, e, 0)); }
        { cudaEvent_t e = next_event(d); B200_CUDA(cudaEventRecord(e, d.in)); B200_CUDA(cudaStreamWaitEvent(stream, e, 0)); }
        if (s) {
            B200_CUDA(cudaEventRecord(d.done, stream));
        }
    }
    for (int s = 1; s < ndev; s++) B200_CUDA(cudaStreamWaitEvent(home_stream, st.dev[s].done, 0));
    if (tracing) {
        b200_writef(STDERR_FILENO, "mgtrace enqueue took %.3f ms (%zu hops)\n", now_ms() - trace_t0, plan.size());
        trace_poll(trace, trace_t0);
    }
    last_variant = fused ? VAR_DMMA_TMA : last_variant;
    g_mg_stats.calls++; g_mg_stats.devices = ndev; g_mg_stats.origin_bytes += origin_bytes; g_mg_stats.forward_bytes += forward_bytes;
    g_mg_stats.hops += plan.size();
    __atomic_fetch_add(&g_stats.hits, host_source ? 0ull : 3ull, __ATOMIC_RELAXED);
    __atomic_fetch_add(&g_stats.misses, host_source ? 3ull : 0ull, __ATOMIC_RELAXED);
    return true;
}

// ---------------------------------------------------------------------------------------------
// Blocked Cholesky WORKLOAD over the devices of the box (BASELINE.json configs[3]: "DSYRK + DTRSM + DGEMM panels, n = 32768,
// sharded across 8 B200"; SURVEY.md section 8(e) row "Blocked Cholesky").  Round 1 ran this as one process per GPU with NCCL
// panel broadcasts (cholesky.py: TiledCholesky, 195 ms = 2.4x at N = 8): an NCCL broadcast kernel cannot get SMs while the
// rank-nb update occupies all of them, so the panel chain serialised.  Here one thread drives every device:
//   * 1-D block-cyclic ownership of the nb-wide block columns (column J on device J mod N; the home GPU works in place);
//     the other devices' columns leave the home GPU once, in the order they are needed, under the first steps;
//   * step J: the owner factors the diagonal block (potrf.cu) and solves the panel below it (trsm_dev) on a HIGH-PRIORITY
//     stream, so these latency-bound kernels get SMs in between the CTAs of the update running on the same device;
//   * the panel travels round the ring owner -> owner+1 -> ... in row pieces, each hop a copy-engine push by the device that
//     has just received the piece (copy engines need no SMs); the next owner is the first to receive it;
//   * every device applies the rank-nb update (one masked DMMA GEMM per block column: SYRK on the diagonal tile, GEMM
//     below) to the columns it owns, the next panel's column FIRST (look-ahead 1), then the others;
//   * a finished column returns to the home allocation as soon as it is factored, under the remaining steps.
// All ordering is CUDA events; nothing synchronises the host until the info words are read.
int multi_cholesky_lower(int n, double* a, int64_t lda, int nb, int ndev_req) {
    if (n <= 0) return 0;
    int ndev = ndev_req < 1 ? 1 : ndev_req;
    std::lock_guard<std::mutex> lock(g_mg.mu);
    TrackerGuard guard;
    if (!mg_init(ndev > 1 ? ndev : 1)) { ndev = 1; if (!mg_init(1)) fatal("multi_cholesky_lower", __FILE__, __LINE__, "device set-up failed"); }
    MgState& st = g_mg;
    cudaStream_t home_stream = current_stream();
    if (nb < 128) nb = 128;
    nb = (nb + 127) / 128 * 128;
    const int NB = (n + nb - 1) / nb;
    const int64_t ldw = ((int64_t)n + 1) / 2 * 2;
    auto own = [&](int J) { return J % ndev; };
    auto loc = [&](int J) { return J / ndev; };
    static const int piece_rows = getenv("B200BLAS_CHOL_PIECE") ? atoi(getenv("B200BLAS_CHOL_PIECE")) : 4096;
    static const int hold_env = getenv("B200BLAS_CHOL_HOLD") ? atoi(getenv("B200BLAS_CHOL_HOLD")) : -1;
    const bool hold_updates = hold_env >= 0 ? hold_env != 0 : ndev >= 4;
    static const bool tracing = getenv("B200BLAS_MG_TRACE") != nullptr;
    std::vector<TraceItem> trace;
    const double trace_t0 = tracing ? now_ms() : 0.0;
    auto trace_ev = [&](cudaEvent_t ev, const char* what, int J, int d) {
        if (!tracing || !ev) return;
        TraceItem it; snprintf(it.label, sizeof it.label, "J=%d %s (dev %d)", J, what, d); it.ev = ev; it.t_ms = -1; trace.push_back(it);
    };

    // ---- per-device storage, streams ----
    struct CDev { double* W; double* P[2]; int* info; cudaStream_t panel, la; };     // la: look-ahead column updates (high priority)
    static CDev cd[kMaxDevices] = {};
    static size_t capW[kMaxDevices] = {}, capP[kMaxDevices] = {};
    cudaEvent_t start;
    { MgDev& h = st.dev[0]; h.next_event = 0; start = next_event(h); B200_CUDA(cudaEventRecord(start, home_stream)); }
    for (int d = 0; d < ndev; d++) {
        MgDev& md = st.dev[d];
        if (d) md.next_event = 0;
        DeviceScope scope(md.id);
        if (d) ws_reset();               // (the home context was reset by the entry point; its workspace may already hold a staged operand)
        const int ncols = (NB - d + ndev - 1) / ndev;                  // block columns owned by d
        if (d) {
            const size_t need = (size_t)ldw * (size_t)ncols * nb * 8;
            if (need > capW[d]) { if (cd[d].W) B200_CUDA(cudaFree(cd[d].W)); B200_CUDA(cudaMalloc((void**)&cd[d].W, need)); capW[d] = need; }
        }
        if (ndev > 1) {
            const size_t need = (size_t)ldw * nb * 8;
            if (need > capP[d]) {
                for (int b = 0; b < 2; b++) { if (cd[d].P[b]) B200_CUDA(cudaFree(cd[d].P[b])); B200_CUDA(cudaMalloc((void**)&cd[d].P[b], need)); }
                capP[d] = need;
            }
        }
        if (!cd[d].panel) {
            int lo = 0, hi = 0;
            B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            B200_CUDA(cudaStreamCreateWithPriority(&cd[d].panel, cudaStreamNonBlocking, hi));
            B200_CUDA(cudaStreamCreateWithPriority(&cd[d].la, cudaStreamNonBlocking, hi));
        }
        cd[d].info = (int*)((char*)device_scalar() + 64);
        B200_CUDA(cudaStreamWaitEvent(md.comp, start, 0));
        B200_CUDA(cudaStreamWaitEvent(md.out, start, 0));
        B200_CUDA(cudaStreamWaitEvent(cd[d].panel, start, 0));
        B200_CUDA(cudaStreamWaitEvent(cd[d].la, start, 0));
        for (int t = 0; t < ndev; t++) if (t != d) B200_CUDA(cudaStreamWaitEvent(fwd_stream(md, t), start, 0));
        B200_CUDA(cudaStreamWaitEvent(push_stream(md), start, 0));
        B200_CUDA(cudaMemsetAsync(cd[d].info, 0, sizeof(int), cd[d].panel));
    }
    auto colptr = [&](int d, int J) -> double* { return d == 0 ? a + (int64_t)J * nb * lda : cd[d].W + (int64_t)loc(J) * nb * ldw; };   // element (i, c): [i + c*ld]
    auto colld = [&](int d) -> int64_t { return d == 0 ? lda : ldw; };
    auto comp_stream = [&](int d) { return d == 0 ? home_stream : st.dev[d].comp; };

    // ---- distribution: the other devices' block columns (rows on and below the diagonal block) leave home in column order ----
    std::vector<cudaEvent_t> dist_ev(NB, nullptr), fact_ev(NB, nullptr), col_ready(NB, nullptr), diag_ready(NB, nullptr);
    for (int J = 0; J < NB; J++) {
        const int d = own(J);
        if (d == 0) continue;
        const int64_t j = (int64_t)J * nb; const int jb = (int)std::min<int64_t>(nb, n - j);
        cudaStream_t s = push_stream(st.dev[0]);       // one ordered stream: columns leave in the order the steps need them
        DeviceScope scope(st.dev[0].id);
        B200_CUDA(cudaMemcpy2DAsync(colptr(d, J) + j, (size_t)ldw * 8, a + j + j * lda, (size_t)lda * 8, (size_t)(n - j) * 8, (size_t)jb, cudaMemcpyDefault, s));
        dist_ev[J] = next_event(st.dev[0]);
        B200_CUDA(cudaEventRecord(dist_ev[J], s));
    }
    // used[d][J]: all of device d's updates with panel J are done (its landing buffer J % 2 may be overwritten by panel J + 2)
    std::vector<cudaEvent_t> used((size_t)ndev * NB, nullptr), arrived((size_t)ndev * NB, nullptr), first_arrived((size_t)ndev * NB, nullptr);
    // last_fwd[d][slot]: the last forward that READS device d's landing buffer `slot` (it must finish before the buffer is refilled)
    std::vector<cudaEvent_t> last_fwd((size_t)ndev * 2, nullptr);
    std::vector<cudaEvent_t> col_done(NB, nullptr);            // last update of column K so far (event on whichever stream ran it)


    auto factor = [&](int J) {
        const int o = own(J);
        const int64_t j = (int64_t)J * nb; const int jb = (int)std::min<int64_t>(nb, n - j); const int64_t rest = n - j - jb;
        MgDev& md = st.dev[o];
        DeviceScope scope(md.id);
        cudaStream_t ps = cd[o].panel;
        if (diag_ready[J]) B200_CUDA(cudaStreamWaitEvent(ps, diag_ready[J], 0));
        if (dist_ev[J]) B200_CUDA(cudaStreamWaitEvent(ps, dist_ev[J], 0));
        double* diag = colptr(o, J) + j;
        potrf_lower_dev(ps, jb, diag, colld(o), cd[o].info, (int)j);
        if (tracing) { cudaEvent_t e = next_event(md); B200_CUDA(cudaEventRecord(e, ps)); trace_ev(e, "potrf done", J, o); trace_ev(diag_ready[J], "diag ready", J, o); trace_ev(col_ready[J], "column ready", J, o); }
        if (col_ready[J]) B200_CUDA(cudaStreamWaitEvent(ps, col_ready[J], 0));
        if (rest > 0) trsm_dev<double>(ps, 'R', 'L', 'T', 'N', (int)rest, jb, 1.0, diag, colld(o), diag + jb, colld(o));
        fact_ev[J] = next_event(md);
        B200_CUDA(cudaEventRecord(fact_ev[J], ps));
        trace_ev(fact_ev[J], "panel solved", J, o);
        if (o != 0) {       // the finished column goes home under the remaining steps
            B200_CUDA(cudaStreamWaitEvent(md.out, fact_ev[J], 0));
            B200_CUDA(cudaMemcpy2DAsync(a + j + j * lda, (size_t)lda * 8, diag, (size_t)ldw * 8, (size_t)(n - j) * 8, (size_t)jb, cudaMemcpyDefault, md.out));
        }
    };

    factor(0);
    for (int J = 0; J < NB; J++) {
        const int o = own(J);
        const int64_t j = (int64_t)J * nb; const int jb = (int)std::min<int64_t>(nb, n - j); const int64_t rest = n - j - jb;
        if (rest <= 0) break;
        const int slot = J & 1;
        // ---- ring broadcast of panel J (rows below the diagonal block) to the devices that still own columns > J ----
        int nrecv = std::min(ndev - 1, NB - 1 - J);                   // owners of columns J+1 .. J+nrecv, in ring order
        for (int64_t r0 = j + jb; r0 < n && nrecv > 0; r0 += piece_rows) {
            const int64_t rows = std::min<int64_t>(piece_rows, n - r0);
            const bool last_piece = r0 + piece_rows >= n;
            int prev = o;
            cudaEvent_t prev_ev = fact_ev[J];
            for (int h = 1; h <= nrecv; h++) {
                const int d = (o + h) % ndev;
                MgDev& ex = st.dev[prev];
                DeviceScope scope(ex.id);
                cudaStream_t s = fwd_stream(ex, d);
                B200_CUDA(cudaStreamWaitEvent(s, prev_ev, 0));
                if (J >= 2 && used[(size_t)d * NB + (J - 2)]) B200_CUDA(cudaStreamWaitEvent(s, used[(size_t)d * NB + (J - 2)], 0));
                if (last_fwd[(size_t)d * 2 + slot]) B200_CUDA(cudaStreamWaitEvent(s, last_fwd[(size_t)d * 2 + slot], 0));
                const double* src = prev == o ? colptr(o, J) + r0 : cd[prev].P[slot] + r0;
                const int64_t sld = prev == o ? colld(o) : ldw;
                B200_CUDA(cudaMemcpy2DAsync(cd[d].P[slot] + r0, (size_t)ldw * 8, src, (size_t)sld * 8, (size_t)rows * 8, (size_t)jb, cudaMemcpyDefault, s));
                cudaEvent_t ev = next_event(ex);
                B200_CUDA(cudaEventRecord(ev, s));
                if (last_piece) arrived[(size_t)d * NB + J] = ev;      // pieces of one (source, destination) pair share a stream: the last implies all
                if (r0 == j + jb) first_arrived[(size_t)d * NB + J] = ev;   // the first piece holds the rows of the next diagonal block
                if (last_piece && h == 1) trace_ev(ev, "panel at next owner", J, d);
                if (last_piece && h == nrecv) trace_ev(ev, "panel at last device", J, d);
                if (last_piece && prev != o) last_fwd[(size_t)prev * 2 + slot] = ev;
                prev = d; prev_ev = ev;
            }
        }
        // ---- rank-nb update of every owned column K > J ----
        // The next panel's column (K = J + 1) is updated on its owner's HIGH-PRIORITY look-ahead stream, not behind that device's
        // other updates: the compute stream is in order, and at step J it still holds the device's step J-1 updates of its other
        // columns (measured at N = 8: 2.5 ms per early step instead of ~1.4, profiles/r02k_chol8_trace.txt).  col_done[K] orders the
        // successive updates of one column across the two streams.
        for (int h = 0; h < ndev; h++) {
            const int d = (o + 1 + h) % ndev;                          // next owner first: its look-ahead work is queued earliest
            MgDev& md = st.dev[d];
            DeviceScope scope(md.id);
            cudaStream_t cs = comp_stream(d);
            bool any = false, comp_has_panel = false;
            cudaEvent_t panel_here = d == o ? fact_ev[J] : arrived[(size_t)d * NB + J];
            for (int K = J + 1; K < NB; K++) {
                if (own(K) != d) continue;
                const int64_t kcol = (int64_t)K * nb; const int kb = (int)std::min<int64_t>(nb, n - kcol);
                const double* Pn = d == o ? colptr(o, J) : cd[d].P[slot];     // panel J, addressed by global row
                const int64_t pld = d == o ? colld(o) : ldw;
                double* Ck = colptr(d, K) + kcol;
                if (K == J + 1) {
                    cudaStream_t ls = cd[d].la;
                    // its diagonal block needs only the first piece of the panel (nb <= piece rows) ...
                    const bool early = d != o && nb <= piece_rows;
                    B200_CUDA(cudaStreamWaitEvent(ls, early ? first_arrived[(size_t)d * NB + J] : panel_here, 0));
                    if (col_done[K]) B200_CUDA(cudaStreamWaitEvent(ls, col_done[K], 0));
                    if (J == 0 && dist_ev[K]) B200_CUDA(cudaStreamWaitEvent(ls, dist_ev[K], 0));
                    // ... and goes first, so that its factorisation (panel stream) runs underneath the update of the rows below it
                    dgemm_dev(ls, 'N', 'T', kb, kb, jb, -1.0, Pn + kcol, pld, Pn + kcol, pld, 1.0, Ck, colld(d), MASK_LOWER);
                    diag_ready[K] = next_event(md);
                    B200_CUDA(cudaEventRecord(diag_ready[K], ls));
                    if (early) B200_CUDA(cudaStreamWaitEvent(ls, panel_here, 0));                        // the rows below need the whole panel
                    const int64_t below = n - kcol - kb;
                    if (below > 0) dgemm_dev(ls, 'N', 'T', (int)below, kb, jb, -1.0, Pn + kcol + kb, pld, Pn + kcol, pld, 1.0, Ck + kb, colld(d), MASK_FULL);
                    col_ready[K] = next_event(md);
                    B200_CUDA(cudaEventRecord(col_ready[K], ls));
                    col_done[K] = col_ready[K];
                    factor(K);                                         // on this device's high-priority panel stream
                    // With many devices the panel chain (diagonal block, solve, ring) is the critical path and this device's
                    // other updates are a small share of the step: they wait until the panel is solved, so its ~50 small
                    // dependent kernels do not queue behind update CTAs (the DGEMM tile owns a whole SM; measured at N = 4:
                    // potrf(512) 1.25 ms under the update against 0.59 ms alone, profiles/r02h_chol4_trace.txt).
                    // (the wait also makes `used` below cover the look-ahead stream's reads of the panel, and gives the compute stream the panel)
                    B200_CUDA(cudaStreamWaitEvent(cs, hold_updates ? fact_ev[K] : col_ready[K], 0));
                    any = comp_has_panel = true;
                    continue;
                }
                if (!comp_has_panel) { B200_CUDA(cudaStreamWaitEvent(cs, panel_here, 0)); comp_has_panel = true; }
                any = true;
                if (J == 0 && dist_ev[K]) B200_CUDA(cudaStreamWaitEvent(cs, dist_ev[K], 0));
                dgemm_dev(cs, 'N', 'T', (int)(n - kcol), kb, jb, -1.0, Pn + kcol, pld, Pn + kcol, pld, 1.0, Ck, colld(d), MASK_LOWER);
                // the column that becomes the look-ahead column of the NEXT step: mark this, its last update on the compute stream
                if (K == J + 2) { col_done[K] = next_event(md); B200_CUDA(cudaEventRecord(col_done[K], cs)); }
            }
            if (any) {
                cudaEvent_t ev = next_event(md);
                B200_CUDA(cudaEventRecord(ev, cs));
                used[(size_t)d * NB + J] = ev;
            }
        }
    }
    // ---- completion: the caller's stream waits for every stream of every device; then the info words ----
    int* pin = (int*)pinned_scalar();
    for (int d = 0; d < ndev; d++) {
        MgDev& md = st.dev[d];
        DeviceScope scope(md.id);
        cudaStream_t cs = comp_stream(d);
        auto join = [&](cudaStream_t s) { if (!s || s == cs) return; cudaEvent_t e = next_event(md); B200_CUDA(cudaEventRecord(e, s)); B200_CUDA(cudaStreamWaitEvent(cs, e, 0)); };
        join(cd[d].panel); join(cd[d].la); join(md.out); join(md.push);
        for (int t = 0; t < ndev; t++) if (t != d) join(md.fwd[t]);
        B200_CUDA(cudaMemcpyAsync(pin + d, cd[d].info, sizeof(int), cudaMemcpyDeviceToHost, cs));
        if (d) { B200_CUDA(cudaEventRecord(md.done, cs)); }
    }
    for (int d = 1; d < ndev; d++) B200_CUDA(cudaStreamWaitEvent(home_stream, st.dev[d].done, 0));
    if (tracing) { b200_writef(STDERR_FILENO, "mgtrace cholesky enqueue took %.3f ms\n", now_ms() - trace_t0); trace_poll(trace, trace_t0); }
    B200_CUDA(cudaStreamSynchronize(home_stream));
    int info = 0;
    for (int d = 0; d < ndev; d++) if (pin[d] > 0 && (info == 0 || pin[d] < info)) info = pin[d];
    return info;
}

template bool multi_gemm<float>(char, char, int, int, int, float, const float*, int64_t, const float*, int64_t, float, float*, int64_t);
template bool multi_gemm<double>(char, char, int, int, int, double, const double*, int64_t, const double*, int64_t, double, double*, int64_t);
template bool multi_gemm<cuFloatComplex>(char, char, int, int, int, cuFloatComplex, const cuFloatComplex*, int64_t, const cuFloatComplex*, int64_t, cuFloatComplex, cuFloatComplex*, int64_t);
template bool multi_gemm<cuDoubleComplex>(char, char, int, int, int, cuDoubleComplex, const cuDoubleComplex*, int64_t, const cuDoubleComplex*, int64_t, cuDoubleComplex, cuDoubleComplex*, int64_t);

}  // namespace b200
