// stager_mock.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// libgpublas_b200/csrc/host_stager.cu (pageable host memory <-> device through a ring of pinned slots packed by host threads) is
// pure host code; here it is compiled with g++ against a stand-in for the seven CUDA runtime calls it makes.  The stand-in stream
// is DEFERRED: cudaMemcpy[2D]Async only queues the copy, it executes when an event recorded after it is synchronised (or at the
// final drain) -- so a slot that is re-packed before its transfer's event has been waited for corrupts the data, exactly as on
// the device.  "Device memory" is ordinary host memory.  The driver pushes matrices and vectors of awkward shapes to the
// "device" and back and compares.  tests/test_host_stager_cpu.py builds and runs it (with AddressSanitizer); libb200blas.so never
// sees this file.
#include <cuda_runtime_api.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <vector>

#include "../../libgpublas_b200/csrc/tracker.h"

namespace {
struct Op { std::function<void()> fn; void* event; };
std::deque<Op> g_queue;          // one stream is enough: the stager is handed one
std::mutex g_qmu;
int g_event_ids = 0;
void drain_until(void* ev) {
    std::lock_guard<std::mutex> lock(g_qmu);
    bool present = false;
    for (auto& o : g_queue) if (o.event == ev) present = true;
    if (ev && !present) return;                       // already completed
    while (!g_queue.empty()) {
        Op o = std::move(g_queue.front());
        g_queue.pop_front();
        if (o.fn) o.fn();
        if (ev && o.event == ev) {
            bool again = false;                       // the event may have been re-recorded later: stop at its LAST record? no -- at the first
            (void)again;
            return;
        }
    }
}
}  // namespace

extern "C" {
cudaError_t cudaMallocHost(void** p, size_t bytes) { *p = malloc(bytes); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)(intptr_t)(++g_event_ids); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    std::lock_guard<std::mutex> lock(g_qmu);
    // a re-recorded event completes at its newest record: drop the older marker
    for (auto& o : g_queue) if (o.event == (void*)e) o.event = nullptr;
    g_queue.push_back(Op{nullptr, (void*)e});
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e) { drain_until((void*)e); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::lock_guard<std::mutex> lock(g_qmu);
    g_queue.push_back(Op{[=] { memcpy(dst, src, n); }, nullptr});
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t) {
    std::lock_guard<std::mutex> lock(g_qmu);
    g_queue.push_back(Op{[=] { for (size_t r = 0; r < height; r++) memcpy((char*)dst + r * dpitch, (const char*)src + r * spitch, width); }, nullptr});
    return cudaSuccess;
}
const char* cudaGetErrorString(cudaError_t) { return "mock"; }
void tracker_enter(void) {}
void tracker_leave(void) {}
}
namespace b200 {
void fatal(const char* what, const char* file, int line, const char* detail) { fprintf(stderr, "fatal %s %s:%d %s\n", what, file, line, detail); abort(); }
}

#include "../../libgpublas_b200/csrc/host_stager.cu"

static unsigned long long g_seed = 88172645463325252ull;
static unsigned char rnd() { g_seed ^= g_seed << 13; g_seed ^= g_seed >> 7; g_seed ^= g_seed << 17; return (unsigned char)(g_seed >> 32); }

// host matrix (spitch) -> "device" (dpitch) -> another host matrix (spitch2); padding bytes must keep their rogue value
static int round_trip(size_t width, size_t height, size_t spitch, size_t dpitch, const char* tag) {
    std::vector<unsigned char> h0(spitch * height), dev(dpitch * height, 0xAB), h1(spitch * height, 0xCD);
    for (auto& b : h0) b = rnd();
    b200::staged_copy_2d(dev.data(), dpitch, h0.data(), spitch, width, height, true, nullptr);
    // as the library does: the caller's stream is synchronised (finish_call) before results are used
    drain_until(nullptr);
    for (size_t r = 0; r < height; r++) {
        if (memcmp(dev.data() + r * dpitch, h0.data() + r * spitch, width)) { printf("FAIL %s: H2D row %zu differs\n", tag, r); return 1; }
        for (size_t c = width; c < dpitch && r * dpitch + c < dev.size(); c++) if (dev[r * dpitch + c] != 0xAB) { printf("FAIL %s: H2D padding of row %zu written\n", tag, r); return 1; }
    }
    b200::staged_copy_2d(h1.data(), spitch, dev.data(), dpitch, width, height, false, nullptr);      // returns with the data on the host
    for (size_t r = 0; r < height; r++) {
        if (memcmp(h1.data() + r * spitch, h0.data() + r * spitch, width)) { printf("FAIL %s: D2H row %zu differs\n", tag, r); return 1; }
        for (size_t c = width; c < spitch && r * spitch + c < h1.size(); c++) if (h1[r * spitch + c] != 0xCD) { printf("FAIL %s: D2H padding of row %zu written\n", tag, r); return 1; }
    }
    return 0;
}

int main() {
    int bad = 0, cases = 0;
    const size_t MiB = (size_t)1 << 20;
    struct { size_t w, h, sp, dp; const char* tag; } shapes[] = {
        {8 * 3000, 2500, 8 * 3001, 8 * 3000, "3000 x 2500 doubles inside lda 3001: three slots"},
        {8 * 1, 700001, 16, 8, "strided rows of one element"},
        {40 * MiB + 24, 1, 40 * MiB + 24, 40 * MiB + 24, "flat vector larger than a slot"},
        {33 * MiB, 5, 33 * MiB + 64, 33 * MiB, "five rows, each larger than a slot: the slots take turns"},
        {32 * MiB, 3, 32 * MiB, 32 * MiB + 128, "rows of exactly one slot"},
        {12 * MiB + 8, 23, 12 * MiB + 40, 12 * MiB + 8, "two rows per slot, odd count"},
        {4096, 40000, 4096, 8192, "many short rows: 8192 rows per slot, five pieces"},
        {17, 3, 64, 32, "tiny"},
        {5 * MiB, 1, 5 * MiB, 5 * MiB, "one row below a slot"},
    };
    for (auto& s : shapes) { cases++; bad += round_trip(s.w, s.h, s.sp, s.dp, s.tag); }
    printf("RESULT cases=%d bad=%d\n", cases, bad);
    return bad != 0;
}
