// tracker_mock.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// libgpublas_b200/csrc/tracker.cpp (the allocator interposers + the registry of managed blocks) linked against a stand-in for
// the four CUDA runtime calls it makes, so that its MANAGED path -- registry insert / lookup / remove, realloc of tracked blocks,
// the aligned allocators, blocks freed on another thread -- runs under LD_PRELOAD on a machine without a GPU:
//   cudaMallocManaged  -> an anonymous mmap with the size kept in a leading page (page-aligned bases, like small managed blocks)
//   cudaFree           -> munmap
// tests/test_preload.py builds this file + tracker.cpp into tests/drivers/_build/libtracker_mock.so and preloads it into the
// allocator drivers.  The real library's behaviour on a device is what the -m gpu tests check; libb200blas.so never sees this file.
#include <cuda_runtime_api.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "../../libgpublas_b200/csrc/tracker.h"

static const size_t kPage = 4096;

extern "C" {
__attribute__((visibility("default"))) cudaError_t cudaMallocManaged(void** p, size_t size, unsigned int) {
    void* base = mmap(nullptr, size + kPage, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return cudaErrorMemoryAllocation;
    *(size_t*)base = size;
    *p = (char*)base + kPage;
    return cudaSuccess;
}
__attribute__((visibility("default"))) cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    char* base = (char*)p - kPage;
    munmap(base, *(size_t*)base + kPage);
    return cudaSuccess;
}
__attribute__((visibility("default"))) cudaError_t cudaGetLastError(void) { return cudaSuccess; }
__attribute__((visibility("default"))) const char* cudaGetErrorString(cudaError_t) { return "mock"; }

void b200_writef(int fd, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n > 0) syscall(SYS_write, fd, buf, (size_t)(n < (int)sizeof buf ? n : (int)sizeof buf - 1));
}
__attribute__((visibility("default"))) int b200blas_is_tracked(const void* p) { return tracker_lookup(p, nullptr, nullptr); }
__attribute__((visibility("default"))) void b200blas_print_help(void) {}
}

// Device bring-up stand-in.  TRACKER_MOCK_INIT_MS=<ms>: the "device" is not ready at start; the first qualifying allocation
// brings it up, which takes <ms> milliseconds during which the bring-up (like the CUDA driver) creates a helper thread that
// allocates large blocks of its own -- those must come from glibc, while application threads that ask for qualifying blocks
// meanwhile must end up with managed ones (the window the round-1 GPU run fell into).  TRACKER_MOCK_NO_DEVICE=1: bring-up fails.
#include <mutex>
#include <pthread.h>
static bool g_mock_ready = true;
static int g_helper_tracked = -1, g_helper_child_tracked = -1;
static void* mock_helper_child(void*) {
    void* p = malloc(1 << 20);
    g_helper_child_tracked = tracker_lookup(p, nullptr, nullptr);
    free(p);
    return nullptr;
}
static void* mock_helper(void* arg) {
    const long ms = (long)(intptr_t)arg;
    int seen = 0;
    for (long t = 0; t < ms; t += 10) {               // a driver worker allocating while the bring-up is in progress ...
        void* p = malloc(256 << 10);
        seen |= tracker_lookup(p, nullptr, nullptr);
        free(p);
        usleep(10000);
    }
    pthread_t th;                                       // ... and spawning a worker of its own
    pthread_create(&th, nullptr, mock_helper_child, nullptr);
    pthread_join(th, nullptr);
    g_helper_tracked = seen;
    return nullptr;
}
static void mock_bring_up() {
    if (getenv("TRACKER_MOCK_NO_DEVICE")) return;
    const char* ms = getenv("TRACKER_MOCK_INIT_MS");
    pthread_t th;
    pthread_create(&th, nullptr, mock_helper, (void*)(intptr_t)(ms ? atol(ms) : 0));
    void* own = malloc(1 << 20);                        // the initialising thread's own allocations stay on the heap too
    if (tracker_lookup(own, nullptr, nullptr)) g_helper_tracked = 1;
    free(own);
    pthread_join(th, nullptr);
    g_mock_ready = true;
}
namespace b200 {
static std::once_flag g_mock_once;
bool device_ready() { return g_mock_ready; }
bool try_init() { std::call_once(g_mock_once, mock_bring_up); return g_mock_ready; }
void ensure_init() { try_init(); }
}  // namespace b200

__attribute__((constructor)) static void mock_ctor() {
    const char* h = getenv("TRACKER_MOCK_HEURISTIC");
    if (h && !strcmp(h, "true")) tracker_set_heuristic(B200_H_TRUE);
    if (h && !strcmp(h, "false")) tracker_set_heuristic(B200_H_FALSE);
    if (getenv("TRACKER_MOCK_INIT_MS") || getenv("TRACKER_MOCK_NO_DEVICE")) g_mock_ready = false;
    tracker_set_tracking(1);
}
__attribute__((destructor)) static void mock_dtor() {
    tracker_set_shutdown();
    if (getenv("TRACKER_MOCK_INIT_MS"))
        b200_writef(STDOUT_FILENO, "MOCK helper_tracked=%d helper_child_tracked=%d\n", g_helper_tracked, g_helper_child_tracked);
}
