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

namespace b200 {
bool device_ready() { return true; }
void ensure_init() {}
}  // namespace b200

__attribute__((constructor)) static void mock_ctor() {
    const char* h = getenv("TRACKER_MOCK_HEURISTIC");
    if (h && !strcmp(h, "true")) tracker_set_heuristic(B200_H_TRUE);
    if (h && !strcmp(h, "false")) tracker_set_heuristic(B200_H_FALSE);
    tracker_set_tracking(1);
}
__attribute__((destructor)) static void mock_dtor() { tracker_set_shutdown(); }
