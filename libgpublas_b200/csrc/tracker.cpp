// tracker.cpp -- (host-only translation unit: nvcc refuses to redefine malloc/free)
// see tracker.h.  Differences from the reference, all on purpose:
//  * the real allocator is reached through glibc's exported __libc_* entry points, not
//    dlsym(RTLD_NEXT) at first use (obj_tracker.c:142-223), so there is no bootstrap recursion;
//  * managed blocks carry no in-band header (the reference stores the size in an 8-byte header and
//    returns base+8, blas2cuda.c:127-148, which breaks 16-byte alignment and therefore TMA):
//    sizes live in the registry and the caller gets the cudaMallocManaged base itself;
//  * the registry is a sorted array under a rwlock with an address-range pre-filter, so free() of
//    ordinary heap pointers never takes the lock (the reference does a tsearch under a rwlock for
//    every free, obj_tracker.c:948-982);
//  * the default heuristic is size-based, not random (obj_tracker.c:52);
//  * posix_memalign / aligned_alloc / memalign / valloc / malloc_usable_size are interposed too (the reference leaves them to
//    glibc, so aligned operands are never tracked).
#include "tracker.h"
#include "runtime.h"
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <errno.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/syscall.h>
#include <time.h>

extern "C" {
void* __libc_malloc(size_t);
void* __libc_calloc(size_t, size_t);
void* __libc_realloc(void*, size_t);
void __libc_free(void*);
void* __libc_memalign(size_t, size_t);
}

namespace {

struct Block { uintptr_t base; size_t size; int resident; /* bulk-migrated to the device since it was allocated */ uint64_t nth, uid; };

pthread_rwlock_t g_lock = PTHREAD_RWLOCK_INITIALIZER;
Block* g_blocks = nullptr;
size_t g_nblocks = 0, g_cap = 0;
volatile uintptr_t g_lo = UINTPTR_MAX, g_hi = 0;   // address envelope of all blocks ever handed out

// Per-thread re-entrancy depth WITHOUT thread-local storage.  A TLS variable read inside malloc must be
// initial-exec (a general-dynamic access may itself call malloc to allocate the module's TLS block), but one
// initial-exec variable marks the whole shared object DF_STATIC_TLS, and the statically linked CUDA runtime
// carries a 4 KiB-aligned TLS block that does not fit glibc's static-TLS surplus -- dlopen() of the library
// (ctypes, Octave, Python extensions) would fail.  So the counter lives in a fixed open-addressing table keyed
// by the thread pointer (%fs:0, the TCB address: unique per live thread, readable without any allocation).
struct ThreadSlot { volatile uintptr_t tp; int inside; };
constexpr size_t kSlots = 4096;                 // live threads beyond this are simply never tracked
ThreadSlot g_slots[kSlots];
int g_overflow_inside = 1 << 30;   // shared by every thread beyond the table: large, so racing ++/-- never reach zero (such threads are never tracked)

inline int& inside_ref() {
    const uintptr_t tp = (uintptr_t)__builtin_thread_pointer();
    size_t h = (size_t)(((tp >> 6) * 0x9E3779B97F4A7C15ull) >> 52);
    for (size_t probe = 0; probe < kSlots; probe++, h = (h + 1) & (kSlots - 1)) {
        uintptr_t cur = g_slots[h].tp;
        if (cur == tp) return g_slots[h].inside;
        if (cur == 0) {
            if (__sync_bool_compare_and_swap(&g_slots[h].tp, (uintptr_t)0, tp)) return g_slots[h].inside;
            if (g_slots[h].tp == tp) return g_slots[h].inside;
        }
    }
    return g_overflow_inside;
}
#define t_inside (inside_ref())
volatile int g_tracking = 0;
volatile int g_heuristic = B200_H_SIZE;
volatile size_t g_threshold = 64 * 1024;
volatile int g_shutdown = 0;      // set at exit, before the CUDA runtime tears itself down
uint64_t g_nth = 0;
b200_tracker_stats g_tstats = {0, 0, 0, 0, 0};

uint8_t* g_decisions = nullptr;   // oracle bitmap: bit n set => n-th allocation is managed
size_t g_ndecisions = 0;

// index of the block containing p, or -1.  caller holds the lock.
long find_block(uintptr_t p) {
    size_t lo = 0, hi = g_nblocks;   // half-open; (the reference's in_excluded_region search has an
    while (lo < hi) {                //  off-by-one on exactly this, obj_tracker.c:409-424)
        size_t mid = lo + (hi - lo) / 2;
        if (p < g_blocks[mid].base) hi = mid;
        else if (p >= g_blocks[mid].base + g_blocks[mid].size) lo = mid + 1;
        else return (long)mid;
    }
    return -1;
}

volatile int g_trace = 0;          // BLAS2CUDA_OPTIONS=trace: T/U/C lines in the reference's TRACE_OUTPUT format
uint64_t g_uid = 0;

// reference obj_tracker_print_info (lib/obj_tracker.c:426-483): "<T|U|C> #nth [ptr] fun=[..] reqsize=[..] tid=[..] time=[..s+..ns] uid=[..]"
void trace_line(char kind, uint64_t nth, const void* ptr, const char* fun, size_t reqsize, uint64_t uid) {
    struct timespec tm;
    clock_gettime(CLOCK_MONOTONIC_RAW, &tm);
    b200_writef(STDOUT_FILENO, "%c #%lu [%p] fun=[%s] reqsize=[%zu] tid=[%d] time=[%lds+%ldns] uid=[%lu]\n", kind, (unsigned long)nth, ptr, fun,
                reqsize, (int)syscall(SYS_gettid), (long)tm.tv_sec, (long)tm.tv_nsec, (unsigned long)uid);
}

bool registry_insert(uintptr_t base, size_t size, uint64_t nth, const char* fun) {
    pthread_rwlock_wrlock(&g_lock);
    if (g_nblocks == g_cap) {
        size_t ncap = g_cap ? g_cap * 2 : 256;
        Block* nb = (Block*)__libc_realloc(g_blocks, ncap * sizeof(Block));
        if (!nb) { pthread_rwlock_unlock(&g_lock); return false; }
        g_blocks = nb; g_cap = ncap;
    }
    size_t pos = g_nblocks;
    while (pos > 0 && g_blocks[pos - 1].base > base) { g_blocks[pos] = g_blocks[pos - 1]; pos--; }
    g_blocks[pos].base = base; g_blocks[pos].size = size; g_blocks[pos].resident = 0;
    g_blocks[pos].nth = nth; g_blocks[pos].uid = ++g_uid;
    const uint64_t uid = g_blocks[pos].uid;
    g_nblocks++;
    if (base < g_lo) g_lo = base;
    if (base + size > g_hi) g_hi = base + size;
    g_tstats.managed_allocs++;
    g_tstats.managed_bytes_live += size;
    if (g_tstats.managed_bytes_live > g_tstats.managed_bytes_peak) g_tstats.managed_bytes_peak = g_tstats.managed_bytes_live;
    pthread_rwlock_unlock(&g_lock);
    if (g_trace) trace_line('T', nth, (void*)base, fun, size, uid);
    return true;
}

// removes the block whose BASE is p; returns its size or 0
size_t registry_remove(uintptr_t p) {
    size_t sz = 0;
    pthread_rwlock_wrlock(&g_lock);
    long i = find_block(p);
    if (i >= 0 && g_blocks[i].base == p) {
        sz = g_blocks[i].size;
        if (g_trace) trace_line('U', g_blocks[i].nth, (void*)p, "free", sz, g_blocks[i].uid);
        memmove(&g_blocks[i], &g_blocks[i + 1], (g_nblocks - i - 1) * sizeof(Block));
        g_nblocks--;
        g_tstats.managed_frees++;
        g_tstats.managed_bytes_live -= sz;
    }
    pthread_rwlock_unlock(&g_lock);
    return sz;
}

bool should_manage(size_t request, uint64_t nth) {
    switch (g_heuristic) {
        case B200_H_TRUE: return true;
        case B200_H_FALSE: return false;
        case B200_H_RANDOM: return (random() & 1) != 0;
        case B200_H_ORACLE: return nth < g_ndecisions && (g_decisions[nth / 8] & (1u << (nth % 8)));
        default: return request >= g_threshold;
    }
}

void* managed_new(size_t request, uint64_t nth = 0, const char* fun = "malloc") {
    // The first qualifying allocation brings the device up.  Allocations CUDA itself makes meanwhile -- on this thread
    // (t_inside) or on the helper threads it creates (marked inside for life by the pthread_create hook below) -- go to
    // glibc.  Every OTHER application thread that asks for a qualifying block during the bring-up blocks here until the
    // device is ready and is then served from managed memory like any later request: once tracking is on, every
    // qualifying allocation is tracked (reference obj_tracker.c:789-840; the reference gets there by initialising in
    // its constructor before tracking is enabled, blas2cuda.c:178-242).  If the device cannot come up at all, tracking
    // is switched off and the heap serves the process -- only a BLAS call is fatal without a device.
    if (!b200::device_ready()) {
        t_inside++;
        const bool ok = b200::try_init();
        t_inside--;
        if (!ok) { g_tracking = 0; return nullptr; }
    }
    void* p = nullptr;
    t_inside++;
    cudaError_t e = cudaMallocManaged(&p, request ? request : 1, cudaMemAttachGlobal);
    t_inside--;
    if (e != cudaSuccess) {
        cudaGetLastError();
        b200_writef(STDERR_FILENO, "b200blas: cudaMallocManaged(%zu) failed: %s -- falling back to the heap\n", request,
                    cudaGetErrorString(e));
        return nullptr;
    }
    if (!registry_insert((uintptr_t)p, request ? request : 1, nth, fun)) {
        t_inside++; cudaFree(p); t_inside--;
        return nullptr;
    }
    return p;
}

bool bypass() { return t_inside || !g_tracking; }

}  // namespace

extern "C" {

int tracker_lookup(const void* ptr, void** base, size_t* size) {
    uintptr_t p = (uintptr_t)ptr;
    if (p < g_lo || p >= g_hi) return 0;
    int found = 0;
    pthread_rwlock_rdlock(&g_lock);
    long i = find_block(p);
    if (i >= 0) {
        found = 1;
        if (base) *base = (void*)g_blocks[i].base;
        if (size) *size = g_blocks[i].size;
    }
    pthread_rwlock_unlock(&g_lock);
    return found;
}
int tracker_test_and_set_resident(const void* ptr) {
    uintptr_t p = (uintptr_t)ptr;
    if (p < g_lo || p >= g_hi) return -1;
    int prev = -1;
    pthread_rwlock_rdlock(&g_lock);
    long i = find_block(p);
    if (i >= 0) prev = __sync_lock_test_and_set(&g_blocks[i].resident, 1);
    pthread_rwlock_unlock(&g_lock);
    return prev;
}
int tracker_peek_resident(const void* ptr) {
    uintptr_t p = (uintptr_t)ptr;
    if (p < g_lo || p >= g_hi) return -1;
    int cur = -1;
    pthread_rwlock_rdlock(&g_lock);
    long i = find_block(p);
    if (i >= 0) cur = __atomic_load_n(&g_blocks[i].resident, __ATOMIC_RELAXED);
    pthread_rwlock_unlock(&g_lock);
    return cur;
}
void tracker_set_trace(int on) { g_trace = on; }
// "C" line for a tracked operand of a BLAS call (reference OBJPRINT_CALL): which allocation the routine `fun` used
void tracker_trace_call(const void* ptr, const char* fun) {
    if (!g_trace) return;
    uintptr_t p = (uintptr_t)ptr;
    if (p < g_lo || p >= g_hi) return;
    pthread_rwlock_rdlock(&g_lock);
    long i = find_block(p);
    uint64_t nth = 0, uid = 0; size_t sz = 0; uintptr_t base = 0;
    if (i >= 0) { nth = g_blocks[i].nth; uid = g_blocks[i].uid; sz = g_blocks[i].size; base = g_blocks[i].base; }
    pthread_rwlock_unlock(&g_lock);
    if (i >= 0) trace_line('C', nth, (void*)base, fun, sz, uid);
}
// what the allocator would decide for the nth allocation of `request` bytes (tests; CPU-only, touches no device)
int tracker_decision(uint64_t nth, size_t request) { return should_manage(request, nth) ? 1 : 0; }
void tracker_enter(void) { t_inside++; }
void tracker_leave(void) { t_inside--; }
void tracker_set_tracking(int on) { g_tracking = on; }
int tracker_get_tracking(void) { return g_tracking; }
void tracker_set_heuristic(int h) { g_heuristic = h; }
void tracker_set_threshold(size_t bytes) { g_threshold = bytes; }
void tracker_get_stats(struct b200_tracker_stats* out) {
    pthread_rwlock_rdlock(&g_lock);
    *out = g_tstats;
    out->allocs_seen = g_nth;
    pthread_rwlock_unlock(&g_lock);
}

int tracker_load_oracle_file(const char* filename) {
    // trace line format "H #<nth> ..." / "D #<nth> ..." (reference oracle.c:41); D = place on device
    t_inside++;
    FILE* f = fopen(filename, "r");
    if (!f) { t_inside--; return 0; }
    char line[4096];
    size_t cap = 0;
    uint8_t* bits = nullptr;
    size_t maxn = 0;
    while (fgets(line, sizeof line, f)) {
        char c; unsigned long long nth;
        if (sscanf(line, " %c #%llu", &c, &nth) != 2 || (c != 'H' && c != 'D')) continue;
        if (nth / 8 >= cap) {
            size_t ncap = cap ? cap : 1024;
            while (nth / 8 >= ncap) ncap *= 2;
            uint8_t* nb = (uint8_t*)__libc_realloc(bits, ncap);
            if (!nb) break;
            memset(nb + cap, 0, ncap - cap);
            bits = nb; cap = ncap;
        }
        if (c == 'D') bits[nth / 8] |= (uint8_t)(1u << (nth % 8));
        if (nth + 1 > maxn) maxn = nth + 1;
    }
    fclose(f);
    g_decisions = bits; g_ndecisions = maxn;
    t_inside--;
    return 1;
}

void* tracker_alloc_managed(size_t bytes) { return managed_new(bytes); }
int tracker_free_managed(void* p) {
    if (!registry_remove((uintptr_t)p)) return 0;
    // Once the process is exiting, blocks are dropped rather than returned: exit-time destructors (including the
    // CUDA runtime's own, which run after its context is gone) still call free() on tracked blocks, and re-entering
    // a half-destroyed runtime crashes.  The driver reclaims everything at process end.
    if (g_shutdown) return 1;
    t_inside++;
    cudaFree(p);
    t_inside--;
    return 1;
}
void tracker_set_shutdown(void) { g_shutdown = 1; g_tracking = 0; }

// ---- threads created by CUDA itself never get managed memory ----
// The reference keeps CUDA's own allocations out of the managed allocator by return address ("excluded
// regions" scanned from /proc/self/maps, obj_tracker.c:352-424), which misses allocations CUDA makes through
// libc helpers.  Here every thread that is created BY a thread that is inside the library (i.e. by the CUDA
// runtime / driver during one of our calls, or by one of its own helper threads) is marked "inside" for its whole life, so a driver worker
// thread can never re-enter cudaMallocManaged from malloc and deadlock on the driver's own locks.
// (A TCB address is reused once its thread has exited, so every new thread re-initialises its slot.)
struct ThreadStart { void* (*fn)(void*); void* arg; int depth; };
static void* thread_trampoline(void* p) {
    ThreadStart ts = *(ThreadStart*)p;
    __libc_free(p);
    t_inside = ts.depth;
    return ts.fn(ts.arg);
}
typedef int (*pthread_create_t)(pthread_t*, const pthread_attr_t*, void* (*)(void*), void*);
__attribute__((visibility("default"))) int pthread_create(pthread_t* thread, const pthread_attr_t* attr, void* (*fn)(void*), void* arg) {
    static pthread_create_t real = nullptr;
    if (!real) {
        t_inside++;
        real = (pthread_create_t)dlsym(RTLD_NEXT, "pthread_create");
        t_inside--;
        if (!real) { b200_writef(STDERR_FILENO, "b200blas: cannot resolve pthread_create\n"); abort(); }
    }
    ThreadStart* ts = (ThreadStart*)__libc_malloc(sizeof(ThreadStart));
    if (!ts) return real(thread, attr, fn, arg);
    ts->fn = fn; ts->arg = arg;
    ts->depth = t_inside > 0 ? (1 << 20) : 0;   // the CREATOR is inside the library: a CUDA helper thread.  Application threads start at 0.
    int rc = real(thread, attr, thread_trampoline, ts);
    if (rc != 0) __libc_free(ts);
    return rc;
}

// ---- the interposed allocator symbols (reference obj_tracker.c:789,842,902,948) ----
void* malloc(size_t request) noexcept {
    if (bypass()) return __libc_malloc(request);
    uint64_t nth = __sync_fetch_and_add(&g_nth, 1);
    if (request && should_manage(request, nth)) {
        void* p = managed_new(request, nth, "malloc");
        if (p) return p;
    }
    return __libc_malloc(request);
}

void* calloc(size_t nmemb, size_t size) noexcept {
    if (bypass()) return __libc_calloc(nmemb, size);
    uint64_t nth = __sync_fetch_and_add(&g_nth, 1);
    size_t total;
    if (__builtin_mul_overflow(nmemb, size, &total)) { errno = ENOMEM; return nullptr; }
    if (total && should_manage(total, nth)) {
        void* p = managed_new(total, nth, "calloc");
        if (p) { memset(p, 0, total); return p; }   // reference calloc_managed, blas2cuda.c:150-154
    }
    return __libc_calloc(nmemb, size);
}

void* realloc(void* ptr, size_t request) noexcept {
    if (ptr == nullptr) return malloc(request);
    void* base; size_t old;
    if (!tracker_lookup(ptr, &base, &old) || base != ptr) return __libc_realloc(ptr, request);
    if (request == 0) { tracker_free_managed(ptr); return nullptr; }
    // reference realloc_managed (blas2cuda.c:156-166): new block, copy, free -- stays managed
    void* np = managed_new(request, __sync_fetch_and_add(&g_nth, 1), "realloc");
    if (!np) np = __libc_malloc(request);
    if (!np) return nullptr;
    memcpy(np, ptr, old < request ? old : request);
    tracker_free_managed(ptr);
    return np;
}

void free(void* ptr) noexcept {
    if (!ptr) return;
    uintptr_t p = (uintptr_t)ptr;
    if (p >= g_lo && p < g_hi && tracker_free_managed(ptr)) return;
    __libc_free(ptr);
}

// ---- the aligned allocators (not interposed by the reference, SURVEY.md section 8b: "real programs use them") ----
// Eigen, FFTW-style codes and C++17 aligned new allocate BLAS operands with posix_memalign / aligned_alloc; left to glibc
// they would be untracked and staged on every call.  A managed block is handed out when the placement heuristic says so
// AND its base happens to satisfy the alignment (cudaMallocManaged bases are at least 256-byte aligned, far more for
// large blocks); otherwise the request goes to glibc.  free() / realloc() already route by registry lookup.
static void* aligned_common(size_t alignment, size_t request, const char* fun) {
    if (bypass()) return __libc_memalign(alignment, request);
    uint64_t nth = __sync_fetch_and_add(&g_nth, 1);
    if (request && should_manage(request, nth)) {
        void* p = managed_new(request, nth, fun);
        if (p && ((uintptr_t)p & (alignment - 1)) == 0) return p;
        if (p) tracker_free_managed(p);
    }
    return __libc_memalign(alignment, request);
}
static inline bool pow2(size_t a) { return a && (a & (a - 1)) == 0; }
__attribute__((visibility("default"))) int posix_memalign(void** out, size_t alignment, size_t request) noexcept {
    if (!pow2(alignment) || alignment % sizeof(void*) != 0) return EINVAL;
    void* p = aligned_common(alignment, request, "posix_memalign");
    if (!p) return ENOMEM;      // posix_memalign reports through its return value and leaves errno alone
    *out = p;
    return 0;
}
__attribute__((visibility("default"))) void* aligned_alloc(size_t alignment, size_t request) noexcept {
    if (!pow2(alignment)) { errno = EINVAL; return nullptr; }
    return aligned_common(alignment, request, "aligned_alloc");
}
__attribute__((visibility("default"))) void* memalign(size_t alignment, size_t request) noexcept {
    if (!pow2(alignment)) { errno = EINVAL; return nullptr; }
    return aligned_common(alignment, request, "memalign");
}
__attribute__((visibility("default"))) void* valloc(size_t request) noexcept { return aligned_common((size_t)sysconf(_SC_PAGESIZE), request, "valloc"); }
// glibc would read its chunk header in front of the pointer: answer for managed blocks from the registry
__attribute__((visibility("default"))) size_t malloc_usable_size(void* ptr) noexcept {
    typedef size_t (*usable_t)(void*);
    static usable_t real = nullptr;
    if (!ptr) return 0;
    void* base; size_t size;
    if (tracker_lookup(ptr, &base, &size) && base == ptr) return size;
    if (!real) {
        t_inside++;
        real = (usable_t)dlsym(RTLD_NEXT, "malloc_usable_size");
        t_inside--;
        if (!real) return 0;
    }
    return real(ptr);
}

}  // extern "C"

// Makes the shared object directly executable (`./libb200blas.so` prints the option help through
// b200blas_entry, lifecycle.cu): the kernel needs a program interpreter to relocate it
// (reference entry.c:4, same mechanism).
extern "C" {
__attribute__((used, section(".interp"))) const char b200blas_interp[] = "/lib64/ld-linux-x86-64.so.2";
void b200blas_print_help(void);
// ELF entry point (-e b200blas_entry): entered without a return address on the stack, so the stack
// must be re-aligned before calling into libc.
__attribute__((force_align_arg_pointer, visibility("default"))) void b200blas_entry(void) {
    b200blas_print_help();
    _exit(0);
}
}
