/* allocs_mt.c -- allocator churn from several threads at once under the interposer: malloc / calloc / realloc /
 * posix_memalign / free with content checks, blocks handed from one thread to another through a shared mailbox (free on a
 * different thread than the allocation).  The reference guards its registry with a rwlock and per-thread re-entrancy flags
 * (lib/obj_tracker.c:65-69, :557-589); this is the stress test of the replacement (tracker.cpp).
 * Usage: allocs_mt <threads> <iterations> <big_every> [late_thread]
 *   every <big_every>-th block is >= 64 KiB (managed under the default heuristic).  With <late_thread> = t >= 0, threads other than
 *   t make only SMALL allocations until thread t has made the process's first big one (which brings the device up) -- they are
 *   mid-loop, spinning through small blocks, when the bring-up starts, and their big blocks follow as soon as it has begun.
 * Prints: RESULT ok=<0|1> threads=<..> iters=<..> tracked_seen=<..> big=<blocks >= 64 KiB requested through malloc/calloc/posix_memalign> */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int (*is_tracked)(const void*);
static int g_iters, g_big_every, g_fail;
static long g_tracked, g_big_direct;
static int g_late = -1;
static volatile int g_first_big_started;
#define MAILBOX 64
static void* volatile g_mail[MAILBOX];          /* blocks waiting to be freed by whoever takes them */
static volatile size_t g_mail_size[MAILBOX];

static int check(const unsigned char* p, size_t n, unsigned char v) {
    for (size_t i = 0; i < n; i += 257) if (p[i] != v) return 0;
    return n == 0 || p[n - 1] == v;
}

static void* worker(void* arg) {
    const unsigned id = (unsigned)(uintptr_t)arg;
    unsigned long long seed = 0x9E3779B97F4A7C15ull * (id + 1);
    long tracked = 0, big_direct = 0;
    if (g_late >= 0 && (int)id != g_late) {            /* churn through small blocks until the late thread starts the first big one */
        while (!g_first_big_started) { void* q = malloc(1 + (size_t)(seed++ % 3000)); if (!q) __sync_fetch_and_add(&g_fail, 1); free(q); }
    }
    for (int it = 0; it < g_iters; it++) {
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        const int big = g_big_every > 0 && it % g_big_every == 0;
        if (big && g_late >= 0 && (int)id == g_late) g_first_big_started = 1;
        if (big && (seed >> 20) % 4 != 3) big_direct++;
        size_t sz = big ? 65536 + (size_t)((seed >> 33) % 200000) : 1 + (size_t)((seed >> 33) % 3000);
        const unsigned char tag = (unsigned char)(id * 31 + it);
        unsigned char* p;
        switch ((seed >> 20) % 4) {
            case 0: p = malloc(sz); break;
            case 1: p = calloc(sz, 1); if (p && !check(p, sz, 0)) __sync_fetch_and_add(&g_fail, 1); break;
            case 2: if (posix_memalign((void**)&p, 64, sz) != 0) p = NULL; else if ((uintptr_t)p & 63) __sync_fetch_and_add(&g_fail, 1); break;
            default: p = malloc(sz / 2 + 1); if (p) { memset(p, tag, sz / 2 + 1); p = realloc(p, sz); if (p && !check(p, sz / 2 + 1, tag)) __sync_fetch_and_add(&g_fail, 1); } break;
        }
        if (!p) { __sync_fetch_and_add(&g_fail, 1); continue; }
        memset(p, tag, sz);
        if (is_tracked && is_tracked(p)) tracked++;
        /* hand the block to the mailbox; free whatever was there (usually another thread's block) */
        const unsigned slot = (unsigned)((seed >> 40) % MAILBOX);
        size_t old_size = g_mail_size[slot];
        void* old = __atomic_exchange_n(&g_mail[slot], (void*)p, __ATOMIC_ACQ_REL);
        g_mail_size[slot] = sz;
        (void)old_size;
        if (old) { volatile unsigned char touch = *(unsigned char*)old; (void)touch; free(old); }
    }
    __sync_fetch_and_add(&g_tracked, tracked);
    __sync_fetch_and_add(&g_big_direct, big_direct);
    return NULL;
}

int main(int argc, char** argv) {
    const int nthreads = argc > 1 ? atoi(argv[1]) : 8;
    g_iters = argc > 2 ? atoi(argv[2]) : 20000;
    g_big_every = argc > 3 ? atoi(argv[3]) : 50;
    g_late = argc > 4 ? atoi(argv[4]) : -1;
    is_tracked = (int (*)(const void*))dlsym(RTLD_DEFAULT, "b200blas_is_tracked");
    pthread_t th[64];
    for (int t = 0; t < nthreads && t < 64; t++) pthread_create(&th[t], NULL, worker, (void*)(uintptr_t)t);
    for (int t = 0; t < nthreads && t < 64; t++) pthread_join(th[t], NULL);
    for (int s = 0; s < MAILBOX; s++) free(g_mail[s]);
    printf("RESULT ok=%d threads=%d iters=%d tracked_seen=%ld big=%ld\n", g_fail == 0, nthreads, g_iters, g_tracked, g_big_direct);
    return g_fail == 0 ? 0 : 1;
}
