/* allocs.c -- allocator-tracker churn, after the reference's lib/tests/allocs.c (512 mallocs of growing
 * size, strided frees) extended with calloc/realloc and a check that contents survive.  Under
 * LD_PRELOAD=libb200blas.so blocks >= the threshold come from the managed allocator.
 * Prints: RESULT ok=<0|1> tracked=<count of blocks the library reports as tracked> */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define N 512
int main(void) {
    int (*is_tracked)(const void*) = (int (*)(const void*))dlsym(RTLD_DEFAULT, "b200blas_is_tracked");
    unsigned char* blk[N]; size_t sz[N];
    int ok = 1, tracked = 0;
    for (int i = 0; i < N; i++) {
        sz[i] = (size_t)(i + 1) * 1024;                     /* 1 KiB .. 512 KiB: crosses the 64 KiB threshold */
        blk[i] = (i % 3 == 0) ? calloc(sz[i], 1) : malloc(sz[i]);
        if (!blk[i]) { ok = 0; break; }
        if (i % 3 == 0) for (size_t b = 0; b < sz[i]; b += 997) if (blk[i][b] != 0) ok = 0;
        memset(blk[i], i & 0xff, sz[i]);
        if (is_tracked && is_tracked(blk[i])) { tracked++; if (!is_tracked(blk[i] + sz[i] / 2)) ok = 0; }   /* interior pointer */
    }
    for (int i = 0; i < N; i += 2) { free(blk[i]); blk[i] = NULL; }          /* strided frees */
    for (int i = 1; i < N; i += 2) {                                          /* realloc: grow and shrink, contents kept */
        size_t ns = (i % 4 == 1) ? sz[i] * 2 : sz[i] / 2 + 1;
        unsigned char* nb = realloc(blk[i], ns);
        if (!nb) { ok = 0; break; }
        size_t keep = ns < sz[i] ? ns : sz[i];
        for (size_t b = 0; b < keep; b += 101) if (nb[b] != (unsigned char)(i & 0xff)) ok = 0;
        blk[i] = nb; sz[i] = ns;
    }
    for (int i = 1; i < N; i += 2) free(blk[i]);
    free(NULL);
    void* z = realloc(NULL, 100000); if (!z) ok = 0; z = realloc(z, 0);      /* realloc edge cases */
    printf("RESULT ok=%d tracked=%d\n", ok, tracked);
    return ok ? 0 : 1;
}
