/* aligned_allocs.c -- the aligned allocators under the interposer (posix_memalign, aligned_alloc, memalign, valloc,
 * malloc_usable_size, and realloc / free of what they return).  The reference interposes only malloc / calloc / realloc /
 * free (lib/obj_tracker.c:789,842,902,948), so aligned BLAS operands are never tracked there; SURVEY.md section 8b.
 * Prints: RESULT ok=<0|1> tracked=<blocks the library reports as tracked> */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <malloc.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(void) {
    int (*is_tracked)(const void*) = (int (*)(const void*))dlsym(RTLD_DEFAULT, "b200blas_is_tracked");
    int ok = 1, tracked = 0;
    const size_t aligns[] = {8, 16, 64, 256, 4096, 1 << 16, 1 << 21};
    const size_t sizes[] = {1, 100, 4096, 70000, 1 << 20, (1 << 22) + 24};
    for (unsigned ai = 0; ai < sizeof aligns / sizeof *aligns; ai++)
        for (unsigned si = 0; si < sizeof sizes / sizeof *sizes; si++) {
            const size_t al = aligns[ai], sz = sizes[si];
            void* p[3] = {NULL, NULL, NULL};
            if (posix_memalign(&p[0], al, sz) != 0) ok = 0;
            p[1] = aligned_alloc(al, (sz + al - 1) / al * al);
            p[2] = memalign(al, sz);
            for (int k = 0; k < 3; k++) {
                if (!p[k] || ((uintptr_t)p[k] & (al - 1))) { ok = 0; continue; }
                memset(p[k], 0x5a + k, sz);
                if (malloc_usable_size(p[k]) < sz) ok = 0;
                if (is_tracked && is_tracked(p[k])) tracked++;
            }
            /* realloc of an aligned block keeps the contents (alignment is not preserved by realloc, as in glibc) */
            unsigned char* q = realloc(p[0], sz * 2 + 1);
            if (!q) ok = 0; else { for (size_t b = 0; b < sz; b += 61) if (q[b] != 0x5a) ok = 0; free(q); }
            for (size_t b = 0; b < sz; b += 61) if (((unsigned char*)p[2])[b] != 0x5c) ok = 0;
            free(p[1]); free(p[2]);
        }
    void* v = valloc(100000);
    if (!v || ((uintptr_t)v & 4095)) ok = 0;
    if (is_tracked && v && is_tracked(v)) tracked++;
    free(v);
    void* bad = (void*)1;
    if (posix_memalign(&bad, 24, 100) != EINVAL || bad != (void*)1) ok = 0;      /* not a power of two: EINVAL, *out untouched */
    if (malloc_usable_size(NULL) != 0) ok = 0;
    printf("RESULT ok=%d tracked=%d\n", ok, tracked);
    return ok ? 0 : 1;
}
