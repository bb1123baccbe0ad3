/* oracle/ref_shims.c -- TEST INFRASTRUCTURE.  The two allocator hooks the reference's lib/oracle.c expects from
 * lib/obj_tracker.c (internal_realloc / internal_free, obj_tracker.c:142-223 resolve them with dlsym(RTLD_NEXT)), provided
 * by libc here so that oracle.c can be built on its own (oracle/build_ref.sh). */
#include <stdlib.h>
void *internal_malloc(size_t n) { return malloc(n); }
void *internal_calloc(size_t a, size_t b) { return calloc(a, b); }
void *internal_realloc(void *p, size_t n) { return realloc(p, n); }
void internal_free(void *p) { free(p); }
