// tracker.h -- object tracker: malloc/calloc/realloc/free interposition, the managed allocator
// and the registry of managed blocks (reference lib/obj_tracker.c, lib/oracle.c,
// blas2cuda.c:127-175).  C API so it is callable from anywhere, including inside malloc.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// reference obj_tracker.c:52 heuristic names (H_RANDOM default there); H_SIZE is the new default
// required by north_star (3): "size-based ... no CPU fallback".
enum b200_heuristic { B200_H_SIZE = 0, B200_H_TRUE, B200_H_FALSE, B200_H_RANDOM, B200_H_ORACLE };

// 1 if p lies inside a block handed out by the managed allocator (interior pointers included,
// reference obj_tracker_objinfo_subptr, obj_tracker.c:602-637); base/size optional outputs.
int tracker_lookup(const void* p, void** base, size_t* size);
// Residency bookkeeping for a tracked block containing p: returns 1 if the block has already been bulk-migrated to
// the device since it was allocated, 0 if not (and marks it), -1 if p is not in a tracked block.
int tracker_test_and_set_resident(const void* p);
int tracker_peek_resident(const void* p);    // the same answer without marking the block
// re-entrancy guard (reference obj_tracker_internal_enter/leave, obj_tracker.c:343-349): while a
// thread is inside, its allocations go straight to glibc.
void tracker_enter(void);
void tracker_leave(void);
void tracker_set_tracking(int on);          // reference obj_tracker_set_tracking
int tracker_get_tracking(void);
void tracker_set_shutdown(void);             // process is exiting: stop tracking, never call into CUDA from free() again
void tracker_set_heuristic(int h);
void tracker_set_threshold(size_t bytes);
int tracker_load_oracle_file(const char* filename);
void tracker_set_trace(int on);                        // T/U/C lines in the reference TRACE_OUTPUT format (obj_tracker.c:426-483)
void tracker_trace_call(const void* ptr, const char* fun);
int tracker_decision(uint64_t nth, size_t request);    // managed-or-not for the nth allocation under the current heuristic   // reference oracle_load_file, oracle.c:26-72
// direct entry points to the managed allocator (reference blas2cuda_manager ctor/dtor); used by
// the C-ABI b200blas_malloc_managed / tests.  NULL on failure.
void* tracker_alloc_managed(size_t bytes);
int tracker_free_managed(void* p);          // 1 if p was a tracked block and has been released
struct b200_tracker_stats { uint64_t allocs_seen, managed_allocs, managed_frees, managed_bytes_live, managed_bytes_peak; };
void tracker_get_stats(struct b200_tracker_stats* out);

#ifdef __cplusplus
}
struct TrackerGuard {   // reference objtracker_guard (runtime-mem.hpp:40)
    TrackerGuard() { tracker_enter(); }
    ~TrackerGuard() { tracker_leave(); }
};
#endif
