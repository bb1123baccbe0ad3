/* um_counters.c -- TEST INFRASTRUCTURE.  A preloadable helper that counts unified-memory traffic of the process with CUPTI
 * (CUPTI_ACTIVITY_KIND_UNIFIED_MEMORY_COUNTER: bytes migrated host->device and device->host, GPU page-fault groups, CPU page
 * faults) and prints the totals at exit:
 *     UMCOUNT htod_bytes=<..> dtoh_bytes=<..> gpu_fault_groups=<..> cpu_faults=<..> records=<..>
 * SURVEY.md section 8(d) asks for "nsys --cuda-um-gpu-page-faults ... proving zero host migrations" for the Level-1/2 chains on
 * managed memory (config 3); nsys is not in this image, CUPTI is.  Used as
 *     LD_PRELOAD="libumcount.so libb200blas.so" ./cg_chain <n> <iters>
 * with two iteration counts: traffic that does not grow with the iteration count is the one-off first-touch migration, and the
 * steady-state iterations fault nothing back.  tools/um_chain_counters.sh builds and runs it on a GPU box. */
#define _GNU_SOURCE
#include <cuda.h>
#include <cupti.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static unsigned long long g_htod, g_dtoh, g_gpu_faults, g_cpu_faults, g_records;
static int g_on;

static void CUPTIAPI buffer_requested(uint8_t** buffer, size_t* size, size_t* max_records) {
    *size = 1 << 20;
    *buffer = (uint8_t*)malloc(*size + 8);
    *max_records = 0;
}
static void CUPTIAPI buffer_completed(CUcontext ctx, uint32_t stream, uint8_t* buffer, size_t size, size_t valid) {
    (void)ctx; (void)stream; (void)size;
    CUpti_Activity* rec = NULL;
    while (cuptiActivityGetNextRecord(buffer, valid, &rec) == CUPTI_SUCCESS) {
        if (rec->kind != CUPTI_ACTIVITY_KIND_UNIFIED_MEMORY_COUNTER) continue;
        CUpti_ActivityUnifiedMemoryCounter2* um = (CUpti_ActivityUnifiedMemoryCounter2*)rec;
        g_records++;
        switch (um->counterKind) {
            case CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_BYTES_TRANSFER_HTOD: g_htod += um->value; break;
            case CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_BYTES_TRANSFER_DTOH: g_dtoh += um->value; break;
            case CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_GPU_PAGE_FAULT: g_gpu_faults += um->value; break;
            case CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_CPU_PAGE_FAULT_COUNT: g_cpu_faults += 1; break;
            default: break;
        }
    }
    free(buffer);
}

__attribute__((constructor)) static void umcount_start(void) {
    if (cuInit(0) != CUDA_SUCCESS) { fprintf(stderr, "UMCOUNT unavailable: cuInit failed\n"); return; }
    int ndev = 0;
    cuDeviceGetCount(&ndev);
    if (ndev < 1) { fprintf(stderr, "UMCOUNT unavailable: no device\n"); return; }
    CUpti_ActivityUnifiedMemoryCounterConfig cfg[4];
    memset(cfg, 0, sizeof cfg);
    const CUpti_ActivityUnifiedMemoryCounterKind kinds[4] = {CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_BYTES_TRANSFER_HTOD,
                                                             CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_BYTES_TRANSFER_DTOH,
                                                             CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_GPU_PAGE_FAULT,
                                                             CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_KIND_CPU_PAGE_FAULT_COUNT};
    for (int i = 0; i < 4; i++) {
        cfg[i].scope = CUPTI_ACTIVITY_UNIFIED_MEMORY_COUNTER_SCOPE_PROCESS_SINGLE_DEVICE;
        cfg[i].kind = kinds[i];
        cfg[i].deviceId = 0;
        cfg[i].enable = 1;
    }
    CUptiResult r = cuptiActivityRegisterCallbacks(buffer_requested, buffer_completed);
    if (r == CUPTI_SUCCESS) r = cuptiActivityConfigureUnifiedMemoryCounter(cfg, 4);
    if (r == CUPTI_SUCCESS) r = cuptiActivityEnable(CUPTI_ACTIVITY_KIND_UNIFIED_MEMORY_COUNTER);
    if (r != CUPTI_SUCCESS) {
        const char* msg = NULL;
        cuptiGetResultString(r, &msg);
        fprintf(stderr, "UMCOUNT unavailable: %s\n", msg ? msg : "CUPTI error");
        return;
    }
    g_on = 1;
}
__attribute__((destructor)) static void umcount_stop(void) {
    if (!g_on) return;
    cuptiActivityFlushAll(1);
    printf("UMCOUNT htod_bytes=%llu dtoh_bytes=%llu gpu_fault_groups=%llu cpu_faults=%llu records=%llu\n", g_htod, g_dtoh, g_gpu_faults, g_cpu_faults, g_records);
    fflush(stdout);
}
