// lifecycle.cu -- library constructor/destructor, BLAS2CUDA_OPTIONS, statistics.csv and the
// control half of the C ABI (reference blas2cuda.c:59-124, :178-277; entry.c:4-11).
#include "abi_common.h"
#include "../../include/b200blas.h"
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <mutex>
#include <execinfo.h>
#include <signal.h>

using namespace b200;

static void print_help() {
    b200_writef(STDERR_FILENO,
        "b200blas options (set BLAS2CUDA_OPTIONS):\n"
        "   You can chain these options with a semicolon (;)\n"
        "   help            -- print help\n"
        "   debug_execfail  -- debug kernel failures (synchronise and check after every call)\n"
        "   debug_exec      -- debug kernel invocations (one line per BLAS call: shapes, variant)\n"
        "   trace_copy      -- trace copies between CPU and GPU\n"
        "   trace           -- object trace on stdout (T/U/C lines, the reference's TRACE_OUTPUT format; input of\n"
        "                      scripts/analyze_trace.py and of heuristic=oracle:<file>)\n"
        "   heuristic=<val> -- one of: 'size' (default), 'random', 'true', 'false', or:\n"
        "                      'oracle:<filename>', where <filename> is the name of an object trace\n"
        "   threshold=<n>   -- heuristic=size: allocations of >= n bytes become managed (default 65536)\n"
        "   variant=<name>  -- force a kernel variant: generic_tile | dmma_tma | dmma_ldg | auto\n"
        "   sgemm_cfg=<n>   -- SGEMM tensor-core tile: -1 size-based (default), 0 128x128, 1 128x256 BK32, 2 128x256 BK16\n"
        "   devices=<n>     -- GPUs used for partitioned Level-3 calls (default 1): ?gemm_ calls of at least multi_min^3 work are\n"
        "                      2-D tile-partitioned over n devices of this box from inside the symbol (peer access required)\n"
        "   multi_min=<n>   -- partition only products with m*n*k >= n^3 (default 8192)\n"
        "   sync=<0|1>      -- block until results are visible before returning (default 1)\n"
        "   prefetch=<0|1|2> -- managed operands: 0 never prefetch, 1 bulk-migrate a tracked block to the device on its\n"
        "                      first use (default), 2 prefetch on every call\n"
        "   pipeline_min=<n> -- host-resident GEMM operands of >= n bytes in total are staged in chunks\n"
        "                      overlapped with compute (default 64 MiB)\n");
}

static int variant_from_name(const char* n) {
    if (!strcmp(n, "generic_tile")) return VAR_GENERIC_TILE;
    if (!strcmp(n, "dmma_tma")) return VAR_DMMA_TMA;
    if (!strcmp(n, "dmma_ldg")) return VAR_DMMA_LDG;
    if (!strcmp(n, "tf32x3_tcgen05")) return VAR_TF32X3_TCGEN05;
    return VAR_NONE;
}

static void set_options(const char* env) {
    if (!env) return;
    char* copy = strdup(env);
    char* save = nullptr;
    for (char* opt = strtok_r(copy, ";", &save); opt; opt = strtok_r(nullptr, ";", &save)) {
        if (!strcmp(opt, "help")) print_help();
        else if (!strcmp(opt, "debug_execfail")) g_opts.debug_execfail = true;
        else if (!strcmp(opt, "debug_exec")) g_opts.debug_exec = true;
        else if (!strcmp(opt, "trace_copy")) g_opts.trace_copy = true;
        else if (!strcmp(opt, "trace")) tracker_set_trace(1);
        else if (!strncmp(opt, "heuristic=", 10)) {
            const char* h = opt + 10;
            if (!strncmp(h, "random", 6)) tracker_set_heuristic(B200_H_RANDOM);
            else if (!strncmp(h, "true", 4)) tracker_set_heuristic(B200_H_TRUE);
            else if (!strncmp(h, "false", 5)) tracker_set_heuristic(B200_H_FALSE);
            else if (!strncmp(h, "size", 4)) tracker_set_heuristic(B200_H_SIZE);
            else if (!strncmp(h, "oracle:", 7)) {
                if (!tracker_load_oracle_file(h + 7)) {
                    b200_writef(STDERR_FILENO, "b200blas:oracle: failed to load '%s'\n", h + 7);
                    abort();
                }
                tracker_set_heuristic(B200_H_ORACLE);
            } else {
                b200_writef(STDERR_FILENO, "b200blas: unsupported heuristic '%s'\n", h);
                abort();
            }
            b200_writef(STDERR_FILENO, "b200blas: selecting heuristic %s\n", h);
        }
        else if (!strncmp(opt, "threshold=", 10)) { g_opts.managed_threshold = strtoull(opt + 10, nullptr, 0); tracker_set_threshold(g_opts.managed_threshold); }
        else if (!strncmp(opt, "variant=", 8)) force_variant = variant_from_name(opt + 8);
        else if (!strncmp(opt, "sgemm_cfg=", 10)) g_opts.sgemm_cfg = atoi(opt + 10);
        else if (!strncmp(opt, "devices=", 8)) g_opts.devices = atoi(opt + 8);
        else if (!strncmp(opt, "multi_min=", 10)) g_opts.multi_gpu_min_dim = strtoull(opt + 10, nullptr, 0);
        else if (!strncmp(opt, "sync=", 5)) g_opts.sync = atoi(opt + 5) != 0;
        else if (!strncmp(opt, "prefetch=", 9)) g_opts.prefetch = atoi(opt + 9);
        else if (!strncmp(opt, "pipeline_min=", 13)) g_opts.pipeline_min_bytes = strtoull(opt + 13, nullptr, 0);
        else b200_writef(STDERR_FILENO, "b200blas: unknown option '%s'. Set BLAS2CUDA_OPTIONS=help.\n", opt);
    }
    free(copy);
}

// Constructor: cheap on purpose.  The reference creates the cuBLAS handle and prints device
// properties in every process that loads it (blas2cuda.c:178-242); here the device comes up lazily
// at the first BLAS call or first managed allocation, so preloading into `sh`, `make`, ... is free.
static void segv_backtrace(int sig) {
    void* frames[64];
    int n = backtrace(frames, 64);
    b200_writef(STDERR_FILENO, "b200blas: signal %d, backtrace:\n", sig);
    backtrace_symbols_fd(frames, n, STDERR_FILENO);
    _exit(128 + sig);
}

__attribute__((constructor)) static void b200blas_ctor() {
    if (getenv("B200BLAS_DEBUG_SEGV")) { signal(SIGSEGV, segv_backtrace); signal(SIGABRT, segv_backtrace); }
    set_options(getenv("BLAS2CUDA_OPTIONS"));
    tracker_set_tracking(1);
}

__attribute__((destructor)) static void b200blas_dtor() {
    tracker_set_shutdown();
    if (g_stats.calls == 0) return;
    // reference blas2cuda.c:266-273: ./statistics.csv with the hit/miss counters
    const char* path = getenv("B200BLAS_STATS_FILE");
    if (!path) path = "statistics.csv";
    FILE* f = fopen(path, "w");
    if (f) {
        fprintf(f, "Hits, Misses, Calls, H2D bytes, D2H bytes, Prefetched bytes\n%llu, %llu, %llu, %llu, %llu, %llu\n",
                g_stats.hits, g_stats.misses, g_stats.calls, g_stats.h2d_bytes, g_stats.d2h_bytes, g_stats.prefetch_bytes);
        fclose(f);
    }
}

// ---- xerbla plumbing ----
static b200blas_xerbla_fn g_xerbla_override = nullptr;
namespace b200 {
void call_xerbla(const char* routine, int info) {
    char srname[8];
    int i = 0;
    for (; i < 6 && routine[i] && routine[i] != '_'; i++) {
        char c = routine[i];
        srname[i] = (c >= 'a' && c <= 'z') ? c - 32 : c;
    }
    for (; i < 6; i++) srname[i] = ' ';
    srname[6] = 0;
    if (g_xerbla_override) { g_xerbla_override(srname, &info, 6); return; }
    typedef void (*xerbla_t)(const char*, int*, size_t);
    xerbla_t x = (xerbla_t)dlsym(RTLD_DEFAULT, "xerbla_");
    if (x) { x(srname, &info, 6); return; }
    // netlib XERBLA's message; netlib stops the program, the CPU BLAS in this image returns
    b200_writef(STDERR_FILENO, " ** On entry to %s parameter number %2d had an illegal value\n", srname, info);
}
}  // namespace b200

extern "C" {

int b200blas_version(void) { return 100; }
void b200blas_set_xerbla(b200blas_xerbla_fn fn) { g_xerbla_override = fn; }
void b200blas_set_stream(void* stream) { set_thread_stream((cudaStream_t)stream, true); }
void b200blas_reset_stream(void) { set_thread_stream(nullptr, false); }
void b200blas_set_sync(int on) { g_opts.sync = on != 0; }
void b200blas_set_options(const char* opts) { set_options(opts); }
const char* b200blas_last_variant(void) { return variant_name(last_variant); }
void b200blas_force_variant(const char* name) { force_variant = name ? variant_from_name(name) : VAR_NONE; }
void b200blas_get_stats(struct b200blas_stats* out) {
    out->hits = g_stats.hits; out->misses = g_stats.misses; out->calls = g_stats.calls;
    out->h2d_bytes = g_stats.h2d_bytes; out->d2h_bytes = g_stats.d2h_bytes; out->prefetch_bytes = g_stats.prefetch_bytes;
    b200_tracker_stats ts; tracker_get_stats(&ts);
    out->managed_allocs = ts.managed_allocs; out->managed_frees = ts.managed_frees; out->managed_bytes_live = ts.managed_bytes_live;
}
void* b200blas_malloc_managed(size_t bytes) { TrackerGuard g; return tracker_alloc_managed(bytes); }
void b200blas_free_managed(void* p) { TrackerGuard g; tracker_free_managed(p); }
int b200blas_is_tracked(const void* p) { return tracker_lookup(p, nullptr, nullptr); }
int b200blas_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
void b200blas_synchronize(void) { B200_CUDA(cudaStreamSynchronize(current_stream())); }

void b200blas_print_help(void) { print_help(); }

// ---- raw device memory + CUDA IPC, for partitioned Level-3 calls across processes (one process per
// GPU): the home rank exports C, every other rank maps it and its GEMM epilogue stores its C tile
// straight into the home allocation over NVLink (SURVEY.md section 8e "C-tile return fused into the
// epilogue (peer stores)"). ----
void* b200blas_device_malloc(size_t bytes) {
    ensure_init();
    TrackerGuard g;
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void b200blas_device_free(void* p) { TrackerGuard g; cudaFree(p); }
int b200blas_ipc_get_handle(void* dev_ptr, void* handle64) {
    ensure_init();
    TrackerGuard g;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, dev_ptr) != cudaSuccess) { cudaGetLastError(); return -1; }
    memcpy(handle64, &h, sizeof h);
    return (int)sizeof h;
}
void* b200blas_ipc_open(const void* handle64) {
    ensure_init();
    TrackerGuard g;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof h);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void b200blas_ipc_close(void* p) { TrackerGuard g; cudaIpcCloseMemHandle(p); }

// Where the driver last placed a managed range: device ordinal, -1 = host, -2 = not managed / unknown.
int b200blas_tracker_decision(unsigned long long nth, size_t request) { return tracker_decision(nth, request); }
int b200blas_residency(const void* p, size_t bytes) {
    ensure_init();
    TrackerGuard g;
    int loc = -2;
    if (cudaMemRangeGetAttribute(&loc, sizeof loc, cudaMemRangeAttributeLastPrefetchLocation, p, bytes) != cudaSuccess) { cudaGetLastError(); return -2; }
    return loc;   // cudaCpuDeviceId == -1, cudaInvalidDeviceId == -2
}

// ---- building blocks of the copy-engine panel push (libgpublas_b200/multigpu.py) ----
// 2-D strided device->device copy on a caller stream; dst may be a CUDA-IPC peer mapping (then the bytes
// cross NVLink on a copy engine, no SM involved).
void b200blas_copy2d_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width_bytes, size_t height, void* stream) {
    ensure_init();
    TrackerGuard g;
    B200_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width_bytes, height, cudaMemcpyDefault, (cudaStream_t)stream));
}
// Stream-ordered 32-bit flag write (value < 65536) into local or peer-mapped device memory: a 4-byte DMA from a
// device-resident table of constants, so it is ordered after the panel copies queued before it on `stream`.
void b200blas_write_flag_async(void* dst_flag, unsigned value, void* stream) {
    ensure_init();
    TrackerGuard g;
    static unsigned* table = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        unsigned* host = (unsigned*)malloc(65536 * sizeof(unsigned));
        for (unsigned i = 0; i < 65536; i++) host[i] = i;
        B200_CUDA(cudaMalloc((void**)&table, 65536 * sizeof(unsigned)));
        B200_CUDA(cudaMemcpy(table, host, 65536 * sizeof(unsigned), cudaMemcpyHostToDevice));
        free(host);
    });
    B200_CUDA(cudaMemcpyAsync(dst_flag, table + (value & 0xffffu), sizeof(unsigned), cudaMemcpyDefault, (cudaStream_t)stream));
}
// Stream-ordered wait until *flag >= value (flag in this device's memory, written remotely by a peer's copy engine):
// a one-thread kernel that polls with system-scope acquire loads, so everything queued behind it on `stream` sees the
// data that was pushed before the flag.
__global__ void b200_wait_flag_kernel(const unsigned* flag, unsigned value) {
    unsigned v;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= value) break;
        __nanosleep(500);
    }
}
void b200blas_wait_flag_async(const void* flag, unsigned value, void* stream) {
    ensure_init();
    TrackerGuard g;
    b200_wait_flag_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const unsigned*)flag, value);
    B200_CUDA(cudaGetLastError());
}
void b200blas_memset_async(void* dst, int byte, size_t bytes, void* stream) {
    ensure_init();
    TrackerGuard g;
    B200_CUDA(cudaMemsetAsync(dst, byte, bytes, (cudaStream_t)stream));
}

}  // extern "C"
