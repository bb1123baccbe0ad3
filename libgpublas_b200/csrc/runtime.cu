// runtime.cu -- see runtime.h.
#include <time.h>
#include "runtime.h"
#include "common.cuh"
#include "kernels.h"
#include "tracker.h"
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unistd.h>
#include <sys/syscall.h>

extern "C" void b200_writef(int fd, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (n < 0) return;
    if (n > (int)sizeof buf - 1) n = sizeof buf - 1;
    syscall(SYS_write, fd, buf, (size_t)n);
}

namespace b200 {

Options g_opts;
Stats g_stats = {0, 0, 0, 0, 0, 0};
thread_local int last_variant = VAR_NONE;
thread_local const char* t_call_name = "blas";
int force_variant = VAR_NONE;

const char* variant_name(int v) {
    switch (v) {
        case VAR_SCALE_ONLY: return "scale_only";
        case VAR_GENERIC_TILE: return "generic_tile";
        case VAR_DMMA_TMA: return "dmma_tma";
        case VAR_DMMA_LDG: return "dmma_ldg";
        case VAR_TF32X3_TCGEN05: return "tf32x3_tcgen05";
        default: return "none";
    }
}

void fatal(const char* what, const char* file, int line, const char* detail) {
    b200_writef(STDERR_FILENO, "b200blas: fatal: %s failed at %s:%d: %s\n", what, file, line, detail ? detail : "");
    abort();
}

// ---------------------------------------------------------------------------------------------
static std::once_flag g_init_once;
static bool g_ready = false;
static const char* g_init_error = nullptr;   // set when the device could not come up: fatal for a BLAS call, not for malloc
static int g_sms = 0;
static int g_device = 0;
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn g_encode = nullptr;

static void do_init() {
    TrackerGuard guard;   // CUDA's own allocations must not be routed to managed memory
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        g_init_error = "no CUDA device: libb200blas has no CPU fallback (by design); unset LD_PRELOAD to run on the CPU BLAS";
        return;
    }
    const char* dev_env = getenv("B200BLAS_DEVICE");
    g_device = dev_env ? atoi(dev_env) : -1;
    if (g_device < 0) {
        // keep whatever device the process already selected (torch.cuda.set_device under torchrun)
        B200_CUDA(cudaGetDevice(&g_device));
    } else {
        B200_CUDA(cudaSetDevice(g_device));
    }
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, g_device));
    g_sms = prop.multiProcessorCount;
    if (prop.major != 10)
        b200_writef(STDERR_FILENO, "b200blas: warning: device %s is sm_%d%d; kernels are built for sm_100a only\n",
                    prop.name, prop.major, prop.minor);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
        g_encode = (encode_tiled_fn)fn;
    if (g_opts.debug_exec)
        b200_writef(STDERR_FILENO, "b200blas: device %d %s, %d SMs, %zu MiB, concurrentManagedAccess=%d, tma=%d\n",
                    g_device, prop.name, g_sms, prop.totalGlobalMem >> 20, prop.concurrentManagedAccess, g_encode != nullptr);
    g_ready = true;
    // Registered after the CUDA runtime registered its own teardown, so it runs BEFORE it (LIFO): stdio
    // buffers that the interposed malloc placed in managed memory (heuristic=true, or a large setvbuf)
    // are flushed and detached while the context still exists; later allocations use the heap.
    atexit([] {
        tracker_set_shutdown();
        fflush(NULL);
        if (tracker_lookup(stdout->_IO_buf_base, nullptr, nullptr)) setvbuf(stdout, nullptr, _IONBF, 0);
    });
}

// A BLAS entry point needs the device: no device is fatal (there is no CPU path).  The allocator only *prefers* managed
// memory: try_init() reports failure and the interposed malloc keeps serving from the heap, so preloading the library into
// a process on a box without a GPU (or with CUDA_VISIBLE_DEVICES empty) never kills it unless it actually calls BLAS.
// Both block in call_once while another thread is bringing the device up.
void ensure_init() {
    std::call_once(g_init_once, do_init);
    if (!g_ready) fatal("cudaGetDeviceCount", __FILE__, __LINE__, g_init_error ? g_init_error : "device initialisation failed");
}
bool try_init() {
    std::call_once(g_init_once, do_init);
    return g_ready;
}
bool device_ready() { return g_ready; }
int sm_count() { return g_sms; }
bool tma_available() { return g_encode != nullptr; }

void set_max_dynamic_smem(const void* kernel, int bytes) {
    struct Entry { const void* k; int dev; int bytes; };
    static Entry table[512];
    static int count = 0;
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < count; i++)
        if (table[i].k == kernel && table[i].dev == dev) {
            if (table[i].bytes >= bytes) return;
            table[i].bytes = bytes;
            B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            return;
        }
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (count < 512) table[count++] = Entry{kernel, dev, bytes};
}

bool encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, void* base, const cuuint64_t* gdim,
                       const cuuint64_t* gstride_bytes, const cuuint32_t* box, const cuuint32_t* estride,
                       CUtensorMapSwizzle swz) {
    if (!g_encode) return false;
    CUresult r = g_encode(map, dt, (cuuint32_t)rank, base, gdim, gstride_bytes, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
struct ThreadCtx {
    cudaStream_t stream = nullptr;      // library-owned, non-blocking
    cudaStream_t ext = nullptr;         // caller-provided (b200blas_set_stream)
    bool external_stream = false;
    char* ws = nullptr; size_t ws_cap = 0, ws_used = 0, call_bytes = 0;
    // blocks retired because the workspace had to grow mid-call; freed at the next reset
    void* retired[16]; int nretired = 0;
    void* pinned = nullptr; void* dscalar = nullptr;
    cudaStream_t aux[2] = {nullptr, nullptr};
    cudaEvent_t events[64]; int nevents = 0;
};
// One context per (thread, device): the home device's is what every single-GPU call uses; partitioned Level-3 calls
// (multi_gemm.cu) switch the calling thread to a peer device with DeviceScope and find that device's own stream, workspace,
// scalar slots and events here, so every *_dev launcher works unchanged on any device.
static thread_local ThreadCtx t_ctxs[kMaxDevices];
static thread_local int t_dev = -1;          // -1: the home device
#define t_ctx (t_ctxs[t_dev < 0 ? g_device : t_dev])

int home_device() { ensure_init(); return g_device; }
int current_device() { return t_dev < 0 ? g_device : t_dev; }
DeviceScope::DeviceScope(int dev) : prev_(t_dev) {
    ensure_init();
    t_dev = dev;
    B200_CUDA(cudaSetDevice(dev));
}
DeviceScope::~DeviceScope() {
    t_dev = prev_;
    cudaSetDevice(prev_ < 0 ? g_device : prev_);
}

cudaStream_t current_stream() {
    ensure_init();
    if (t_ctx.external_stream) return t_ctx.ext;
    if (!t_ctx.stream) {
        TrackerGuard guard;
        B200_CUDA(cudaSetDevice(current_device()));
        B200_CUDA(cudaStreamCreateWithFlags(&t_ctx.stream, cudaStreamNonBlocking));
    }
    return t_ctx.stream;
}
void set_thread_stream(cudaStream_t s, bool external) {
    ensure_init();
    t_ctx.ext = s;               // may legitimately be 0: the legacy default stream
    t_ctx.external_stream = external;
}

cudaStream_t aux_stream(int which) {
    ensure_init();
    ThreadCtx& c = t_ctx;
    if (!c.aux[which]) {
        TrackerGuard guard;
        B200_CUDA(cudaSetDevice(current_device()));
        B200_CUDA(cudaStreamCreateWithFlags(&c.aux[which], cudaStreamNonBlocking));
    }
    return c.aux[which];
}
cudaEvent_t pooled_event(int idx) {
    ThreadCtx& c = t_ctx;
    if (idx >= 64) fatal("pooled_event", __FILE__, __LINE__, "event pool exhausted");
    while (c.nevents <= idx) {
        TrackerGuard guard;
        B200_CUDA(cudaEventCreateWithFlags(&c.events[c.nevents], cudaEventDisableTiming));
        c.nevents++;
    }
    return c.events[idx];
}

void ws_reset() {
    ThreadCtx& c = t_ctx;
    if (c.nretired) {
        TrackerGuard guard;
        B200_CUDA(cudaStreamSynchronize(current_stream()));
        for (int i = 0; i < c.nretired; i++) cudaFree(c.retired[i]);
        c.nretired = 0;
    }
    c.ws_used = 0;
    c.call_bytes = 0;
}
void* ws_alloc(size_t bytes) {
    ThreadCtx& c = t_ctx;
    bytes = (bytes + 255) & ~(size_t)255;
    if (c.ws_used + bytes > c.ws_cap) {
        TrackerGuard guard;
        // grow: earlier sub-allocations of this call may still be in use by queued work, so the old
        // block is retired (freed at the next call) rather than freed now.
        // The new block is sized for everything this call has asked for so far, so the NEXT call of the same
        // shape fits in one block and allocates nothing (steady state: zero cudaMalloc/cudaFree per call).
        size_t ncap = c.ws_cap ? c.ws_cap : (size_t)64 << 20;
        while (ncap < c.call_bytes + bytes) ncap *= 2;
        if (c.ws) {
            if (c.nretired == 16) fatal("workspace growth", __FILE__, __LINE__, "too many regrowths in one call");
            c.retired[c.nretired++] = c.ws;
        }
        B200_CUDA(cudaMalloc((void**)&c.ws, ncap));
        c.ws_cap = ncap; c.ws_used = 0;
    }
    void* p = c.ws + c.ws_used;
    c.ws_used += bytes;
    c.call_bytes += bytes;
    return p;
}
void* pinned_scalar() {
    if (!t_ctx.pinned) { TrackerGuard guard; B200_CUDA(cudaMallocHost(&t_ctx.pinned, 256)); }
    return t_ctx.pinned;
}
static thread_local bool t_result_observed = false;
void* armed_scalar(size_t bytes) {
    void* p = pinned_scalar();
    memset(p, 0xFF, bytes);      // all-ones words: a NaN no arithmetic produces / an index no reduction returns
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    return p;
}
void wait_scalar(const void* slot, size_t bytes, size_t word_bytes) {
    // the kernel stores the result word by word (a complex value is two stores): every word must have changed
    const int nw = (int)(bytes / word_bytes);
    struct timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned it = 1;; it++) {
        bool done = true;
        if (word_bytes == 4) { for (int w = 0; w < nw; w++) done = done && ((const volatile uint32_t*)slot)[w] != 0xFFFFFFFFu; }
        else { for (int w = 0; w < nw; w++) done = done && ((const volatile uint64_t*)slot)[w] != ~0ull; }
        if (done) { __atomic_thread_fence(__ATOMIC_ACQUIRE); t_result_observed = true; return; }
        __builtin_ia32_pause();
        if ((it & 1023u) == 0) {   // bounded: a result that really is all-ones, or a faulted kernel, ends in the synchronise below
            struct timespec t1;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if ((t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec) > 20000000ll) break;
        }
    }
    B200_CUDA(cudaStreamSynchronize(current_stream()));
    t_result_observed = true;
}
void* device_scalar() {
    if (!t_ctx.dscalar) {
        TrackerGuard guard;
        B200_CUDA(cudaMalloc(&t_ctx.dscalar, 256));
        B200_CUDA(cudaMemset(t_ctx.dscalar, 0, 256));   // bytes 128.. hold the "last block" tickets of the reductions
    }
    return t_ctx.dscalar;
}

void finish_call() {
    // BLAS is synchronous: results must be visible to the CPU when the symbol returns.  The
    // reference skips this on concurrentManagedAccess devices (runtime.h:33-34) and so races
    // with the caller on every modern GPU; here the per-thread stream is drained unless the
    // caller opted out (device-resident pipelines, b200blas_set_sync(0)).
    const bool observed = t_result_observed;   // a reduction whose value the host has already read (wait_scalar)
    t_result_observed = false;
    if (g_opts.sync && !(observed && !g_opts.debug_execfail)) {
        B200_CUDA(cudaStreamSynchronize(current_stream()));
    } else if (g_opts.debug_execfail) {
        B200_CUDA(cudaStreamSynchronize(current_stream()));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fatal("kernel launch", __FILE__, __LINE__, cudaGetErrorString(e));
    __atomic_fetch_add(&g_stats.calls, 1ull, __ATOMIC_RELAXED);
}

void log_exec(const char* routine, const char* fmt, ...) {
    if (!g_opts.debug_exec) return;
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    b200_writef(STDERR_FILENO, "b200blas: %s %s variant=%s\n", routine, buf, variant_name(last_variant));
}

// ---------------------------------------------------------------------------------------------
Residency classify(const void* p) {
    // 1. the tracker knows every managed block it handed out (reference obj_tracker_objinfo_subptr,
    //    lib/obj_tracker.c:602-637) -- a lock-free range lookup, no CUDA call
    if (tracker_lookup(p, nullptr, nullptr)) return RES_MANAGED;
    // 2. anything else: ask the driver (device pointers from cudaMalloc/torch, managed memory
    //    allocated by the application itself, pinned host memory)
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return RES_HOST_PAGEABLE; }
    switch (at.type) {
        case cudaMemoryTypeDevice: return RES_DEVICE;
        case cudaMemoryTypeManaged: return RES_MANAGED;
        case cudaMemoryTypeHost: return RES_HOST_PINNED;
        default: return RES_HOST_PAGEABLE;
    }
}


// Managed-operand residency (north_star (3)).  cudaMemPrefetchAsync over an already-resident range is NOT free:
// the driver walks every 2 MiB page of the range (measured on B200: ~0.4 ms per 512 MiB vector, ~20 ms for the
// 8 GiB matrix of config 3 -- 15x the DGEMV it precedes).  So a tracked block is bulk-migrated once, on its first
// use after allocation (the CPU has just initialised it), given the device as preferred location, and afterwards
// left to the page-fault path: pages the CPU touches between calls come back on demand, everything else stays in
// HBM and chained Level-1/2 calls run at device-memory speed.  prefetch=2 restores prefetch-on-every-call.
void make_resident(const void* p, size_t bytes, cudaStream_t s) {
    if (!g_opts.prefetch || bytes < ((size_t)2 << 20)) return;
    if (g_opts.prefetch == 1) {
        const int prev = tracker_test_and_set_resident(p);
        if (prev != 0) return;          // already migrated once, or not one of our blocks (the application manages those)
    }
    TrackerGuard guard;
    int dev = current_device();
    void* base = nullptr; size_t bsize = 0;
    if (g_opts.prefetch == 1 && tracker_lookup(p, &base, &bsize)) {    // whole block: later calls may use other parts of it
        cudaMemAdvise(base, bsize, cudaMemAdviseSetPreferredLocation, dev);
        if (cudaMemPrefetchAsync(base, bsize, dev, s) == cudaSuccess) __atomic_fetch_add(&g_stats.prefetch_bytes, (unsigned long long)bsize, __ATOMIC_RELAXED);
        else cudaGetLastError();
        return;
    }
    if (cudaMemPrefetchAsync(p, bytes, dev, s) == cudaSuccess) __atomic_fetch_add(&g_stats.prefetch_bytes, (unsigned long long)bytes, __ATOMIC_RELAXED);
    else cudaGetLastError();
}

Operand::Operand(const void* host, int64_t rows, int64_t cols, int64_t ld, size_t elem, int access)
    : host_(host), dev_(nullptr), rows_(rows), cols_(cols), ld_(ld), dld_(ld), elem_(elem), access_(access),
      staged_(false), done_(false) {
    if (rows <= 0 || cols <= 0 || host == nullptr) { dev_ = (void*)host; done_ = true; return; }
    Residency r = classify(host);
    pageable_ = r == RES_HOST_PAGEABLE;
    cudaStream_t s = current_stream();
    if (r == RES_DEVICE || r == RES_MANAGED) {
        dev_ = (void*)host;
        __atomic_fetch_add(&g_stats.hits, 1ull, __ATOMIC_RELAXED);
        if (r == RES_MANAGED) tracker_trace_call(host, t_call_name);
        if (r == RES_MANAGED) make_resident(host, (size_t)((cols - 1) * ld + rows) * elem, s);
        done_ = true;   // nothing to write back
        return;
    }
    // miss: stage a compact copy.  Leading dimension rounded so that rows*elem is a multiple of 16 B
    // (TMA-addressable) -- an unaligned host lda never forces the slow loader.
    __atomic_fetch_add(&g_stats.misses, 1ull, __ATOMIC_RELAXED);
    staged_ = true;
    int64_t per16 = 16 / (int64_t)elem; if (per16 < 1) per16 = 1;
    dld_ = (rows + per16 - 1) / per16 * per16;
    dev_ = ws_alloc((size_t)dld_ * cols * elem);
    if (access & ACC_IN) {
        TrackerGuard guard;
        // one column (every vector, packed matrices): a flat copy -- a 2-D copy's pitch is capped at cudaDeviceProp::memPitch (2 GiB)
        const bool bounce = pageable_ && staged_copy_worthwhile((size_t)rows * cols * elem);     // host_stager.cu
        if (bounce) staged_copy_2d(dev_, (size_t)dld_ * elem, host, (size_t)ld * elem, (size_t)rows * elem, (size_t)cols, true, s);
        else if (cols == 1) B200_CUDA(cudaMemcpyAsync(dev_, host, (size_t)rows * elem, cudaMemcpyHostToDevice, s));
        else B200_CUDA(cudaMemcpy2DAsync(dev_, (size_t)dld_ * elem, host, (size_t)ld * elem, (size_t)rows * elem, (size_t)cols,
                                         cudaMemcpyHostToDevice, s));
        __atomic_fetch_add(&g_stats.h2d_bytes, (unsigned long long)(rows * cols * elem), __ATOMIC_RELAXED);
        if (g_opts.trace_copy)
            b200_writef(STDOUT_FILENO, "b200blas: copy %zu B from %p (CPU) ---> %p (GPU)\n", (size_t)(rows * cols * elem), host, dev_);
    }
}

void Operand::release() {
    if (done_) return;
    done_ = true;
    if (staged_ && (access_ & ACC_OUT)) {
        TrackerGuard guard;
        const bool bounce = pageable_ && !(cols_ == 1 && vec_inc_ > 1) && staged_copy_worthwhile((size_t)rows_ * cols_ * elem_);
        if (bounce) staged_copy_2d((void*)host_, (size_t)ld_ * elem_, dev_, (size_t)dld_ * elem_, (size_t)rows_ * elem_, (size_t)cols_, false, current_stream());
        else if (cols_ == 1 && vec_inc_ > 1 && vec_n_ > 0)   // strided vector: only its own elements go back (see runtime.h)
            B200_CUDA(cudaMemcpy2DAsync((void*)host_, (size_t)vec_inc_ * elem_, dev_, (size_t)vec_inc_ * elem_, elem_, (size_t)vec_n_,
                                        cudaMemcpyDeviceToHost, current_stream()));
        else if (cols_ == 1) B200_CUDA(cudaMemcpyAsync((void*)host_, dev_, (size_t)rows_ * elem_, cudaMemcpyDeviceToHost, current_stream()));
        else B200_CUDA(cudaMemcpy2DAsync((void*)host_, (size_t)ld_ * elem_, dev_, (size_t)dld_ * elem_, (size_t)rows_ * elem_,
                                         (size_t)cols_, cudaMemcpyDeviceToHost, current_stream()));
        __atomic_fetch_add(&g_stats.d2h_bytes, (unsigned long long)(rows_ * cols_ * elem_), __ATOMIC_RELAXED);
        if (g_opts.trace_copy)
            b200_writef(STDOUT_FILENO, "b200blas: copy %zu B from %p (GPU) ---> %p (CPU)\n", (size_t)(rows_ * cols_ * elem_), dev_, host_);
    }
}

}  // namespace b200
