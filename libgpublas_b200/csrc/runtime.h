// runtime.h -- host runtime of libb200blas: lazy device init, per-thread stream + workspace,
// operand residency (the replacement for the reference's gpuptr<T>, runtime-mem.hpp:19-176, and
// call_kernel, runtime.h:28-36), options and statistics (blas2cuda.c:59-124, :266-273).
#pragma once
#include <cuda_runtime_api.h>
#include <cuda.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

struct Options {
    bool debug_exec = false;       // reference key: log every kernel invocation
    bool debug_execfail = false;   // reference key: sync + check after every launch
    bool trace_copy = false;       // reference key: log host<->device copies
    bool sync = true;              // block until results are visible before returning (BLAS semantics)
    int prefetch = 1;              // managed operands: 0 never prefetch, 1 bulk-migrate a tracked block on first use, 2 prefetch on every call
    size_t managed_threshold = 64 * 1024;   // tracker: allocations >= this go to managed memory
    int sgemm_cfg = -1;            // SGEMM tcgen05 tile configuration: -1 size-based, 0 128x128 BK32 x3, 1 128x256 BK32 x2, 2 128x256 BK16 x4
    int devices = 1;               // GPUs used by partitioned Level-3 calls
    size_t multi_gpu_min_dim = 8192;
    size_t pipeline_min_bytes = (size_t)64 << 20;   // host-resident GEMM operands above this are staged in overlapped chunks
};
extern Options g_opts;

struct Stats {
    unsigned long long hits, misses;        // reference b2c_hits / b2c_misses (runtime-mem.hpp:83,112)
    unsigned long long calls, h2d_bytes, d2h_bytes, prefetch_bytes;
};
extern Stats g_stats;

constexpr int kMaxDevices = 16;
void ensure_init();                 // idempotent, thread-safe; aborts if no CUDA device
int home_device();                  // the device single-GPU calls run on (B200BLAS_DEVICE, else the process's current device at first use)
int current_device();               // the device this thread's library state is bound to right now (home unless inside a DeviceScope)
// Switches the calling thread (CUDA current device + the library's per-thread stream / workspace / scalar slots) to `dev`
// for the scope's lifetime: how partitioned calls drive the peer GPUs from the caller's thread.
struct DeviceScope {
    explicit DeviceScope(int dev);
    ~DeviceScope();
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
private:
    int prev_;
};
bool try_init();                    // same, but returns false instead of aborting when there is no device (allocator path)
bool device_ready();                // true once ensure_init() has succeeded
int sm_count();
cudaStream_t current_stream();      // per-thread stream (or the one set by b200blas_set_stream)
void set_thread_stream(cudaStream_t s, bool external);
void finish_call();                 // the synchronous-return step (replaces call_kernel's tail)
// Per-thread helper streams (0 = host->device copies, 1 = device->host copies) and a pool of timing-less
// events, for calls that overlap staging with compute (staged_gemm.cu).  Created lazily, never destroyed.
cudaStream_t aux_stream(int which);
cudaEvent_t pooled_event(int idx);

// cudaFuncAttributeMaxDynamicSharedMemorySize, applied once per (kernel, device): thread-safe, and correct when a process
// drives several devices (the attribute belongs to the function's per-device instance).
void set_max_dynamic_smem(const void* kernel, int bytes);

bool tma_available();
bool encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, void* base, const cuuint64_t* gdim,
                       const cuuint64_t* gstride_bytes, const cuuint32_t* box, const cuuint32_t* estride,
                       CUtensorMapSwizzle swz);

// Per-thread grow-only device scratch: bump allocator reset at the start of every BLAS call.
void* ws_alloc(size_t bytes);       // 256-B aligned device memory valid until the call returns
void ws_reset();
void* pinned_scalar();              // 64 B of pinned host memory for scalar results (per thread)
// Scalar results without a stream synchronisation: armed_scalar() hands out the pinned slot pre-filled with an "unwritten"
// pattern, the reduction's finishing block stores the value straight into it (zero-copy), and wait_scalar() spins on the
// slot until every word has changed (bounded; falls back to cudaStreamSynchronize, which also surfaces kernel faults).
// Once the value is seen the call's only kernel has done its work, so finish_call() skips its own synchronisation.
void* armed_scalar(size_t bytes);
void wait_scalar(const void* slot, size_t bytes, size_t word_bytes);   // word_bytes: size of the stores the kernel makes (4 or 8)
void* device_scalar();              // 64 B of device memory for scalar results (per thread)

// Pageable host memory <-> device through pinned bounce buffers filled by a pool of host threads (host_stager.cu).  dst/src: `height`
// rows of `width` bytes.  to_device: returns once everything is queued on `stream`; from device: returns once the data is on the host.
void staged_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, bool to_device, cudaStream_t stream);
bool staged_copy_worthwhile(size_t bytes);

enum Access { ACC_IN = 1, ACC_OUT = 2, ACC_INOUT = 3 };
enum Residency { RES_DEVICE = 0, RES_MANAGED = 1, RES_HOST_PINNED = 2, RES_HOST_PAGEABLE = 3 };
Residency classify(const void* p);
void make_resident(const void* p, size_t bytes, cudaStream_t s);   // managed operand residency policy (runtime.cu)

// A column-major matrix operand (vectors are 1 x n with ld = |inc|) made device-accessible.
// Tracked-managed / device pointers are used in place (hit); host pointers are staged into the
// workspace with a 2-D copy that also re-packs them to a TMA-friendly leading dimension (miss),
// and written back on release if the access mode says so.
struct VecTag {};
class Operand {
public:
    Operand(const void* host, int64_t rows, int64_t cols, int64_t ld, size_t elem, int access);
    // BLAS vector of n elements with increment inc: occupies 1+(n-1)|inc| elements from the pointer.  A staged strided
    // vector is read as its whole extent but written back ELEMENT BY ELEMENT (a 2-D copy of n rows of one element, pitch
    // |inc| elements): the gaps between its elements are not this operand's to write -- in dswap_(n,&A[i],lda,&A[j],lda)
    // they hold the other operand (LAPACK row swaps, drot in dbdsqr), and other threads may own them.
    Operand(VecTag, const void* host, int64_t n, int64_t inc, size_t elem, int access)
        : Operand(host, n > 0 ? 1 + (n - 1) * (inc < 0 ? -inc : inc) : 0, 1, n > 0 ? 1 + (n - 1) * (inc < 0 ? -inc : inc) : 1, elem, access) {
        vec_n_ = n; vec_inc_ = inc < 0 ? -inc : inc;
    }
    void* dev() const { return dev_; }
    int64_t ld() const { return dld_; }
    void release();                 // enqueue write-back (if any); idempotent
    bool staged() const { return staged_; }
private:
    const void* host_; void* dev_; int64_t rows_, cols_, ld_, dld_; size_t elem_; int access_; bool staged_, done_;
    int64_t vec_n_ = 0, vec_inc_ = 1;
    bool pageable_ = false;
};
struct VecOperand : Operand {
    VecOperand(const void* p, int64_t n, int64_t inc, size_t elem, int access) : Operand(VecTag{}, p, n, inc, elem, access) {}
};

void log_exec(const char* routine, const char* fmt, ...);
extern thread_local const char* t_call_name;   // routine name of the call in progress (for the trace's C lines)

}  // namespace b200

// async-signal-safe formatted write (reference common.h:15-28 writef): usable inside malloc
extern "C" void b200_writef(int fd, const char* fmt, ...);
