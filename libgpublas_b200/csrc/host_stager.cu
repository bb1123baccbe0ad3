// host_stager.cu -- copies between PAGEABLE host memory and the device through a ring of pinned bounce buffers that a small pool
// of host threads fills / drains.
//
// The reference's miss path -- an operand in plain malloc'd memory -- is a whole-array blocking cudaMemcpy (runtime-mem.hpp:84-112),
// and SURVEY.md section 8(f) rank 2 notes that real applications hit it for most calls.  cudaMemcpy[2D]Async from pageable memory
// is no better: the driver stages through its own bounce buffer with ONE host thread (~11 GB/s measured: DGEMM 16384^3 from
// pageable buffers ran at 15 TFLOP/s against 34 from pinned ones, profiles/r02a_bench_n1.json).  Here 8 threads copy disjoint
// column ranges of each piece into a pinned slot (host memory bandwidth, not one core's, is the limit), the slot goes to the
// device by cudaMemcpy2DAsync on the caller's stream, and the next piece is packed meanwhile (4 slots x 32 MiB).
#include "runtime.h"
#include "common.cuh"
#include "tracker.h"
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <unistd.h>

namespace b200 {
namespace {

class HostPool {
public:
    explicit HostPool(int n) {
        for (int i = 0; i < n; i++) workers_.emplace_back([this] { loop(); });
        for (auto& t : workers_) t.detach();          // live for the process; they only ever sleep on the condition variable
    }
    // fn(part) for part in [0, nparts), on the pool and the calling thread; returns when all parts are done and no worker
    // still looks at this job (the job lives on this stack frame)
    void run(int nparts, const std::function<void(int)>& fn) {
        Job job;
        job.fn = &fn; job.nparts = nparts;
        {
            std::lock_guard<std::mutex> lock(mu_);
            job_ = &job; generation_++;
        }
        cv_.notify_all();
        work(job);
        std::unique_lock<std::mutex> lock(mu_);
        job_ = nullptr;                                   // late wakers find nothing to do
        done_cv_.wait(lock, [&] { return job.done.load() >= nparts && active_ == 0; });
    }
private:
    struct Job { const std::function<void(int)>* fn = nullptr; int nparts = 0; std::atomic<int> next{0}, done{0}; };
    static void work(Job& job) {
        for (;;) {
            const int p = job.next.fetch_add(1);
            if (p >= job.nparts) break;
            (*job.fn)(p);
            job.done.fetch_add(1);
        }
    }
    void loop() {
        TrackerGuard guard;                               // never route these threads' allocations to managed memory
        unsigned long seen = 0;
        for (;;) {
            Job* job;
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&] { return generation_ != seen; });
                seen = generation_;
                job = job_;
                if (!job) continue;
                active_++;                                // the owner of the job waits for active_ == 0 before its frame goes away
            }
            work(*job);
            {
                std::lock_guard<std::mutex> lock(mu_);
                active_--;
            }
            done_cv_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    Job* job_ = nullptr;
    int active_ = 0;
    unsigned long generation_ = 0;
};

constexpr int kSlots = 4;
constexpr size_t kSlotBytes = (size_t)32 << 20;
struct Stager {
    std::mutex mu;
    HostPool* pool = nullptr;
    int nthreads = 0;
    char* slot[kSlots] = {};
    cudaEvent_t ev[kSlots] = {};
    bool busy[kSlots] = {};
} g_stager;

void stager_init() {
    Stager& st = g_stager;
    if (st.pool) return;
    int hw = (int)std::thread::hardware_concurrency();
    st.nthreads = hw >= 16 ? 8 : (hw >= 4 ? hw / 2 : 1);
    st.pool = new HostPool(st.nthreads - 1);
    for (int i = 0; i < kSlots; i++) {
        B200_CUDA(cudaMallocHost((void**)&st.slot[i], kSlotBytes));
        B200_CUDA(cudaEventCreateWithFlags(&st.ev[i], cudaEventDisableTiming));
    }
}

// rows of `width` bytes: dst[r*dpitch ..] = src[r*spitch ..] for r in [0, height), split over the pool
void parallel_copy_rows(char* dst, size_t dpitch, const char* src, size_t spitch, size_t width, size_t height) {
    Stager& st = g_stager;
    const size_t total = width * height;
    int parts = st.nthreads;
    if (total < ((size_t)1 << 20) || height < 2) parts = 1;
    if ((size_t)parts > height) parts = (int)height;
    const bool flat = dpitch == width && spitch == width;
    auto body = [&](int p) {
        if (flat) {
            const size_t b0 = total * p / parts, b1 = total * (p + 1) / parts;
            memcpy(dst + b0, src + b0, b1 - b0);
        } else {
            const size_t r0 = height * p / parts, r1 = height * (p + 1) / parts;
            for (size_t r = r0; r < r1; r++) memcpy(dst + r * dpitch, src + r * spitch, width);
        }
    };
    if (parts == 1) body(0);
    else st.pool->run(parts, body);
}

}  // namespace

bool staged_copy_worthwhile(size_t bytes) { return bytes >= ((size_t)4 << 20); }

// dst/src are "height" rows (columns of a column-major matrix) of "width" bytes.  to_device: returns once every piece has been
// queued on `stream` (the host source has been read completely); from device: returns once the data is in host memory.
void staged_copy_2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, bool to_device, cudaStream_t stream) {
    if (width == 0 || height == 0) return;
    TrackerGuard guard;
    Stager& st = g_stager;
    std::lock_guard<std::mutex> lock(st.mu);
    stager_init();
    if (width > kSlotBytes) {      // a single row larger than a slot (a flat vector copy): re-shape it into slot-sized rows
        const size_t chunk = kSlotBytes;
        size_t turn = 0;
        for (size_t r = 0; r < height; r++)
            for (size_t off = 0; off < width; off += chunk, turn++) {
                const size_t w = std::min(chunk, width - off);
                const int i = (int)(turn % kSlots);          // the slots take turns: chunk c+1 is packed while chunk c is on the wire
                if (st.busy[i]) { B200_CUDA(cudaEventSynchronize(st.ev[i])); st.busy[i] = false; }
                char* d = (char*)dst + r * dpitch + off; const char* s = (const char*)src + r * spitch + off;
                if (to_device) {
                    parallel_copy_rows(st.slot[i], w, s, w, w, 1);
                    B200_CUDA(cudaMemcpyAsync(d, st.slot[i], w, cudaMemcpyHostToDevice, stream));
                    B200_CUDA(cudaEventRecord(st.ev[i], stream)); st.busy[i] = true;
                } else {
                    B200_CUDA(cudaMemcpyAsync(st.slot[i], s, w, cudaMemcpyDeviceToHost, stream));
                    B200_CUDA(cudaEventRecord(st.ev[i], stream));
                    B200_CUDA(cudaEventSynchronize(st.ev[i]));
                    parallel_copy_rows(d, w, st.slot[i], w, w, 1);
                }
            }
        return;
    }
    const size_t rows_per = std::max<size_t>(1, kSlotBytes / width);
    const size_t npieces = (height + rows_per - 1) / rows_per;
    if (to_device) {
        for (size_t p = 0; p < npieces; p++) {
            const int i = (int)(p % kSlots);
            const size_t r0 = p * rows_per, nr = std::min(rows_per, height - r0);
            if (st.busy[i]) { B200_CUDA(cudaEventSynchronize(st.ev[i])); st.busy[i] = false; }     // the slot's previous transfer has left it
            parallel_copy_rows(st.slot[i], width, (const char*)src + r0 * spitch, spitch, width, nr);
            B200_CUDA(cudaMemcpy2DAsync((char*)dst + r0 * dpitch, dpitch, st.slot[i], width, width, nr, cudaMemcpyHostToDevice, stream));
            B200_CUDA(cudaEventRecord(st.ev[i], stream));
            st.busy[i] = true;
        }
    } else {
        // device -> slot transfers run one piece ahead of the slot -> host unpacking
        for (int i = 0; i < kSlots; i++) if (st.busy[i]) { B200_CUDA(cudaEventSynchronize(st.ev[i])); st.busy[i] = false; }
        auto issue = [&](size_t p) {
            const int i = (int)(p % kSlots);
            const size_t r0 = p * rows_per, nr = std::min(rows_per, height - r0);
            B200_CUDA(cudaMemcpy2DAsync(st.slot[i], width, (const char*)src + r0 * spitch, spitch, width, nr, cudaMemcpyDeviceToHost, stream));
            B200_CUDA(cudaEventRecord(st.ev[i], stream));
        };
        issue(0);
        for (size_t p = 0; p < npieces; p++) {
            if (p + 1 < npieces) issue(p + 1);
            const int i = (int)(p % kSlots);
            const size_t r0 = p * rows_per, nr = std::min(rows_per, height - r0);
            B200_CUDA(cudaEventSynchronize(st.ev[i]));
            parallel_copy_rows((char*)dst + r0 * dpitch, dpitch, st.slot[i], width, width, nr);
        }
    }
}

}  // namespace b200
