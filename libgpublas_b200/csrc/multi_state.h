// multi_state.h -- devices, streams and panel buffers shared by the partitioned Level-3 drivers (multi_gemm.cu: ?gemm_ and the
// Cholesky workload; multi_level3.cu: ?syrk_, ?trsm_, ?trmm_).  One process, one calling thread drives every device; all
// cross-device ordering is CUDA events.
#pragma once
#include "runtime.h"
#include <mutex>
#include <vector>

namespace b200 {

struct MgDev {
    int id = -1;
    cudaStream_t comp = nullptr, in = nullptr, out = nullptr, push = nullptr;
    cudaStream_t fwd[kMaxDevices] = {};           // one forwarding stream per destination slot (Cholesky's ring and column traffic)
    char* panelA = nullptr; size_t capA = 0;
    char* panelB = nullptr; size_t capB = 0;
    char* ctile = nullptr; size_t capC = 0;
    uint32_t* flags = nullptr;                    // [0,2048): A row-groups, [2048,4096): B column bands
    uint32_t* consts = nullptr;                   // consts[v] = v (device copy, for flag writes into a peer)
    std::vector<cudaEvent_t> events; size_t next_event = 0;
    cudaEvent_t done = nullptr;
};
struct MgState {
    std::mutex mu;                                // one partitioned call at a time
    int ndev = 0; bool ready = false, failed = false;
    MgDev dev[kMaxDevices];
    uint32_t* host_consts = nullptr;              // pinned: flag writes that follow a host->device copy
    uint32_t epoch = 0;
};
extern MgState g_mg;

cudaEvent_t next_event(MgDev& d);                 // from the device's pool; the pool restarts at every partitioned call
void ensure_cap(char** p, size_t* cap, size_t need);
bool mg_init(int ndev);                           // brings up `ndev` devices (home first) with peer access between every pair
cudaStream_t push_stream(MgDev& d);               // the device's one ordered stream for pushes to its peers

}  // namespace b200
