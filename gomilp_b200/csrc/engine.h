// engine.h — internals shared by the translation units of libgomilp_b200.so (engine.cu: C ABI, staging, launch;
// bnb_device.cu: device-side node check / scan and the multi-GPU wavefront). Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gomilp_b200.h"
#include "kernels.h"

namespace gm_engine {

constexpr int kMaxDevices = 16;

struct Root {
    int device = 0;
    double *c = nullptr, *A = nullptr, *b = nullptr;
    int m0 = 0, n0 = 0;
    // warm-start state: final bases / inverses of the previous wave, kept in HBM for the children
    double* prev_bi = nullptr;
    long long* prev_basis = nullptr;
    int64_t prev_nodes = 0;
    int prev_m = 0;
};

struct StreamEvents {
    cudaStream_t s = nullptr;
    cudaEvent_t e[6] = {};
    int* h_counts = nullptr;  // 64 pinned ints: arrival counters of a streamed batch (run_host_batch_streamed)
    // scratch of the device-side B&B scheduler (bnb_device.cu), kept with the stream across calls: wave buffers only
    // grow, so a second search of a similar size allocates nothing (stream-ordered allocation of GBs is not free)
    void* arena = nullptr;
    void (*arena_free)(void*) = nullptr;
    cudaError_t init();
    ~StreamEvents();
};

struct DeviceCtx {
    std::mutex mu;
    bool ready = false;
    bool coop_ok = false;
    int id = -1;
    int sms = 0;
    size_t smem_optin = 0;
    std::vector<StreamEvents*> pool;  // idle streams (+ events), reused across calls
};

struct StreamLease {
    DeviceCtx* dev;
    StreamEvents* se = nullptr;
    cudaError_t err = cudaSuccess;
    explicit StreamLease(DeviceCtx* d);
    ~StreamLease();
    StreamLease(const StreamLease&) = delete;
    StreamLease& operator=(const StreamLease&) = delete;
};

struct Engine {
    std::mutex mu;
    DeviceCtx dev[kMaxDevices];
    int default_device = -1;
    gm_options opt{0, 0, 0, 0, 0, 0, 0, 0};
    std::map<gm_root_t, Root> roots;
    gm_root_t next_root = 1;
};

struct TraceReq {
    bool armed = false, in_flight = false;
    int64_t lp = 0, cap = 0;
    int* d_buf = nullptr;
    cudaStream_t stream = nullptr;
    std::vector<int32_t> rows;
    // gm_profile_arm: leader clock cycles per activity of the cooperative tier
    bool prof_armed = false, prof_in_flight = false;
    long long* d_prof = nullptr;
    int64_t prof_count = 0;
    cudaStream_t prof_stream = nullptr;
    std::vector<long long> prof;
};

extern Engine g;
extern thread_local int t_device;
extern thread_local std::string t_err;
extern thread_local gm_timing t_timing;
extern thread_local TraceReq t_trace;
extern thread_local int t_robust;  // > 0: waves launched by this thread solve with BatchParams::robust (GM_BNB_ROBUST)
struct RobustScope {
    explicit RobustScope(bool on) : on_(on) { if (on_) ++t_robust; }
    ~RobustScope() { if (on_) --t_robust; }
    bool on_;
};

int fail(cudaError_t e, const char* what);
#define CK(call)                                              \
    do {                                                      \
        cudaError_t e_ = (call);                              \
        if (e_ != cudaSuccess) return gm_engine::fail(e_, #call); \
    } while (0)

int current_device(DeviceCtx** out);
int device_of_root(gm_root_t h, Root* out, DeviceCtx** dev);
int launch_wave(DeviceCtx& d, gm::BatchParams P, cudaStream_t stream, cudaEvent_t ev0, cudaEvent_t ev1, gm_timing* tm);
void finish_trace();
float ms(cudaEvent_t a, cudaEvent_t b);

cudaError_t gm_bnb_kernels_prepare();  // bnb_device.cu

}  // namespace gm_engine
