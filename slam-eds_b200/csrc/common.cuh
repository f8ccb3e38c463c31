// Shared host-side plumbing of libedsgpu.so (context, error handling, buffers).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost a pointer test unless a profiler is attached
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/edsgpu.h"

struct edsgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 0;
    int64_t launches = 0;
    std::vector<struct edsgpu_frames*> frames_list;  // their build streams are part of edsgpu_synchronize()
    int lm_clusters[4] = {0, 0, 0, 0};  // co-resident clusters of 1/2/4/8 tracker CTAs (0 = not queried yet)
    std::string last_error;
    // pinned staging (grown on demand) for host-facing entry points
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // device scratch (grown on demand)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
};

inline edsgpu_status edsgpu_fail(edsgpu_ctx* ctx, edsgpu_status st, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return st;
}

#define EDS_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            char buf__[512];                                                                         \
            snprintf(buf__, sizeof(buf__), "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,        \
                     cudaGetErrorString(e__));                                                       \
            return edsgpu_fail((ctx), e__ == cudaErrorMemoryAllocation ? EDSGPU_OUT_OF_MEMORY        \
                                                                        : EDSGPU_CUDA_ERROR, buf__); \
        }                                                                                            \
    } while (0)

#define EDS_REQUIRE(ctx, cond, msg)                                                  \
    do {                                                                             \
        if (!(cond)) return edsgpu_fail((ctx), EDSGPU_INVALID_ARGUMENT, (msg));      \
    } while (0)

// RAII device guard: every entry point runs on the context's device.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// NVTX range over an entry point of the C ABI (shows up in Nsight Systems timelines; SURVEY.md section 5, tracing)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
#define EDS_RANGE(name) NvtxRange nvtx_range__(name)

edsgpu_status edsgpu_ensure_pinned(edsgpu_ctx* ctx, size_t bytes);
edsgpu_status edsgpu_ensure_scratch(edsgpu_ctx* ctx, size_t bytes);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
