// Device-resident objects shared between event_frame.cu and tracker.cu.
#pragma once
#include "common.cuh"
#include <vector>

struct edsgpu_lut {
    edsgpu_ctx* ctx = nullptr;
    int H = 0, W = 0;
    float* mapx = nullptr;  // nullptr = identity
    float* mapy = nullptr;
};

uint64_t edsgpu_next_uid();  // process-wide, never reused: lets caches detect recycled addresses

struct edsgpu_frames {
    edsgpu_ctx* ctx = nullptr;
    uint64_t uid = edsgpu_next_uid();
    int H = 0, W = 0, capacity = 0;
    int levels = 1;  // pyramid levels (EventFrame.cpp:342-357); per-level arrays below are indexed [level * capacity + slot]
    // Fixed-point (2^-40) brightness increments, [acc_slots][H*W].  An accumulator only lives from the scatter of a window to
    // its blur, so a large set of slots shares a RING of accumulators sized to stay resident in L2 (acc_slots < capacity): a
    // build runs in chunks of acc_slots windows, clear -> scatter -> blur, and the 8 bytes per pixel never travel to HBM.
    // Small sets keep one accumulator per slot (acc_slots == capacity), which is what the read-back entry points need.
    long long* acc = nullptr;
    int acc_slots = 0;
    std::vector<int> slot_acc;    // [capacity] accumulator that still holds the slot's window, -1 = recycled
    // blurred, un-normalised fp32 frames: one block-linear CUDA array per slot, written through a
    // surface by the blur kernel and sampled by the tracker with 2x2 texture gathers (clamp-to-edge
    // addressing = the clamped ceres::Grid2D)
    cudaArray_t* arrays = nullptr;            // [levels * capacity] host
    cudaTextureObject_t* tex = nullptr;       // [levels * capacity] host
    cudaSurfaceObject_t* surf = nullptr;      // [levels * capacity] host
    cudaSurfaceObject_t* surf_dev = nullptr;  // [levels * capacity] device copy for the blur / morphology kernels
    cudaTextureObject_t* tex_dev = nullptr;   // [levels * capacity] device copy (the morphology kernel samples level 0)
    double* partials = nullptr;   // [capacity][ntiles] per-CTA sum of squares of the kernel that is building a level
    unsigned* tickets = nullptr;  // [capacity] last-CTA election
    double* norms = nullptr;      // [levels * capacity][2] = {norm, 1/norm}
    double k0 = 0.0, k1 = 1.0;    // Gaussian taps of the last create (for read-back)
    double* exp_table = nullptr;  // [exp_E] Gaussian-in-time weight of event i of a window of exp_E events (Utils.hpp:542-546)
    int exp_E = 0;
    // host-facing create: two device staging buffers filled by a copy stream, so that the H2D copy of
    // the next batch of events overlaps the kernels still working on the current one
    void* events_dev[2] = {nullptr, nullptr};
    size_t events_bytes[2] = {0, 0};
    int stage_idx = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copied = nullptr;
    cudaEvent_t stage_free[2] = {nullptr, nullptr};
    // Event frames are BUILT on their own stream, so that the frames of the next windows can be built
    // while the context's stream is still solving the current ones (on the SMs the solve leaves free).
    // Ordering is per slot: a build waits for the last reader of the slots it overwrites, a reader
    // waits for the last build of the slots it samples.  Events come from small round-robin pools; a
    // recycled event can only over-synchronise.
    static constexpr int kEventPool = 8;
    cudaStream_t build_stream = nullptr;
    cudaEvent_t order_evt = nullptr;  // "everything the caller queued on the context's stream so far"
    cudaEvent_t built_pool[kEventPool] = {}, read_pool[kEventPool] = {};
    int built_next = 0, read_next = 0;
    std::vector<cudaEvent_t> slot_built, slot_read;  // [capacity] last build / last reader, nullptr = none
};

// make `stream` wait for the last build of slots [first, first + count)
edsgpu_status edsgpu_frames_wait_built(const edsgpu_frames* fr, int first, int count, cudaStream_t stream);
// record, after a reader of slots [first, first + count) has been queued on `stream`, that later builds must wait for it
edsgpu_status edsgpu_frames_mark_read(const edsgpu_frames* fr, int first, int count, cudaStream_t stream);
