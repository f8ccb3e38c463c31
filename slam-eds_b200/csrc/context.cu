// Context lifetime for libedsgpu.so.  No CPU fallback: creation fails without a CUDA device.
#include <atomic>

#include "common.cuh"
#include "frames.cuh"

extern "C" {

const char* edsgpu_version(void) { return "edsgpu 0.1 (sm_100a)"; }

edsgpu_status edsgpu_create(int device, void* stream, edsgpu_ctx** out) {
    if (!out) return EDSGPU_INVALID_ARGUMENT;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return EDSGPU_CUDA_ERROR;
    edsgpu_ctx* ctx = new edsgpu_ctx();
    ctx->device = device;
    DeviceGuard g(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return EDSGPU_CUDA_ERROR; }
    ctx->num_sms = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return EDSGPU_CUDA_ERROR; }
        ctx->own_stream = true;
    }
    *out = ctx;
    return EDSGPU_OK;
}

void edsgpu_destroy(edsgpu_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* edsgpu_last_error_string(const edsgpu_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

edsgpu_status edsgpu_synchronize(edsgpu_ctx* ctx) {
    if (!ctx) return EDSGPU_INVALID_ARGUMENT;
    DeviceGuard g(ctx->device);
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (edsgpu_frames* fr : ctx->frames_list) EDS_CUDA(ctx, cudaStreamSynchronize(fr->build_stream));  // event frames are built on their own stream
    return EDSGPU_OK;
}

int64_t edsgpu_launch_count(const edsgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

uint64_t edsgpu_next_uid() {
    static std::atomic<uint64_t> counter{1};
    return counter.fetch_add(1);
}

edsgpu_status edsgpu_ensure_pinned(edsgpu_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->pinned_bytes) return EDSGPU_OK;
    // the old block may still be the source/target of an in-flight async copy
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 0;
    size_t want = align_up(bytes + bytes / 2, 4096);
    EDS_CUDA(ctx, cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_bytes = want;
    return EDSGPU_OK;
}

edsgpu_status edsgpu_ensure_scratch(edsgpu_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return EDSGPU_OK;
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->scratch) cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    EDS_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return EDSGPU_OK;
}
