// The final gather of the sharded batch in the C++ host layer: one ncclAllGather of the state records
// over NVLink / NVSwitch, nothing else.  Built into libedsgpu_nccl.so (see include/edsgpu_nccl.h).
#include <nccl.h>

#include "../../include/edsgpu_nccl.h"
#include "common.cuh"

struct edsgpu_comm {
    edsgpu_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    bool owned = false;
    int world = 1, rank = 0;
    double* buf = nullptr;  // device: send (rows x 14) | recv (world x rows x 14) | out (world x rows x 14)
    size_t buf_rows = 0;
};

namespace {

constexpr int W14 = 14;

// recv[r][l][:] -> out[l * world + r][:] for ids below num_sequences
__global__ void reorder_states_kernel(const double* __restrict__ recv, int world, int rows, int num_sequences, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_sequences * W14) return;
    const int s = i / W14, k = i % W14;
    out[i] = recv[((size_t)(s % world) * rows + s / world) * W14 + k];
}

__global__ void pad_states_kernel(const double* __restrict__ local, int n_local, int rows, double* __restrict__ send) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * W14) send[i] = (i < n_local * W14) ? local[i] : 0.0;
}

edsgpu_status nccl_fail(edsgpu_ctx* ctx, ncclResult_t r, const char* what) {
    return edsgpu_fail(ctx, EDSGPU_CUDA_ERROR, std::string(what) + ": " + ncclGetErrorString(r));
}

}  // namespace

extern "C" {

edsgpu_status edsgpu_comm_unique_id(char id_out[EDSGPU_COMM_ID_BYTES]) {
    static_assert(sizeof(ncclUniqueId) == EDSGPU_COMM_ID_BYTES, "ncclUniqueId size");
    if (!id_out) return EDSGPU_INVALID_ARGUMENT;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return EDSGPU_CUDA_ERROR;
    memcpy(id_out, &id, sizeof(id));
    return EDSGPU_OK;
}

edsgpu_status edsgpu_comm_create(edsgpu_ctx* ctx, int world_size, int rank, const char id[EDSGPU_COMM_ID_BYTES], edsgpu_comm** out) {
    if (!ctx || !out || !id) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, world_size >= 1 && rank >= 0 && rank < world_size, "comm_create: bad world size / rank");
    DeviceGuard g(ctx->device);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t c = nullptr;
    ncclResult_t r = ncclCommInitRank(&c, world_size, uid, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclCommInitRank");
    edsgpu_comm* h = new edsgpu_comm();
    h->ctx = ctx; h->comm = c; h->owned = true; h->world = world_size; h->rank = rank;
    *out = h;
    return EDSGPU_OK;
}

edsgpu_status edsgpu_comm_adopt(edsgpu_ctx* ctx, void* nccl_comm, int world_size, int rank, edsgpu_comm** out) {
    if (!ctx || !out || !nccl_comm) return EDSGPU_INVALID_ARGUMENT;
    EDS_REQUIRE(ctx, world_size >= 1 && rank >= 0 && rank < world_size, "comm_adopt: bad world size / rank");
    edsgpu_comm* h = new edsgpu_comm();
    h->ctx = ctx; h->comm = (ncclComm_t)nccl_comm; h->owned = false; h->world = world_size; h->rank = rank;
    *out = h;
    return EDSGPU_OK;
}

void edsgpu_comm_destroy(edsgpu_comm* comm) {
    if (!comm) return;
    DeviceGuard g(comm->ctx->device);
    cudaStreamSynchronize(comm->ctx->stream);
    if (comm->buf) cudaFree(comm->buf);
    if (comm->owned && comm->comm) ncclCommDestroy(comm->comm);
    delete comm;
}

edsgpu_status edsgpu_gather_states_nccl(edsgpu_comm* comm, const double* local_states_dev, int n_local, int num_sequences,
                                        double* global_states_dev) {
    if (!comm) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = comm->ctx;
    EDS_REQUIRE(ctx, local_states_dev && global_states_dev && num_sequences > 0, "gather_states_nccl: bad arguments");
    const int world = comm->world;
    const int mine = (num_sequences - comm->rank + world - 1) / world;  // ids rank, rank + world, ... below num_sequences
    EDS_REQUIRE(ctx, n_local == mine, "gather_states_nccl: n_local is not this rank's share of num_sequences (round-robin dealing)");
    DeviceGuard g(ctx->device);
    const size_t rows = (size_t)(num_sequences + world - 1) / world;
    if (comm->buf_rows < rows) {
        EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (comm->buf) cudaFree(comm->buf);
        comm->buf = nullptr; comm->buf_rows = 0;
        EDS_CUDA(ctx, cudaMalloc(&comm->buf, sizeof(double) * W14 * rows * (1 + (size_t)world)));
        comm->buf_rows = rows;
    }
    double* send = comm->buf;
    double* recv = comm->buf + W14 * rows;
    const int n = (int)rows * W14;
    pad_states_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(local_states_dev, n_local, (int)rows, send);
    EDS_CUDA(ctx, cudaGetLastError());
    ncclResult_t r = ncclAllGather(send, recv, (size_t)n, ncclDouble, comm->comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, r, "ncclAllGather");
    const int m = num_sequences * W14;
    reorder_states_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(recv, world, (int)rows, num_sequences, global_states_dev);
    EDS_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return EDSGPU_OK;
}

edsgpu_status edsgpu_batch_gather_states_nccl(edsgpu_batch* batch, edsgpu_comm* comm, int num_sequences, double* global_states_host) {
    if (!batch || !comm || !global_states_host) return EDSGPU_INVALID_ARGUMENT;
    edsgpu_ctx* ctx = comm->ctx;
    int n_local = 0;
    edsgpu_status st = edsgpu_batch_count(batch, &n_local);
    if (st != EDSGPU_OK) return st;
    DeviceGuard g(ctx->device);
    const size_t bytes = sizeof(double) * W14 * ((size_t)n_local + (size_t)num_sequences);
    st = edsgpu_ensure_scratch(ctx, bytes);
    if (st != EDSGPU_OK) return st;
    double* local = (double*)ctx->scratch;
    double* global = local + (size_t)W14 * n_local;
    st = edsgpu_batch_pack_states_dev(batch, local);
    if (st == EDSGPU_OK) st = edsgpu_gather_states_nccl(comm, local, n_local, num_sequences, global);
    if (st != EDSGPU_OK) return st;
    EDS_CUDA(ctx, cudaMemcpyAsync(global_states_host, global, sizeof(double) * W14 * (size_t)num_sequences, cudaMemcpyDeviceToHost, ctx->stream));
    EDS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return EDSGPU_OK;
}

}  // extern "C"
