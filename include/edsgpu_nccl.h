/* edsgpu_nccl.h -- the one collective of the path, in the C++ host layer.
 *
 * Sequences are independent: ranks share nothing while they track.  At the end the [sequences x 14] state records
 * (px(3) qx(4) vx(6) tau, Tracker.hpp:47-49 + config.loss_params) are gathered with ONE ncclAllGather over NVLink /
 * NVSwitch.  These entry points live in their own library (libedsgpu_nccl.so, linked against libnccl.so.2) so that
 * libedsgpu.so itself has no NCCL dependency; a single-GPU user never loads it.  The reference has no multi-process layer:
 * there is nothing it replaces, it is what the Rock task running N trackers would call after its last window.
 */
#ifndef EDSGPU_NCCL_H
#define EDSGPU_NCCL_H

#include "edsgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define EDSGPU_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */

typedef struct edsgpu_comm edsgpu_comm; /* one rank's end of an NCCL communicator, bound to a context (device + stream) */

/* Rank 0 makes the id (ncclGetUniqueId) and hands the 128 bytes to the other ranks by any means (MPI, a file, a socket). */
edsgpu_status edsgpu_comm_unique_id(char id_out[EDSGPU_COMM_ID_BYTES]);
/* ncclCommInitRank on the context's device; collective: every rank calls it with the same id. */
edsgpu_status edsgpu_comm_create(edsgpu_ctx* ctx, int world_size, int rank, const char id[EDSGPU_COMM_ID_BYTES], edsgpu_comm** out);
/* Adopt a communicator the application already has (nccl_comm is an ncclComm_t); it is not destroyed with the handle. */
edsgpu_status edsgpu_comm_adopt(edsgpu_ctx* ctx, void* nccl_comm, int world_size, int rank, edsgpu_comm** out);
void edsgpu_comm_destroy(edsgpu_comm* comm);

/* Sequence s lives on rank s mod world_size as local row s / world_size (round-robin dealing).  local_states_dev:
 * [n_local x 14] doubles on this rank's device, n_local = number of ids s < num_sequences with s mod world == rank.
 * global_states_dev: [num_sequences x 14] in global sequence order, on every rank.  Asynchronous on the context's stream
 * (pad, ncclAllGather, reorder kernel). */
edsgpu_status edsgpu_gather_states_nccl(edsgpu_comm* comm, const double* local_states_dev, int n_local, int num_sequences,
                                        double* global_states_dev);
/* The same for the trackers of a batch (edsgpu_batch_pack_states_dev + gather), result copied to the host; synchronises. */
edsgpu_status edsgpu_batch_gather_states_nccl(edsgpu_batch* batch, edsgpu_comm* comm, int num_sequences, double* global_states_host);

#ifdef __cplusplus
}
#endif
#endif
