"""Sharding of independent sequences over ranks: sequence s runs on rank s mod G,
no collective on the data path, one all-gather of the [sequences x 14] state records at the end
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def local_sequences(num_sequences, world_size, rank):
    """Global ids of the sequences rank `rank` owns (round-robin)."""
    return list(range(rank, num_sequences, world_size))


def gather_states(local_states, num_sequences=None):
    """local_states: [n_local, 14] tensor of this rank's sequences in local order (device of the backend).
    Returns [num_sequences, 14] in GLOBAL sequence order on every rank.  When num_sequences is not a multiple of the world
    size the ranks own different numbers of sequences: every rank pads to ceil(num_sequences / world) rows for the
    collective, the padding maps to ids >= num_sequences and is cut off.  Without num_sequences all ranks must pass the
    same number of rows."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_states.clone()
    world = dist.get_world_size()
    n_local, width = local_states.shape
    n_rows = n_local if num_sequences is None else (num_sequences + world - 1) // world
    if n_local > n_rows:
        raise ValueError("gather_states: %d local rows but only %d sequences over %d ranks" % (n_local, num_sequences, world))
    send = local_states.contiguous()
    if n_local < n_rows:
        send = torch.cat([send, send.new_zeros(n_rows - n_local, width)], 0)
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send)
    stacked = torch.stack(parts, 0)            # [world, n_rows, 14]
    out = stacked.transpose(0, 1).reshape(-1, width)  # global id = local * world + rank
    return out if num_sequences is None else out[:num_sequences]


def max_over_ranks(value, device="cpu"):
    """Timing rule of the bench contract: the slowest rank defines the step time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
