"""Sharding of independent sequences over ranks: sequence s runs on rank s mod G,
no collective on the data path, one all-gather of the [sequences x 14] state records at the end
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def local_sequences(num_sequences, world_size, rank):
    """Global ids of the sequences rank `rank` owns (round-robin)."""
    return list(range(rank, num_sequences, world_size))


def gather_states(local_states, num_sequences=None):
    """local_states: [n_local, 14] tensor (same n_local on every rank, device of the backend).
    Returns [world*n_local, 14] in GLOBAL sequence order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_states.clone()
    world = dist.get_world_size()
    parts = [torch.empty_like(local_states) for _ in range(world)]
    dist.all_gather(parts, local_states.contiguous())
    stacked = torch.stack(parts, 0)            # [world, n_local, 14]
    out = stacked.transpose(0, 1).reshape(-1, local_states.shape[1])  # global id = local * world + rank
    return out if num_sequences is None else out[:num_sequences]


def max_over_ranks(value, device="cpu"):
    """Timing rule of the bench contract: the slowest rank defines the step time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
