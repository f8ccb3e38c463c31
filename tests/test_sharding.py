"""CPU, world_size 2 over gloo: the N>1 path of bench.py (round-robin sequence sharding, the one
final all-gather, max-over-ranks timing) without a GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from edsgpu import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = shard.local_sequences(n_total, world, rank)
    local = torch.tensor([[100.0 * g + k for k in range(14)] for g in ids], dtype=torch.float64)
    full = shard.gather_states(local, n_total)
    t = shard.max_over_ranks(1.0 + rank)
    q.put((rank, ids, full.numpy(), t))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_ownership():
    assert shard.local_sequences(8, 2, 0) == [0, 2, 4, 6] and shard.local_sequences(8, 2, 1) == [1, 3, 5, 7]
    owned = sorted(sum((shard.local_sequences(64, 8, r) for r in range(8)), []))
    assert owned == list(range(64))
    assert all(len(shard.local_sequences(64, 8, r)) == 8 for r in range(8))


import pytest


@pytest.mark.parametrize("n_total", [8, 7])  # 7: the ranks own 4 and 3 sequences
def test_gather_states_world2_gloo(n_total):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.array([[100.0 * g + k for k in range(14)] for g in range(n_total)])
    for rank, ids, full, t in results:
        assert ids == list(range(rank, n_total, world))
        np.testing.assert_array_equal(full, expect)  # global order on every rank
        assert t == 2.0                              # max over ranks


def test_single_process_is_identity():
    x = torch.arange(28, dtype=torch.float64).reshape(2, 14)
    assert torch.equal(shard.gather_states(x), x)
    assert shard.max_over_ranks(3.5) == 3.5
