"""GPU, two ranks on two devices: the final gather of a sharded batch through the C ABI (libedsgpu_nccl.so,
include/edsgpu_nccl.h) -- one ncclAllGather over NVLink, global sequence order on every rank, uneven shares padded.
Skipped on a single-GPU box (NCCL refuses two ranks on one device); tests/test_sharding.py covers the same dealing
logic with gloo on CPU."""
import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, n_total, q_id, q_out):
    import edsgpu
    torch.cuda.set_device(rank)
    ctx = edsgpu.Context(rank)
    if rank == 0:
        uid = edsgpu.Comm.unique_id()
        for _ in range(world - 1):
            q_id.put(uid)
    else:
        uid = q_id.get(timeout=120)
    comm = edsgpu.Comm(ctx, world, rank, uid)
    ids = list(range(rank, n_total, world))
    local = torch.tensor([[100.0 * g + k for k in range(14)] for g in ids], dtype=torch.float64, device="cuda:%d" % rank)
    out = torch.full((n_total, 14), -1.0, dtype=torch.float64, device="cuda:%d" % rank)
    torch.cuda.synchronize()
    comm.gather_states_dev(local.data_ptr(), len(ids), n_total, out.data_ptr())
    ctx.synchronize()
    q_out.put((rank, out.cpu().numpy()))
    comm.close()
    ctx.close()


@pytest.mark.parametrize("n_total", [8, 7])
def test_gather_states_over_nccl_two_ranks(n_total):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    mpc = mp.get_context("spawn")
    q_id, q_out = mpc.Queue(), mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, n_total, q_id, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q_out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.array([[100.0 * g + k for k in range(14)] for g in range(n_total)])
    for rank, full in results:
        np.testing.assert_array_equal(full, expect)


def test_single_rank_gather_is_a_copy(gpu_ctx):
    """world = 1: the communicator works on one device too and the gather is the identity (exercised by the 1-GPU run)."""
    import edsgpu
    comm = edsgpu.Comm(gpu_ctx, 1, 0, edsgpu.Comm.unique_id())
    local = torch.arange(5 * 14, dtype=torch.float64, device="cuda:0").reshape(5, 14)
    out = torch.zeros_like(local)
    torch.cuda.synchronize()
    comm.gather_states_dev(local.data_ptr(), 5, 5, out.data_ptr())
    gpu_ctx.synchronize()
    assert torch.equal(out, local)
    with pytest.raises(edsgpu.EdsGpuError):
        comm.gather_states_dev(local.data_ptr(), 4, 5, out.data_ptr())  # not this rank's share
    comm.close()
