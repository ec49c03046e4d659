"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in gen6d_b200/dist.py: shard
partitioning, rank-major gathers, and the exactness of cross-shard InstanceNorm statistics."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, fn_name, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from gen6d_b200.dist import Comm
        globals()[fn_name](Comm(), rank, world)
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def run2(fn_name, world=2):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == 'ok' for r in res), res


def _check_comm(comm, rank, world):
    assert comm.world == world and comm.rank == rank
    t = torch.full((3,), float(rank + 1), dtype=torch.float64)
    assert comm.all_reduce_sum(t).tolist() == [3.0, 3.0, 3.0]
    g = comm.all_gather_cat(torch.arange(4).reshape(2, 2) + 10 * rank, dim=0)
    assert g.tolist() == [[0, 1], [2, 3], [10, 11], [12, 13]]          # rank-major
    assert comm.shard_range(8) == (4 * rank, 4 * rank + 4)
    with pytest.raises(ValueError):
        comm.shard_range(7)


def _check_exact_instnorm(comm, rank, world):
    """Reference-sharded InstanceNorm3d: all-reduced fp64 moments == statistics of the whole tensor
    (what the selector's towers do between convs), unlike per-shard normalisation."""
    from gen6d_b200.dist import merge_moments
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 16, 8, 5, 5, generator=g) * 3 + 1            # [1, C, S, h, w], S = 8 slices
    b, e = comm.shard_range(8)
    mine = x[:, :, b:e]
    flat = mine.double().permute(1, 0, 2, 3, 4).reshape(16, -1)
    s1, s2 = comm.all_reduce_sum(flat.sum(1)), comm.all_reduce_sum((flat * flat).sum(1))
    scale, shift = merge_moments(s1, s2, 8 * 25)
    got = mine * scale.float().view(1, 16, 1, 1, 1) + shift.float().view(1, 16, 1, 1, 1)
    want = F.instance_norm(x)[:, :, b:e]
    assert torch.allclose(got, want, atol=1e-5)
    approx = F.instance_norm(mine)                                   # the "single all-gather" shortcut
    assert (approx - want).abs().max() > 1e-3                        # ... is measurably different


def _check_pose_shard(comm, rank, world):
    from gen6d_b200.dist import pose_shard
    poses = torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3)
    out = pose_shard(lambda p: torch.cat([p * 2, p.sum(1, keepdim=True)], 1), [poses], comm)
    assert torch.equal(out, torch.cat([poses * 2, poses.sum(1, keepdim=True)], 1))


def test_comm_primitives_world2():
    run2('_check_comm')


def test_exact_cross_shard_instance_norm_world2():
    run2('_check_exact_instnorm')


def test_pose_shard_gather_order_world2():
    run2('_check_pose_shard')


def test_local_comm_is_identity():
    from gen6d_b200.network.selector import LocalComm
    c = LocalComm()
    t = torch.ones(2)
    assert c.all_reduce_sum(t) is t and c.all_gather_cat(t) is t and c.shard_range(6) == (0, 6)
