"""CPU: the multi-GPU host logic (shard bounds, ragged gather) on a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from direct_data_driven_mpc_b200.sharding import gather_shards, shard_bounds, shard_sizes


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_u = torch.arange(total * 3 * 2, dtype=torch.float64).reshape(total, 3, 2)
        full_s = torch.arange(total, dtype=torch.int32)
        lo, hi = shard_bounds(total, world, rank)
        got_u = gather_shards(full_u[lo:hi].clone(), total)                 # all_gather
        got_s = gather_shards(full_s[lo:hi].clone(), total, dst=0)          # gather to rank 0
        ok = torch.equal(got_u, full_u)
        ok = ok and ((rank == 0 and torch.equal(got_s, full_s)) or (rank != 0 and got_s is None))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 9])
def test_gather_equals_concatenation_world2(total):
    """N-rank result == concatenation of the single-rank shard runs, ragged shards included."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
