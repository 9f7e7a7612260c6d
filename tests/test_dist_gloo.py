"""CPU, world_size 2, gloo: the batch-sharding / final-gather host logic of the N>1 path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from psld_b200.distributed import gather_samples, max_over_ranks, rank_seed, shard, shard_bounds
    full = torch.arange(n_total * 6 * 2 * 2, dtype=torch.float32).reshape(n_total, 6, 2, 2)
    local = shard(full, rank, world)
    lo, hi = shard_bounds(n_total, rank, world)
    assert local.shape[0] == hi - lo
    got = gather_samples(local)
    assert torch.equal(got, full[:, :3]), (rank, got.shape)
    assert max_over_ranks(float(rank + 1), "cpu") == float(world)
    assert rank_seed(7, rank) == 7 + rank
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_shard_and_gather_world2(n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total), nprocs=world, join=True)


def test_shard_bounds_cover():
    from psld_b200.distributed import shard_bounds
    for n in (1, 7, 256, 2048):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
