"""world_size-2 gloo test of the N>1 host logic (sequence sharding + max-over-ranks timing); runs on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spann3r_b200 import shard
    mine = shard.shard_indices(7, world, rank)
    everything = shard.gather_objects(mine)
    t = shard.max_over_ranks(10.0 + 5.0 * rank)
    ret[rank] = (mine, everything, t)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert ret[0][0] == [0, 2, 4, 6] and ret[1][0] == [1, 3, 5]
    for r in range(world):
        flat = sorted(i for part in ret[r][1] for i in part)
        assert flat == list(range(7))              # every sequence exactly once
        assert ret[r][2] == 15.0                   # max over ranks, identical on both ranks


def test_single_process_identity():
    from spann3r_b200 import shard
    assert shard.shard_indices(5, 1, 0) == [0, 1, 2, 3, 4]
    assert shard.max_over_ranks(3.5) == 3.5
    assert shard.gather_objects("x") == ["x"]
