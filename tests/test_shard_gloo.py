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


def _fake_forward(frames):
    """Stand-in for Spann3R.forward on CPU: per-sample output that depends only on that sample's own frames."""
    preds = []
    for i, f in enumerate(frames):
        m = f["img"].mean(dim=(1, 2, 3))
        preds.append({"conf": m[:, None, None] + i, "pts3d" if i == 0 else "pts3d_in_other_view": m[:, None, None, None].expand(-1, 1, 1, 3)})
    return preds, None


def _make_seqs(n, frames=3):
    g = torch.Generator().manual_seed(0)
    return [[{"img": torch.rand(1, 3, 32, 48, generator=g)} for _ in range(frames)] for _ in range(n)]


def _worker_run(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spann3r_b200 import shard
    out = shard.run_sharded(_fake_forward, _make_seqs(7), per_gpu_batch=2)
    ret[rank] = {i: [float(p["conf"].reshape(-1)[0]) for p in preds] for i, preds in out.items()}
    dist.barrier()
    dist.destroy_process_group()


def test_run_sharded_two_ranks_equals_single_process():
    """BASELINE config[2] host logic: sequences dealt to 2 ranks and advanced 2 at a time in lockstep give, per sequence,
    what a single process running them one by one gives."""
    from spann3r_b200 import shard
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_run, args=(world, _free_port(), ret), nprocs=world, join=True)
    single = shard.run_sharded(_fake_forward, _make_seqs(7), per_gpu_batch=1, rank=0, world_size=1)
    merged = {**ret[0], **ret[1]}
    assert sorted(ret[0]) == [0, 2, 4, 6] and sorted(ret[1]) == [1, 3, 5]
    assert sorted(merged) == list(range(7))
    for i in range(7):
        assert merged[i] == [float(p["conf"].reshape(-1)[0]) for p in single[i]]
        assert set(single[i][0]) == {"conf", "pts3d"} and single[i][0]["pts3d"].shape[0] == 1
    import pytest
    with pytest.raises(ValueError):
        shard.run_sharded(_fake_forward, [_make_seqs(1)[0], _make_seqs(1, frames=2)[0]], per_gpu_batch=2, rank=0, world_size=1)


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
