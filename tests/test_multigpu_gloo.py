"""The N > 1 path on CPU: world_size-2 gloo processes shard the perturbation directions, all-gather their pose
records and reassemble them in direction order (the same xslam_b200.parallel code bench.py runs over NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_dirs, comps, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xslam_b200 import parallel
    mine = parallel.shard_directions(n_dirs, rank, world)
    # a fake per-rank record: real part identical on all ranks, derivative rows encode (direction, component)
    rec = torch.zeros((1 + len(mine) * comps, 16))
    rec[0] = torch.arange(16.0)
    for i, d in enumerate(mine):
        for c in range(comps):
            rec[1 + i * comps + c] = 100.0 * d + c
    full = parallel.gather_records(rec, n_dirs, comps, rank, world, dist)
    ok = full.shape == (1 + n_dirs * comps, 16) and bool((full[0] == torch.arange(16.0)).all())
    for d in range(n_dirs):
        for c in range(comps):
            ok = ok and bool((full[1 + d * comps + c] == 100.0 * d + c).all())
    ret[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("n_dirs,comps", [(6, 1), (55, 3), (5, 3)])
def test_direction_sharding_and_gather(n_dirs, comps):
    world = 2
    port = 29500 + (os.getpid() + n_dirs) % 500
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, n_dirs, comps, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_shards_partition_all_directions():
    sys.path.insert(0, ROOT)
    from xslam_b200 import parallel
    for n in (1, 6, 21, 55, 64):
        for w in (1, 2, 4, 8):
            shards = [parallel.shard_directions(n, r, w) for r in range(w)]
            assert sorted(sum(shards, [])) == list(range(n))
            assert max(len(s) for s in shards) == parallel.max_dirs_per_rank(n, w)
