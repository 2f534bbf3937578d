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


def _planned_worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from xslam_b200 import parallel
    pairs = [(i, j) for i in range(n) for j in range(i, n)]
    plan = parallel.plan_hessian_shards(n, pairs, world)  # every rank computes the same plan
    mine = plan[rank]
    L = parallel.planned_record_floats(plan)
    # fake record of this rank: real row, one row per local parameter (value 1000 + global id), one per local pair (2000 + pair id)
    rec = torch.zeros(L)
    rows = rec.view(-1, 16)
    rows[0] = torch.arange(16.0)
    for i, q in enumerate(mine["params"]):
        rows[1 + i] = 1000.0 + q
    for i, k in enumerate(mine["pair_ids"]):
        rows[1 + len(mine["params"]) + i] = 2000.0 + k
    out = torch.zeros(world * L)
    dist.all_gather_into_tensor(out, rec)
    full = parallel.assemble_planned_records(out.numpy().reshape(world, L), plan, n, len(pairs))
    ok = full.shape == (1 + n + len(pairs), 16) and bool((full[0] == np.arange(16.0)).all())
    ok = ok and all((full[1 + q] == 1000.0 + q).all() for q in range(n))
    ok = ok and all((full[1 + n + k] == 2000.0 + k).all() for k in range(len(pairs)))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_blocked_hessian_shards_gather_over_gloo():
    world = 2
    port = 29500 + (os.getpid() + 77) % 500
    ret = mp.Manager().dict()
    mp.spawn(_planned_worker, args=(world, port, 10, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world))


def test_blocked_hessian_shards_cover_every_pair_once():
    sys.path.insert(0, ROOT)
    from xslam_b200 import parallel
    for n in (3, 6, 10):
        pairs = [(i, j) for i in range(n) for j in range(i, n)]
        for w in (1, 2, 4, 8):
            plan = parallel.plan_hessian_shards(n, pairs, w, iters=4000)
            assert sorted(k for sh in plan for k in sh["pair_ids"]) == list(range(len(pairs)))
            assert {q for sh in plan for q in sh["params"]} == set(range(n))
            for sh in plan:  # the local pairs address the rank's own parameter list, sorted by the first parameter
                assert [(sh["params"][a], sh["params"][b]) for a, b in sh["local_pairs"]] == [pairs[k] for k in sh["pair_ids"]]
                assert sh["local_pairs"] == sorted(sh["local_pairs"])
            planes = max(len(sh["params"]) + len(sh["pair_ids"]) for sh in plan)
            assert planes <= n + (len(pairs) + w - 1) // w  # never more planes per rank than round-robin shards
    plan = parallel.plan_hessian_shards(10, [(i, j) for i in range(10) for j in range(i, 10)], 8)
    assert max(len(sh["params"]) + len(sh["pair_ids"]) for sh in plan) <= 12


def test_round_robin_is_kept_where_blocks_do_not_pay():
    """2 ranks, 10 parameters: blocks would leave 37 planes on the heavier rank against 38, so the plan stays round robin (which
    spreads the pairs of every parameter evenly); at 4 and 8 ranks the blocks are taken."""
    sys.path.insert(0, ROOT)
    from xslam_b200 import parallel
    pairs = [(i, j) for i in range(10) for j in range(i, 10)]
    p2 = parallel.plan_hessian_shards(10, pairs, 2)
    assert [sh["pair_ids"] for sh in p2] == [list(range(0, 55, 2)), list(range(1, 55, 2))]
    for w, planes in ((4, 21), (8, 12)):
        plan = parallel.plan_hessian_shards(10, pairs, w)
        assert max(len(sh["params"]) + len(sh["pair_ids"]) for sh in plan) <= planes
