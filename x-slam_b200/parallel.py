"""Multi-GPU layer: perturbation directions are independent given the (replicated, deterministic) real state, so
they are sharded across the ranks of one node and only the per-frame derivative records are exchanged
(SURVEY.md §8e).  torch.distributed is the plumbing: NCCL over NVLink on the GPUs, gloo in the CPU tests."""
import numpy as np


def shard_directions(n_dirs, rank, world):
    """Direction indices owned by `rank` (round-robin, so every rank holds ceil/floor(n_dirs/world))."""
    return list(range(rank, n_dirs, world))


def max_dirs_per_rank(n_dirs, world):
    return (n_dirs + world - 1) // world


def record_length(n_dirs_local, comps):
    """floats in one pose record: (1 + ncomp) 4x4 matrices (component 0 = real part)."""
    return (1 + n_dirs_local * comps) * 16


def gather_records(local_record, n_dirs, comps, rank, world, dist=None, out=None, send=None):
    """All-gathers the ranks' pose records (each [(1 + local_dirs*comps), 16]) and reassembles the full
    [(1 + n_dirs*comps), 16] record in direction order.  Works on torch tensors of any device; with world == 1
    it is the identity."""
    import torch
    if world == 1:
        return local_record.reshape(-1, 16)
    if dist is None:
        import torch.distributed as dist  # noqa: F811
    md = max_dirs_per_rank(n_dirs, world)
    L = record_length(md, comps)
    if send is None:
        send = torch.zeros((L,), dtype=local_record.dtype, device=local_record.device)
    if out is None:
        out = torch.zeros((world * L,), dtype=local_record.dtype, device=local_record.device)
    flat = local_record.reshape(-1)
    send[: flat.numel()].copy_(flat)
    dist.all_gather_into_tensor(out, send)  # flat output: accepted by both NCCL and gloo
    out = out.view(world, L)
    full = torch.zeros((1 + n_dirs * comps, 16), dtype=local_record.dtype, device=local_record.device)
    full[0] = out[0, :16]
    for r in range(world):
        dirs = shard_directions(n_dirs, r, world)
        rec = out[r].reshape(-1, 16)
        for i, d in enumerate(dirs):
            full[1 + d * comps: 1 + (d + 1) * comps] = rec[1 + i * comps: 1 + (i + 1) * comps]
    return full


def real_parts_agree(gathered_out, atol=0.0):
    """Replica check: every rank must have produced the same real pose (rows 0 of the gathered records)."""
    ref = gathered_out[0, :16]
    return bool(((gathered_out[:, :16] - ref).abs() <= atol).all())


# ------------------------------------------------------------------ in-library NCCL layer (csrc/comm.cpp)
def _load_torch_nccl_first():
    """The library binds NCCL with dlopen("libnccl.so.2").  In a process that will also import PyTorch, PyTorch's bundled NCCL
    must be the copy that soname resolves to (an older system NCCL loaded first would later break `import torch`), so torch is
    imported before the first NCCL call whenever it is installed."""
    try:
        import torch  # noqa: F401
    except Exception:
        pass


class Comm:
    """The library's own communicator (xs_comm: NCCL bound at run time).  The 128-byte id of rank 0 reaches the other ranks
    through whatever the launcher offers: torch.distributed (any backend), or a file for the C++ driver."""

    def __init__(self, rank, world, uid):
        import ctypes as C
        from . import _capi
        _load_torch_nccl_first()
        self.lib = _capi.load()
        self.rank, self.world = rank, world
        self.h = self.lib.xs_comm_create(rank, world, C.c_char_p(bytes(uid)))
        if not self.h:
            raise _capi.XsError("xs_comm_create: " + self.lib.xs_last_error().decode())

    @staticmethod
    def unique_id():
        import ctypes as C
        from . import _capi
        _load_torch_nccl_first()
        buf = C.create_string_buffer(128)
        _capi.check(_capi.load().xs_comm_unique_id(buf), "xs_comm_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, dist, device=None):
        """Rank 0 draws the id, torch.distributed broadcasts it (gloo: CPU tensor, nccl: tensor on `device`)."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8).clone()
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, 0)
        return cls(rank, world, bytes(t.cpu().numpy().tobytes()))

    def close(self):
        if self.h:
            self.lib.xs_comm_destroy(self.h)
            self.h = None

    __del__ = close


def shard_pairs(pairs, rank, world):
    """Second-order pairs owned by `rank` of a Hessian batch (round robin keeps every share sorted by i); every rank also
    carries all first-order components."""
    return list(pairs[rank::world])


def hessian_record_floats(n_params, n_pairs, world):
    """floats per rank in the gather: the largest record, (1 + n + ceil(m / world)) 4x4 matrices."""
    return (1 + n_params + (n_pairs + world - 1) // world) * 16


def assemble_hessian_records(gathered, n_params, pairs, world):
    """gathered: [world, record_floats] (numpy or torch).  Returns the full record [(1 + n + m), 16] in the order of `pairs`:
    the real part and the first-order components are replicated (rank 0's copy is taken), pair k lives on rank k % world."""
    g = gathered.reshape(world, -1, 16)
    full = g[0][: 1 + n_params + len(pairs)].clone() if hasattr(g, "clone") else g[0][: 1 + n_params + len(pairs)].copy()
    full = full[: 1 + n_params]
    rows = []
    for k in range(len(pairs)):
        r, local = k % world, k // world
        rows.append(g[r][1 + n_params + local])
    if hasattr(g, "clone"):
        import torch
        return torch.cat([full, torch.stack(rows)]) if rows else full
    return np.concatenate([full, np.stack(rows)]) if rows else full


def assemble_list_records(gathered, n_dirs, comps, world):
    """gathered: [world, record_floats] of a CSFD / DCSFD list run sharded round robin (direction d on rank d % world at local
    index d // world).  Returns the full record [(1 + n_dirs * comps), 16]."""
    g = np.asarray(gathered).reshape(world, -1, 16)
    rows = [g[0][0]]
    for d in range(n_dirs):
        r, local = d % world, d // world
        for c in range(comps):
            rows.append(g[r][1 + local * comps + c])
    return np.stack(rows)


# ------------------------------------------------------------------ blocked shards of a Hessian batch
def plan_hessian_shards(n_params, pairs, world, iters=20000):
    """Which rank carries which second-order pairs - and therefore which parameters - of a Hessian batch.

    A rank that holds pair (i, j) needs the first-order components of i and j as well, so round-robin shards end up carrying
    all n first-order planes on every rank (17 planes per rank for 10 parameters / 55 pairs on 8 ranks).  Here the pairs are
    dealt so that a rank's pairs share few parameters (blocks of the upper triangle): a deterministic local search on
    (max planes per rank, sum of squares), planes = |parameters touched| + |pairs|, reaches 12 planes per rank for that case.
    Every rank computes the same plan.  Returns one dict per rank:
      params: sorted global parameter ids, pair_ids: sorted indices into `pairs`, local_pairs: the pairs in local parameter indices."""
    import random
    pairs = [tuple(p) for p in pairs]
    m = len(pairs)
    if world == 1:
        shards = [list(range(m))]
    else:
        order = sorted(range(m), key=lambda k: (max(pairs[k]), min(pairs[k])))
        shards = [[] for _ in range(world)]
        for t, k in enumerate(order):
            shards[t * world // m].append(k)

        def score(sh):
            c = [len({q for k in s_ for q in pairs[k]}) + len(s_) for s_ in sh]
            return max(c), sum(x * x for x in c)

        rnd = random.Random(12345)
        best = score(shards)
        for _ in range(iters):
            a, b = rnd.randrange(world), rnd.randrange(world)
            if a == b or not shards[a]:
                continue
            ia = rnd.randrange(len(shards[a]))
            if shards[b] and rnd.random() < 0.5:
                ib = rnd.randrange(len(shards[b]))
                shards[a][ia], shards[b][ib] = shards[b][ib], shards[a][ia]
                sc = score(shards)
                if sc <= best:
                    best = sc
                else:
                    shards[a][ia], shards[b][ib] = shards[b][ib], shards[a][ia]
            else:
                k = shards[a].pop(ia)
                shards[b].append(k)
                sc = score(shards)
                if sc <= best:
                    best = sc
                else:
                    shards[b].pop()
                    shards[a].insert(ia, k)
    # blocks pay off when they take planes off the heaviest rank; where they do not (2 ranks: 37 against 38 planes) the
    # round-robin shards are kept - they spread the pairs of every parameter, and with them the costlier intrinsic pairs, evenly
    if world > 1:
        rr = [list(range(r, m, world)) for r in range(world)]
        planes = lambda sh: max(len({q for k in s_ for q in pairs[k]}) + len(s_) for s_ in sh)
        if planes(shards) > 0.9 * planes(rr):
            shards = rr
    plan = []
    for s_ in shards:
        ids = sorted(s_, key=lambda k: pairs[k])
        params = sorted({q for k in ids for q in pairs[k]})
        loc = {g: i for i, g in enumerate(params)}
        plan.append({"params": params, "pair_ids": ids, "local_pairs": [(loc[pairs[k][0]], loc[pairs[k][1]]) for k in ids]})
    # every parameter's first-order component must come from somewhere: parameters no pair touches go to the lightest rank
    seen = {q for sh in plan for q in sh["params"]}
    for g in range(n_params):
        if g not in seen:
            sh = min(plan, key=lambda d: len(d["params"]) + len(d["pair_ids"]))
            sh["params"] = sorted(sh["params"] + [g])
            loc = {q: i for i, q in enumerate(sh["params"])}
            sh["local_pairs"] = [(loc[pairs[k][0]], loc[pairs[k][1]]) for k in sh["pair_ids"]]
    return plan


def planned_record_floats(plan):
    """floats per rank in the gather: the largest record of the plan, (1 + n_local + m_local) 4x4 matrices."""
    return max(1 + len(sh["params"]) + len(sh["pair_ids"]) for sh in plan) * 16


def assemble_planned_records(gathered, plan, n_params, n_pairs):
    """gathered: [world, record_floats] (numpy).  Returns the full record [(1 + n + m), 16]: the real part from rank 0, the
    first-order component of a parameter from the lowest rank that carries it, pair k from the rank that owns it."""
    g = np.asarray(gathered).reshape(len(plan), -1, 16)
    full = np.zeros((1 + n_params + n_pairs, 16), g.dtype)
    full[0] = g[0][0]
    done = set()
    for r, sh in enumerate(plan):
        nl = len(sh["params"])
        for i, q in enumerate(sh["params"]):
            if q not in done:
                full[1 + q] = g[r][1 + i]
                done.add(q)
        for i, k in enumerate(sh["pair_ids"]):
            full[1 + n_params + k] = g[r][1 + nl + i]
    return full
