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
