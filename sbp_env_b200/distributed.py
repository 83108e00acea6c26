"""Multi-GPU sharding of the naturally parallel workloads.

One process per GPU (``torch.distributed``; NCCL on the GPU box, gloo in the CPU
tests).  PRM candidate-edge batches and independent query sets are partitioned
across ranks with NO collective on the data path; the only exchange is the
all-gather that assembles the per-rank visibility masks / neighbour blocks into the
roadmap adjacency.  Sequential single-tree planners do not shard ("replicas only").
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced slice [first, first+count) of n units for this rank."""
    base, rem = divmod(int(n), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def shard_by_cost(costs, world: int):
    """Partition units into `world` contiguous chunks of near-equal total cost (e.g.
    edges weighted by their pixel count).  Returns the `world + 1` chunk boundaries."""
    c = np.asarray(costs, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    return np.concatenate([[0], cuts, [len(c)]]).astype(np.int64)


def all_gather_rows(local, counts, group=None):
    """Assemble row blocks of unequal length from all ranks (rank order) on every rank.

    ``local`` is this rank's [count_r, ...] tensor; ``counts`` the per-rank row counts.
    Blocks are padded to the maximum count so that one fixed-stride all-gather
    (NCCL ``ncclAllGather`` / gloo) moves them."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    mx = int(max(counts))
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * mx: r * mx + int(counts[r])] for r in range(world)]
    return torch.cat(parts, 0)


def pack_adjacency(nbr, vis):
    """One int32 per candidate edge: ``(nbr + 1) * 2 + vis`` (nbr in [-1, 2^30)), so that the
    neighbour block and its visibility mask travel in ONE collective of 4 bytes per edge
    instead of two of 8 + 1."""
    import torch

    return (((nbr + 1) << 1) | vis.to(nbr.dtype)).to(torch.int32)


def unpack_adjacency(code):
    """Inverse of :func:`pack_adjacency`: ``(nbr int64, vis bool)``."""
    import torch

    return (code >> 1).to(torch.int64) - 1, (code & 1).bool()


def all_gather_adjacency(nbr, vis, counts=None, group=None, unpack=True):
    """All-gather of the per-rank (nbr, vis) row blocks as one packed int32 collective.
    ``counts`` (rows per rank) is only needed when the blocks differ in length.  With
    ``unpack=False`` the packed int32 adjacency is returned as is (decode it where it is
    consumed with :func:`unpack_adjacency`)."""
    import torch
    import torch.distributed as dist

    code = pack_adjacency(nbr, vis)
    if counts is None or len(set(int(c) for c in counts)) == 1:
        world = dist.get_world_size(group)
        out = torch.empty((world * code.shape[0],) + tuple(code.shape[1:]), dtype=code.dtype,
                          device=code.device)
        dist.all_gather_into_tensor(out, code, group=group)
    else:
        out = all_gather_rows(code, counts, group=group)
    return unpack_adjacency(out) if unpack else out


def sharded_visible_batch(visible_fn, P, Q, group=None, costs=None):
    """Every rank holds the full edge list (P, Q); each checks its chunk with
    ``visible_fn`` and the masks are all-gathered, so every rank returns the full mask.

    ``visible_fn(P_chunk, Q_chunk) -> bool tensor`` is ``cc.visible_batch`` on the GPU
    box; the CPU (gloo) tests plug in the oracle to exercise the partition/assembly."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = P.shape[0]
    if costs is None:
        bounds = np.array([shard_range(n, r, world)[0] for r in range(world)] + [n], dtype=np.int64)
    else:
        bounds = shard_by_cost(costs, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    local = visible_fn(P[lo:hi], Q[lo:hi])
    if not torch.is_tensor(local):
        local = torch.as_tensor(np.asarray(local))
    local = local.to(torch.uint8)
    counts = (bounds[1:] - bounds[:-1]).tolist()
    return all_gather_rows(local, counts, group=group).bool()


def sharded_prm_connect(connect_fn, n_nodes: int, group=None):
    """Rows (nodes) are split across ranks; ``connect_fn(first, count) -> (nbr, vis)``
    is :func:`sbp_env_b200.planning.prm_connect_knn` restricted to this rank's rows.
    Returns the all-gathered ``nbr`` [n][k] and ``vis`` [n][k] on every rank."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = shard_range(n_nodes, rank, world)
    nbr, vis = connect_fn(first, count)
    counts = [shard_range(n_nodes, r, world)[1] for r in range(world)]
    return all_gather_adjacency(nbr.to(torch.int64), vis, counts, group=group)
