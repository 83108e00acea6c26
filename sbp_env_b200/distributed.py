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
    neighbour block and its visibility mask travel as 4 bytes per edge -- in ONE collective,
    or in one device-to-host copy -- instead of 8 + 1.  CUDA tensors are packed by one launch
    of the library's pack kernel (``sbp_adjacency_peer_scatter`` with itself as the only peer)."""
    import torch

    if nbr.is_cuda and nbr.dtype == torch.int64 and vis.dtype in (torch.bool, torch.uint8):
        import ctypes as C

        from . import _lib
        from .runtime import get_context, ptr

        ctx = get_context(nbr.device.index)
        ctx.use_torch_stream()
        nbr = nbr.contiguous()
        vis8 = vis.contiguous().view(torch.uint8)
        out = torch.empty(nbr.shape, dtype=torch.int32, device=nbr.device)
        dst = (C.c_void_p * 1)(out.data_ptr())
        _lib.check(_lib.load().sbp_adjacency_peer_scatter(ctx.handle, ptr(nbr), ptr(vis8), nbr.numel(),
                                                          dst, 1, 0), "sbp_adjacency_peer_scatter")
        return out
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


class PeerAdjacencyGather:
    """All-gather of equally sized (nbr, vis) blocks WITHOUT a collective library call on the
    data path: one kernel packs this rank's block and stores it over NVLink into every
    rank's symmetric buffer (``sbp_adjacency_peer_scatter``), then the ranks meet at a
    signal-pad barrier of torch's symmetric memory.  Two buffers alternate, so a rank may
    still be reading step s while step s + 1 is being written.

    >>> ag = PeerAdjacencyGather(rows_per_rank=m, k=k, device=local_rank)
    >>> packed = ag.gather(nbr, vis)          # int32 [world * m][k], decode with unpack_adjacency
    """

    def __init__(self, rows_per_rank: int, k: int, device: int, group=None, multicast: bool = True):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from .runtime import get_context

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.m, self.k = int(rows_per_rank), int(k)
        self.block = self.m * self.k
        self.stride = (self.block + 3) // 4 * 4                    # 16-byte aligned row offsets
        self.ctx = get_context(device)
        dev = torch.device("cuda", device)
        # per half: the world blocks + 64 ints of rendezvous signals (one per rank)
        self._half_ints = self.world * self.stride + 64
        self.buf = symm_mem.empty((2, self._half_ints), dtype=torch.int32, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)                                   # every rank's signals are zero
        self._state = torch.zeros(4, dtype=torch.int32, device=dev)   # CTA ticket, epoch of half 0 / 1
        base = [int(p) for p in self.hdl.buffer_ptrs]
        import ctypes as C

        half = self._half_ints * 4
        self._ptrs = [(C.c_void_p * self.world)(*[b + h * half for b in base]) for h in (0, 1)]
        # NVSwitch multicast (NVLS) mapping of the same buffer, when the fabric offers it: the
        # block is then sent once and replicated by the switch
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if multicast else 0
        self._mc = [C.c_void_p(mc + h * half) for h in (0, 1)] if mc else None
        self.kind = "NVSwitch multicast stores" if mc else "NVLink peer stores"
        self._step = 0

    # ---- fused form: the producing kernel stores straight into the multicast mapping -----
    @property
    def fused(self) -> bool:
        """True when the fabric offers the multicast mapping the fused form needs."""
        return self._mc is not None

    def target(self, h: int):
        """``packed_out`` for :func:`planning.prm_connect_knn` / ``PRMConnectGraph``: the multicast
        address of this rank's block in buffer ``h`` (0 / 1).  The visibility kernel then sends
        every adjacency code through the switch as the edge is decided, so the exchange overlaps
        the computation and :meth:`complete` is only the barrier."""
        if self._mc is None:
            raise RuntimeError("no multicast mapping: use gather()")
        return (int(self._mc[h].value) + self.rank * self.stride * 4, True)

    def target_with_rendezvous(self, h: int):
        """Like :meth:`target`, plus the block that makes the kernel itself meet the other ranks
        (its last CTA publishes an epoch through the switch and waits for everyone's): after the
        step the adjacency is complete on this rank, :meth:`view` needs no barrier."""
        addr, mc = self.target(h)
        sig_off = self.world * self.stride * 4
        sync = (int(self._mc[h].value) + sig_off,                      # signals, multicast address
                int(self.buf[h].data_ptr()) + sig_off,                 # signals, local address
                int(self._state.data_ptr()),                           # CTA ticket
                int(self._state.data_ptr()) + 4 * (1 + h),             # epoch of this half
                self.rank, self.world)
        return (addr, mc, sync)

    def complete(self, h: int):
        """Barrier after a step that wrote through :meth:`target` (h); returns the assembled
        int32 [world * m][k] block like :meth:`gather`."""
        self.ctx.use_torch_stream()
        self.hdl.barrier(channel=h)
        return self.view(h)

    def view(self, h: int):
        """The assembled int32 [world * m][k] block of buffer half ``h`` (no synchronisation)."""
        out = self.buf[h][: self.world * self.stride].view(self.world, self.stride)[:, : self.block]
        return out.reshape(self.world, self.m, self.k).reshape(self.world * self.m, self.k)

    def gather(self, nbr, vis):
        import torch

        from . import _lib
        from .runtime import ptr

        if nbr.shape != (self.m, self.k) or vis.shape != (self.m, self.k):
            raise ValueError("block shape differs from the one the buffers were sized for")
        self.ctx.use_torch_stream()
        h = self._step & 1
        self._step += 1
        nbr = nbr.contiguous()
        vis8 = vis.contiguous().view(torch.uint8) if vis.dtype == torch.bool else vis.contiguous()
        if self._mc is not None:
            _lib.check(_lib.load().sbp_adjacency_multicast(self.ctx.handle, ptr(nbr), ptr(vis8), self.block,
                                                           self._mc[h], self.rank * self.stride),
                       "sbp_adjacency_multicast")
        else:
            _lib.check(_lib.load().sbp_adjacency_peer_scatter(self.ctx.handle, ptr(nbr), ptr(vis8), self.block,
                                                              self._ptrs[h], self.world, self.rank * self.stride),
                       "sbp_adjacency_peer_scatter")
        self.hdl.barrier(channel=h)
        return self.view(h)


class PeerAdjacencyBuffer:
    """The roadmap's packed adjacency ``int32 [n][k]`` replicated on every rank through symmetric
    memory: every rank's connection kernel stores the rows it owns through the NVSwitch multicast
    mapping (``target(h)`` -> ``packed_out`` of ``prm_connect_knn(part=...)``), so every rank's
    copy is complete after the barrier (``complete(h)``); two halves alternate so that step s may
    still be read while step s + 1 is written.  With ``order_out=`` the rows are in cell-sorted
    order and a rank's rows are one contiguous block stored with coalesced lines (the efficient
    form over NVLink: 4-byte stores scattered in node order took 3x the time of the walk itself).

    >>> buf = PeerAdjacencyBuffer(n, k, device=local_rank)
    >>> prm_connect_knn(cc, tree, k, part=(rank, world), packed_out=buf.target(h), want_rows=False,
    ...                 order_out=order)
    >>> packed = buf.complete(h)                 # int32 [n][k] on every rank, row p = node order[p]
    """

    def __init__(self, n_rows: int, k: int, device: int, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from .runtime import get_context

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 64:
            raise ValueError("PeerAdjacencyBuffer: at most 64 ranks")
        self.n, self.k = int(n_rows), int(k)
        self.ctx = get_context(device)
        dev = torch.device("cuda", device)
        # per half: the packed rows [n][k], then the row permutation [n] (every rank stores its own
        # slice of both through the same mapping)
        self._rows = (self.n * self.k + 3) // 4 * 4
        self._half = self._rows + (self.n + 3) // 4 * 4
        self.buf = symm_mem.empty((2, self._half), dtype=torch.int32, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        self._mc = [mc + h * self._half * 4 for h in (0, 1)] if mc else None

    @property
    def fused(self) -> bool:
        """True when the fabric offers the multicast mapping."""
        return self._mc is not None

    def target(self, h: int):
        if self._mc is None:
            raise RuntimeError("no multicast mapping on this fabric")
        return (self._mc[h], True)

    def order_target(self, h: int) -> int:
        """Multicast address of half h's row permutation (``order_out`` of ``prm_connect_knn``)."""
        if self._mc is None:
            raise RuntimeError("no multicast mapping on this fabric")
        return self._mc[h] + self._rows * 4

    def local(self, h: int):
        """This rank's own copy of half h as ``packed_out`` (no exchange: for fabrics without the
        multicast mapping the blocks are then all-gathered by the caller)."""
        return self.view(h)

    def complete(self, h: int):
        self.ctx.use_torch_stream()
        self.hdl.barrier(channel=h)
        return self.view(h)

    def view(self, h: int):
        return self.buf[h][: self.n * self.k].view(self.n, self.k)

    def order_view(self, h: int):
        return self.buf[h][self._rows: self._rows + self.n]


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
