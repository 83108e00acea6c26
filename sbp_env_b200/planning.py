"""Batched entry points the PRM / RRT* planners call on whole candidate sets
(BASELINE.json north_star, item 3) and the roadmap-connection primitive built
from them.

  visible_batch(cc, P, Q)   = [cc.visible(p, q) for p, q in zip(P, Q)]
  knn(tree, X, k)           = per row: stable argsort of np.linalg.norm(poses - x)[:k]
  near(tree, X, r)          = per row: prmPlanner.nearest_neighbours (prmPlanner.py:28-44)
  prm_connect_knn(...)      = the double loop of PRMPlanner.build_graph
                              (prmPlanner.py:91-101) with k-NN instead of the PRM*
                              radius, one kNN launch + one visible launch
"""
from __future__ import annotations

import numpy as np

from .runtime import _is_torch_tensor


def visible_batch(cc, P, Q):
    return cc.visible_batch(P, Q)


def feasible_batch(cc, P):
    return cc.feasible_batch(P)


def knn(tree, X, k, return_dist=True):
    return tree.knn(X, k, return_dist=return_dist)


def near(tree, X, r, metric="full", closed=False):
    return tree.near(X, r, metric=metric, closed=closed)


def prm_connect_knn(cc, tree, k: int, rows=None, return_dist: bool = False, queries=None,
                    mark=None, packed_out=None, part=None, want_rows: bool = True, order_out=None):
    """Candidate roadmap edges of every stored node to its k nearest other nodes and
    their visibility, entirely on the device.

    For row node v and neighbour m_g the check is ``cc.visible(m_g.pos, v.pos)``
    (argument order of prmPlanner.py:97).  A node's own entry (distance 0 to itself,
    ``m_g is v`` at prmPlanner.py:94-95) is dropped by asking for k+1 neighbours and
    removing the row index; when duplicates of the row node precede it in index order
    the last column is dropped instead, so rows always hold k candidates.

    :param rows: optional ``(first, count)`` to process a slice of the nodes (the
        multi-GPU shard of this rank); neighbours are still searched among all nodes.
    :param queries: optional CUDA tensor [m][d] of external samples to connect to the
        stored roadmap instead of the stored rows (no self entry to drop).
    :param packed_out: optional destination of the packed adjacency ``(nbr + 1) * 2 + vis``
        (int32 per candidate edge), written by the visibility kernel itself as the edges are
        decided: a CUDA int32 tensor [m][k], or ``(address, True)`` for the NVSwitch multicast
        address of a symmetric buffer (``distributed.PeerAdjacencyGather.target``).  2-D image
        checker only.
    :param part: ``(i, parts)``: connect only run i of the ``parts`` contiguous runs of the
        CELL-SORTED row order (spatially compact shards, one per rank; outputs are [n][k] in node
        order with only that run's rows written).  2-D image checker only.
    :param order_out: int32 CUDA tensor [n]: rows of all outputs in CELL-SORTED order, row p = node
        ``order_out[p]`` (written by the call); a part's rows are then one contiguous block stored
        with coalesced lines -- the layout for ``packed_out`` in pinned host memory or behind the
        multicast mapping (``Roadmap.build_from_packed(..., row_ids=order_out)`` consumes it).
    :param want_rows: ``False`` skips the ``nbr`` / ``vis`` outputs (only ``packed_out`` is written).
    :param mark: optional callable ``mark(name)`` invoked between the phases
        ("knn", "edges", "visible") on the launching stream (bench.py records CUDA
        events there).
    :return: ``nbr`` int64 [m][k] (CUDA), ``vis`` bool [m][k] (CUDA) and optionally
        ``dist`` fp64 [m][k] (only for ``queries``; stored rows include the self column).
    """
    tree.ctx.use_torch_stream()
    # the 2-D image checker resolves the kNN rows inside its visibility kernel (one launch)
    fused = (hasattr(cc, "knn_edges_visible") and hasattr(tree, "_h") and tree.dim == 2
             and not cc._mocked("visible"))
    # all stored rows (or a cell-sorted part of them) with the 2-D checker: the whole step is one
    # library call -- kNN on the cell grid, then the row-ordered visibility kernel
    if fused and queries is None and rows is None and not return_dist and 1 <= k <= 31 \
            and hasattr(cc, "prm_connect") and len(tree) > 0:
        out = cc.prm_connect(tree, k, part=part or (0, 1), packed_out=packed_out, want_rows=want_rows,
                             order_out=order_out)
        if mark:
            for name in ("knn", "edges", "visible"):
                mark(name)
        return out
    if part is not None and tuple(part) != (0, 1):
        raise ValueError("part= needs the 2-D image checker on a device node store (all rows)")
    if order_out is not None:
        raise ValueError("order_out= needs the 2-D image checker on a device node store (all rows)")
    if queries is not None:
        X = queries
        count = X.shape[0]
        idx, dist = tree.knn(X, k, return_dist=return_dist)
        first = 0
    else:
        n = len(tree)
        first, count = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
        X = tree.rows_device(first, count)
        idx, dist = tree.knn(X, k + 1, return_dist=False)
        dist = None
    if mark:
        mark("knn")
    if packed_out is not None and not fused:
        raise ValueError("packed_out needs the 2-D image checker on a device node store")
    if fused:
        if mark:
            mark("edges")                            # no separate edge-gather launch
        nbr, vis = cc.knn_edges_visible(tree, idx, first_row=first, queries=queries,
                                        packed_out=packed_out)
    else:
        nbr, P, Q, valid = tree.knn_edges(idx, first_row=first, queries=queries)
        if mark:
            mark("edges")
        vis = cc.visible_batch(P, Q).reshape(count, k)
        if _is_torch_tensor(valid):
            import torch

            valid = valid.view(torch.bool)           # 0 / 1 bytes written by k_knn_edges
        vis = vis & valid.reshape(count, k)
    if mark:
        mark("visible")
    if return_dist:
        return nbr, vis, dist
    return nbr, vis


def connect_to_roadmap(cc, tree, X, k: int = 8):
    """Roadmap node each external point connects to: the nearest of its ``k`` nearest stored
    nodes that is visible from it, -1 when none is -- the batched form of
    ``PRMPlanner.get_solution``'s ``get_nearest_free`` over the neighbourhood of the start / goal
    (prmPlanner.py:103-126, :131-138: closest visible node; there by ``engine.dist`` among the
    nodes inside ``epsilon``, here among the k nearest).  ``X``: CUDA tensor [q][d]; returns a
    CUDA int64 tensor [q]."""
    import torch

    nbr, vis = prm_connect_knn(cc, tree, k, queries=X)
    first = vis.to(torch.uint8).argmax(dim=1, keepdim=True)           # first visible column
    pick = nbr.gather(1, first).squeeze(1)
    return torch.where(vis.any(dim=1), pick, torch.full_like(pick, -1))


def roadmap_edges_to_host(nbr, vis, first_row: int = 0):
    """(src, dst) int64 arrays of the visible directed edges m_g -> v (prmPlanner.py:99-101)."""
    if _is_torch_tensor(nbr):
        nbr, vis = nbr.cpu().numpy(), vis.cpu().numpy()
    rows = np.repeat(np.arange(first_row, first_row + nbr.shape[0], dtype=np.int64), nbr.shape[1])
    sel = vis.reshape(-1)
    return nbr.reshape(-1)[sel], rows[sel]


def prm_connect_radius(cc, tree, radius: float):
    """The reference's own PRM* connection rule in two launches: for every stored node v,
    ``nearest_neighbours(nodes, poses, v.pos, radius)`` (all dims, strict ``<``,
    index-ascending, prmPlanner.py:28-44, :92) minus v itself (``m_g is v``, :94-95), and
    ``cc.visible(m_g.pos, v.pos)`` for each pair (:97).

    Returns CUDA tensors ``(src, dst, vis)``: candidate directed edges m_g -> v in the
    order the reference's double loop visits them, and their visibility.  (at
    the reference's gamma the radius exceeds every shipped map, i.e. this is the complete
    graph; use it with small n or a sane radius.)"""
    import torch

    tree.ctx.use_torch_stream()
    n = len(tree)
    X = tree.rows_device(0, n)
    row_ptr, idx = tree.near(X, radius, metric="full", closed=False)
    dev = X.device
    dst = torch.repeat_interleave(torch.arange(n, device=dev, dtype=torch.int64), row_ptr[1:] - row_ptr[:-1])
    keep = idx != dst                      # drop v itself; duplicates of v at other indices stay
    src, dst = idx[keep], dst[keep]
    vis = cc.visible_batch(X[src].contiguous(), X[dst].contiguous())
    return src, dst, vis


class CapturedStep:
    """A fixed sequence of launches of this library (and torch copies), captured once into a
    CUDA graph and replayed with a single ``cudaGraphLaunch``.

    ``fn()`` is called ``warmup`` times eagerly on a side stream (so that every work-space
    reaches its final size and nothing allocates or synchronises during capture), then once
    under capture; whatever it returns is kept in ``out`` and overwritten by each replay.
    Shapes, sizes and addresses are frozen; the *contents* of the input tensors, of the node
    stores and of the maps may change between replays.
    """

    def __init__(self, fn, device: int = 0, warmup: int = 2):
        import torch

        dev = torch.device("cuda", device)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from .runtime import get_context

        self._fn, self._dev, self._side = fn, dev, side
        self._ctx = get_context(device)
        self._capture()

    def _capture(self):
        import torch

        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=self._side):
            self.out = self._fn()
        # the library's internal work buffers are frozen into the graph by address: if a later
        # eager call makes one of them grow, the context's scratch generation moves and the
        # graph is captured again before its next replay (never replayed on freed memory)
        self._generation = self._ctx.scratch_generation

    def replay(self):
        if self._ctx.scratch_generation != self._generation:
            import torch

            torch.cuda.synchronize(self._dev)
            with torch.cuda.stream(self._side):
                self._fn()                       # eager: work buffers settle at their new size
            torch.cuda.synchronize(self._dev)
            self._capture()
            self._recaptured()
        self.graph.replay()
        return self.out

    def _recaptured(self):
        pass


class PRMConnectGraph(CapturedStep):
    """:func:`prm_connect_knn` for a fixed problem shape as a :class:`CapturedStep`.

    One connection step is a dozen short kernels (cell grid, cell-grid kNN, the brute-force
    kernels behind it, visibility); issued one by one from Python it takes twice the ~0.1 ms the
    GPU needs for a 50 000-node roadmap, so the inner loop is captured and a step becomes a single
    graph launch.  Results are identical to the eager call.  ``tree`` must not change size
    and ``queries`` (when given) must be updated in place.

    >>> g = PRMConnectGraph(cc, tree, k=10)
    >>> nbr, vis = g.replay()          # tensors owned by the graph, overwritten by each replay
    """

    def __init__(self, cc, tree, k: int, rows=None, queries=None, warmup: int = 2, packed_out=None,
                 after=None, part=None, want_rows: bool = True, order_out=None):
        """``after``: optional callable run (and captured) right after the connection step, on
        the same stream -- e.g. the barrier of ``PeerAdjacencyGather.complete``."""
        if "visible" in getattr(cc, "__dict__", {}):
            raise ValueError("a mocked checker cannot be captured into a CUDA graph")
        self.cc, self.tree, self.k = cc, tree, int(k)
        self._n = len(tree)
        before = [0]

        def fn():
            before[0] = int(cc.stats.visible_cnt)
            out = prm_connect_knn(cc, tree, self.k, rows=rows, queries=queries, packed_out=packed_out,
                                  part=part, want_rows=want_rows, order_out=order_out)
            if after is not None:
                after()
            return out

        super().__init__(fn, device=tree.ctx.device, warmup=warmup)
        self._edges = int(cc.stats.visible_cnt) - before[0]     # edges checked by one step
        self.nbr, self.vis = self.out if self.out is not None else (None, None)

    def _recaptured(self):
        self.nbr, self.vis = self.out if self.out is not None else (None, None)

    def replay(self):
        if len(self.tree) != self._n:
            raise RuntimeError("the node store changed size since the graph was captured")
        super().replay()
        self.cc.stats.visible_cnt += self._edges        # Stats accounting of visible_batch
        return self.nbr, self.vis
