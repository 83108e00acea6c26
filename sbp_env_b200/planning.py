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
                    mark=None):
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
    :param mark: optional callable ``mark(name)`` invoked between the phases
        ("knn", "edges", "visible") on the launching stream (bench.py records CUDA
        events there).
    :return: ``nbr`` int64 [m][k] (CUDA), ``vis`` bool [m][k] (CUDA) and optionally
        ``dist`` fp64 [m][k] (only for ``queries``; stored rows include the self column).
    """
    tree.ctx.use_torch_stream()
    if queries is not None:
        X = queries
        count = X.shape[0]
        idx, dist = tree.knn(X, k, return_dist=return_dist)
        if mark:
            mark("knn")
        nbr, P, Q, valid = tree.knn_edges(idx, queries=X)
    else:
        n = len(tree)
        first, count = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
        X = tree.rows_device(first, count)
        idx, dist = tree.knn(X, k + 1, return_dist=False)
        if mark:
            mark("knn")
        nbr, P, Q, valid = tree.knn_edges(idx, first_row=first)
        dist = None
    if mark:
        mark("edges")
    vis = cc.visible_batch(P, Q).reshape(count, k) & valid.bool().reshape(count, k)
    if mark:
        mark("visible")
    if return_dist:
        return nbr, vis, dist
    return nbr, vis


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
