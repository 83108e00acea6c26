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
