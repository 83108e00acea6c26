"""The planners' per-node loops, restated on top of the GPU primitives.

Each function reproduces the *result* of one reference loop (file:line under
/root/reference/sbp_env/planners/) while replacing its O(N) Python iteration over the
tree by one fused kernel launch (`NodeStore.near_visible`) plus O(#neighbours) host
arithmetic.  Costs and tie-breaking follow the reference exactly: distances that enter
a cost are evaluated on the host with the reference's own formula
(`engine.euclidean_dist`, engine.py:13-24, first two coordinates, scalar ``** 2``), and
candidates are visited in node-index order with the same strict / non-strict
comparisons.

SURVEY.md D2: in the shipped reference `RRTPlanner.rewire` is a no-op (it queries an
R-tree nothing ever inserts into); :func:`rewire` implements the *linear* variant
(rrtPlanner.py:250-266) and is therefore opt-in for the planner that adopts it.
"""
from __future__ import annotations

import numpy as np

from .engine import euclidean_dist


def choose_least_cost_parent(cc, tree, costs, newpos, nn_idx, radius, dist=euclidean_dist,
                             poses=None):
    """rrtPlanner.py:173-196 (linear path).

    :param tree: :class:`NodeStore` holding the tree's poses (device)
    :param costs: host sequence, ``costs[i]`` = cost of node i
    :param nn_idx: index of the current nearest node, or ``None``
    :param poses: optional host copy of the poses (``planner.poses``); fetched from the
        device for the few candidates when omitted
    :return: ``(parent_idx, new_cost, (idx, vis, c))`` -- the candidate set is returned so
        that :func:`rewire` can reuse it, like ``_new_node_dist_to_all_others``.
    """
    newpos = np.asarray(newpos, dtype=np.float64)
    idx, vis = tree.near_visible(cc, newpos, radius, metric="xy", closed=True, direction=0)
    P = _rows(tree, poses, idx)
    c = np.array([dist(newpos, p) for p in P], dtype=np.float64)
    nn = nn_idx
    c_nn = dist(newpos, _rows(tree, poses, [nn])[0]) if nn is not None else None
    for j, p in enumerate(idx):
        # the device gate uses the correctly rounded product, engine.dist libm pow: re-apply
        # the reference's own `<=` on its own value (differs only within 1 ulp of radius)
        if c[j] <= radius and vis[j]:
            if nn is None or costs[p] + c[j] < costs[nn] + c_nn:
                nn, c_nn = int(p), c[j]
    if nn is None:
        raise LookupError("ERROR: Provided nn=None, and cannot find any valid nn by this function. "
                          "This newnode is not close to the root tree...?")
    new_cost = costs[nn] + dist(_rows(tree, poses, [nn])[0], newpos)
    return nn, new_cost, (idx, vis, c)


def rewire(cc, tree, costs, parents, new_idx, radius, dist=euclidean_dist, poses=None,
           candidates=None):
    """rrtPlanner.py:250-266 (linear variant): re-parent every node that gets cheaper
    through the new node.  Mutates ``costs`` / ``parents``; returns the rewired indices.

    In the 2-D checker ``visible(n, new)`` equals ``visible(new, n)`` (direction
    symmetric), so the candidate set of :func:`choose_least_cost_parent` can be passed in;
    for the arm checker the edge is re-checked in the rewire direction."""
    newpos = _rows(tree, poses, [new_idx])[0]
    arm = hasattr(cc, "stick_robot_length_config")
    if candidates is None or arm:
        idx, vis = tree.near_visible(cc, newpos, radius, metric="xy", closed=True, direction=1)
        c = np.array([dist(newpos, p) for p in _rows(tree, poses, idx)], dtype=np.float64)
    else:
        idx, vis, c = candidates
    # The reference compares Node objects, and Node.__eq__ / __hash__ go by POSITION
    # (utils/common.py:187-192): `x not in already_rewired` (= {newnode}) skips every node at the
    # new node's position, `n != newnode.parent` every node at the parent's position (goal-biased
    # sampling does produce such duplicates).
    P = _rows(tree, poses, idx) if len(idx) else np.empty((0, len(newpos)))
    ppos = _rows(tree, poses, [parents[new_idx]])[0] if parents[new_idx] is not None and parents[new_idx] >= 0 else None
    skip = [int(n) == new_idx or bool(np.all(P[j] == newpos)) or
            (ppos is not None and bool(np.all(P[j] == ppos))) for j, n in enumerate(idx)]
    if candidates is not None and not arm:
        # Stats parity: the reference's loop evaluates cc.visible(n.pos, new.pos) for every node
        # that passed `n != newnode.parent and dist <= radius` (rrtPlanner.py:256-260); the reused
        # candidate set already holds those answers, so only the counter moves
        cc.stats.visible_cnt += sum(1 for j in range(len(idx)) if not skip[j] and c[j] <= radius)
    out = []
    for j, n in enumerate(idx):
        n = int(n)
        if skip[j]:
            continue
        if c[j] <= radius and vis[j] and costs[new_idx] + c[j] < costs[n]:
            parents[n] = new_idx
            costs[n] = costs[new_idx] + c[j]
            out.append(n)
    return out


def birrt_join(cc, other_tree, newpos, epsilon):
    """birrtPlanner.py:82-87: nearest node of the other tree (all dims); joinable when it
    is closer than epsilon and visible(other_pose, newpos).  Returns the index or None."""
    newpos = np.asarray(newpos, dtype=np.float64).reshape(1, other_tree.dim)
    idx, d = other_tree.knn(newpos, 1)
    i = int(idx[0, 0])
    if i < 0 or not d[0, 0] < epsilon:
        return None
    other = other_tree.poses_rows([i])[0]
    return i if cc.visible(other, newpos[0]) else None


def prm_nearest_free(cc, tree, pos, radius, exclude=(), dist=euclidean_dist, poses=None):
    """prmPlanner.py:131-138 + :103-126: among the nodes with ``norm(pose - pos) < radius``
    (all dims, strict) the closest one -- by ``engine.dist``, first minimum -- that is
    visible from ``pos``.  Returns the index or None."""
    pos = np.asarray(pos, dtype=np.float64)
    idx, vis = tree.near_visible(cc, pos, radius, metric="full", closed=False, direction=0)
    nn, best = None, 999999
    P = _rows(tree, poses, idx)
    for j, n in enumerate(idx):
        if int(n) in exclude or not vis[j]:
            continue
        cst = dist(pos, P[j])
        if nn is None or cst < best:
            nn, best = int(n), cst
    return nn


def nearest_in_trees(trees, pos, radius, own=None, dist=euclidean_dist):
    """rrdtPlanner.py:685-721: one arg-min scan per disjoint tree (skipping `own`), kept
    when ``engine.dist < radius``.  Returns ``[(tree_index, node_index), ...]``."""
    pos = np.asarray(pos, dtype=np.float64)
    out = []
    for ti, t in enumerate(trees):
        if ti == own or len(t) == 0:
            continue
        i = t.nearest(pos)
        if dist(pos, t.poses_rows([i])[0]) < radius:
            out.append((ti, i))
    return out


def _rows(tree, poses, idx):
    idx = np.asarray(idx, dtype=np.int64)
    if poses is not None:
        return np.asarray(poses)[idx]
    return tree.poses_rows(idx)
