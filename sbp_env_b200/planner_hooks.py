"""Instance-level hooks that put the batched GPU primitives under an UNMODIFIED reference
planner object.

``accelerate_rrt(planner)`` takes a ``RRTPlanner`` (RRT*, Informed-RRT*, and every planner that
inherits its ``run_once``; /root/reference/sbp_env/planners/rrtPlanner.py) after ``Env`` has
initialised it and replaces, on that INSTANCE only:

  (the node arrays)             :106-113   mirrored lazily into a device ``NodeStore``
  find_nearest_neighbour_idx    :268-279   ``NodeStore.nearest`` (one wide-kNN launch) when the
                                           array passed in is the planner's own node prefix
  choose_least_cost_parent      :173-196   linear path: the O(N) Python loop over all nodes
                                           (91 % of the reference's RRT* time)
                                           becomes ONE fused near + visible launch plus host
                                           arithmetic on the few nodes inside the radius
                                           (``planner_ops.choose_least_cost_parent``)

Everything else -- sampling, ``step_from_to``, goal test, ``rewire`` (a no-op in the shipped
reference), Stats -- runs as before, and the seeded run is identical: same
nodes, costs, parents, ``c_max`` and counters (tests/test_reference_dropin.py).  The other
call forms of ``choose_least_cost_parent`` (``skip_optimality``, ``use_rtree``, ``poses=``) are
passed through to the original method.
"""
from __future__ import annotations

import numpy as np

from . import planner_ops
from .node_store import NodeStore


class _NodeCosts:
    """``costs[i]`` of planner_ops read straight from the planner's Node objects."""

    def __init__(self, nodes):
        self.nodes = nodes

    def __getitem__(self, i):
        return self.nodes[i].cost


class _LazyDistances(dict):
    """``planner._new_node_dist_to_all_others`` (rrtPlanner.py:174-176): the reference fills it
    for every node; only entries that are looked up are computed here."""

    def __init__(self, dist):
        super().__init__()
        self._dist = dist

    def __missing__(self, key):
        v = self._dist(key[0].pos, key[1].pos)
        self[key] = v
        return v


class _Tree:
    """One tree of a planner (its Node list and pose array) and the device store that mirrors
    it.  The planners append to their arrays directly (``poses[len(nodes)] = pos;
    nodes.append(node)``: rrtPlanner.py:106-113, birrtPlanner.py:67-69, :113-114), so the mirror
    is brought up to date lazily, at the start of every hooked query."""

    def __init__(self, nodes, poses, store):
        self.nodes, self.poses, self.store = nodes, poses, store
        self.index_of = {}
        self.costs = _NodeCosts(nodes)

    def sync(self):
        have, n = len(self.store), len(self.nodes)
        if have < n:
            self.store.extend(self.poses[have:n])
            for i in range(have, n):
                self.index_of[id(self.nodes[i])] = i
        return self.store


def _accelerate(planner, trees):
    args = planner.args
    cc = args.engine.cc
    orig_nearest = planner.find_nearest_neighbour_idx
    orig_choose = planner.choose_least_cost_parent

    def tree_of_poses(poses):
        base = getattr(poses, "base", None)
        for t in trees:
            if base is t.poses and len(poses) == len(t.nodes):
                return t
        return None

    def find_nearest_neighbour_idx(pos, poses):
        # the planners pass a prefix of their own pose array (rrtPlanner.py:77, birrtPlanner.py:54)
        t = tree_of_poses(poses)
        if t is None:
            return orig_nearest(pos, poses)
        return t.sync().nearest(pos)

    def choose_least_cost_parent(newnode, nn=None, nodes=None, skip_optimality=False, use_rtree=False,
                                 poses=None):
        t = next((t for t in trees if nodes is t.nodes), None)
        if skip_optimality or use_rtree or poses is not None or t is None:
            return orig_choose(newnode, nn, nodes, skip_optimality=skip_optimality, use_rtree=use_rtree,
                               poses=poses)
        store = t.sync()
        planner._new_node_dist_to_all_others = _LazyDistances(args.engine.dist)
        nn_idx = None if nn is None else t.index_of[id(nn)]
        parent, cost, _ = planner_ops.choose_least_cost_parent(
            cc, store, t.costs, newnode.pos, nn_idx, args.radius, dist=args.engine.dist, poses=t.poses)
        nn = nodes[parent]
        newnode.cost = cost
        assert newnode is not nn
        newnode.parent = nn
        return newnode, nn

    planner.find_nearest_neighbour_idx = find_nearest_neighbour_idx
    planner.choose_least_cost_parent = choose_least_cost_parent
    planner._sbp_accelerated = True
    planner._sbp_trees = trees


def _new_store(planner, poses):
    cc = planner.args.engine.cc
    return NodeStore(planner.args.engine.get_dimension(), capacity=len(poses), device=getattr(cc, "_device", 0))


def accelerate_rrt(planner, store=None):
    """Hook ``planner`` (see module docstring).  ``store`` may be passed in (tests); by default
    a ``NodeStore`` of the engine's dimension on the checker's device is created; it mirrors
    ``planner.poses[:len(planner.nodes)]``.  Returns the store."""
    if getattr(planner, "_sbp_accelerated", False):
        return planner._sbp_trees[0].store
    tree = _Tree(planner.nodes, planner.poses, store if store is not None else _new_store(planner, planner.poses))
    tree.sync()
    _accelerate(planner, [tree])
    return tree.store


def accelerate_birrt(planner, stores=None):
    """The same hooks for ``BiRRTPlanner`` (planners/birrtPlanner.py:36-124), whose ``run_once``
    alternates between the start tree (``nodes`` / ``poses``) and the goal tree
    (``goal_tree_nodes`` / ``goal_tree_poses``): each tree gets its own device store, picked by
    the identity of the array / list the planner passes in.  The join test (:80-87) and the
    re-parenting walk after the join (:101-118, which calls ``choose_least_cost_parent`` on the
    start tree) run unchanged on top.  Returns the two stores."""
    if getattr(planner, "_sbp_accelerated", False):
        return tuple(t.store for t in planner._sbp_trees)
    s0, s1 = stores if stores is not None else (_new_store(planner, planner.poses),
                                                _new_store(planner, planner.goal_tree_poses))
    trees = [_Tree(planner.nodes, planner.poses, s0),
             _Tree(planner.goal_tree_nodes, planner.goal_tree_poses, s1)]
    for t in trees:
        t.sync()
    _accelerate(planner, trees)
    return s0, s1


class _ArrayMirror:
    """Device mirror of one append-only pose array, keyed by the identity of the array (RRdT*
    owns one array per tree: rrdtPlanner.py:858-887).  The planner's prefix is brought up to
    date lazily; the last mirrored row is compared first, so that an array that was rewritten
    rather than appended to is mirrored again from scratch."""

    def __init__(self, base, store):
        self.base, self.store, self.last = base, store, None

    def sync(self, n):
        have = len(self.store)
        if have > n or (have and not np.array_equal(self.base[have - 1], self.last)):
            self.store.truncate(0)
            have = 0
        if have < n:
            self.store.extend(self.base[have:n])
        if n:
            self.last = np.array(self.base[n - 1], copy=True)
        return self.store


def accelerate_rrdt(planner, min_nodes: int = 1024, store_factory=None):
    """Hook for ``RRdTPlanner`` (planners/rrdtPlanner.py): its many trees each own a pose array
    and are searched one by one with ``find_nearest_neighbour_idx(pos, tree.poses[:len(tree.nodes)])``
    (:571, :709-711).  Arrays whose prefix has at least ``min_nodes`` rows -- in practice the
    root tree, which absorbs the others -- get a device mirror (created on first sight, keyed by
    array identity, synced lazily); small trees stay on numpy, where a launch would cost more
    than the scan.  Everything else runs unchanged; results are those of ``np.argmin`` over
    ``np.linalg.norm`` (first minimum).  Returns the dict of mirrors."""
    if getattr(planner, "_sbp_accelerated", False):
        return planner._sbp_mirrors
    cc = planner.args.engine.cc
    d = planner.args.engine.get_dimension()
    orig_nearest = planner.find_nearest_neighbour_idx
    mirrors = {}
    if store_factory is None:
        store_factory = lambda base: NodeStore(d, capacity=len(base), device=getattr(cc, "_device", 0))

    def find_nearest_neighbour_idx(pos, poses):
        base = getattr(poses, "base", None)
        n = len(poses)
        if base is None or n < min_nodes or base.ndim != 2 or base.shape[1] != d:
            return orig_nearest(pos, poses)
        m = mirrors.get(id(base))
        if m is None or m.base is not base:
            m = mirrors[id(base)] = _ArrayMirror(base, store_factory(base))
        return m.sync(n).nearest(pos)

    planner.find_nearest_neighbour_idx = find_nearest_neighbour_idx
    planner._sbp_accelerated = True
    planner._sbp_mirrors = mirrors
    return mirrors


def accelerate_prm(planner, store=None, chunk_rows: int = 4096):
    """Hook a reference ``PRMPlanner`` object (planners/prmPlanner.py): ``build_graph``
    (:80-101) -- for every node a Python scan of ALL nodes (``nearest_neighbours``, :28-44) and
    one ``cc.visible`` call per neighbour -- becomes, per block of ``chunk_rows`` nodes, one
    batched radius query (``NodeStore.near``: all dims, strict ``<``, index-ascending, exactly
    the order of the reference's double loop) and one ``visible_batch`` launch over the
    candidate pairs ``visible(m_g.pos, v.pos)``.  The edges are then inserted into the planner's
    own networkx ``DiGraph`` in the reference's order with ``engine.dist`` weights, so that
    ``get_solution`` (and anything else that reads ``planner.graph``) sees the identical graph.

    The radius is the reference's PRM* radius ``gamma * (log(n+1)/(n+1))^(1/d)`` (:87-89), which
    for the shipped maps exceeds the map: the candidate set is then all n^2 pairs,
    which is why the rows are processed in blocks.  Returns the node store."""
    args = planner.args
    cc = args.engine.cc
    d = args.engine.get_dimension()
    if store is None:
        store = NodeStore(d, capacity=len(planner.poses), device=getattr(cc, "_device", 0))

    def build_graph():
        n = len(planner.nodes)
        radius = planner.gamma * np.power(np.log(n + 1) / (n + 1), 1 / d)       # :87-89
        poses = np.ascontiguousarray(planner.poses[:n], dtype=np.float64)
        have = len(store)
        # the mirror is reused only when it is a bit-identical prefix of the planner's array
        # (PRM appends between builds; anything else -- a rewritten array -- is mirrored afresh)
        if have > n or (have and not np.array_equal(store.poses, poses[:have])):
            store.truncate(0)
            have = 0
        if have < n:
            store.extend(poses[have:])
        nodes, dist, graph = planner.nodes, args.engine.dist, planner.graph
        for first in range(0, n, chunk_rows):
            rows = poses[first:first + chunk_rows]
            row_ptr, idx = store.near(rows, float(radius), metric="full", closed=False)
            dst = np.repeat(np.arange(first, first + len(rows), dtype=np.int64), np.diff(row_ptr))
            keep = idx != dst                        # `if m_g is v: continue` (:94-95)
            src, dst = idx[keep], dst[keep]
            vis = np.asarray(cc.visible_batch(poses[src], poses[dst]), dtype=bool)   # visible(m_g, v), :97
            for a, b in zip(src[vis].tolist(), dst[vis].tolist()):
                graph.add_weighted_edges_from([(nodes[a], nodes[b], dist(nodes[a].pos, nodes[b].pos))])

    planner.build_graph = build_graph
    planner._sbp_store = store
    return store
