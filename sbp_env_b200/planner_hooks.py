"""Instance-level hooks that put the batched GPU primitives under an UNMODIFIED reference
planner object.

``accelerate_rrt(planner)`` takes a ``RRTPlanner`` (RRT*, Informed-RRT*, and every planner that
inherits its ``run_once``; /root/reference/sbp_env/planners/rrtPlanner.py) after ``Env`` has
initialised it and replaces, on that INSTANCE only:

  add_newnode                   :106-113   also appends the pose to a device ``NodeStore``
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


def accelerate_rrt(planner, store=None):
    """Hook ``planner`` (see module docstring).  ``store`` may be passed in (tests); by default
    a ``NodeStore`` of the engine's dimension on the checker's device is created and filled
    with the nodes the planner already holds.  Returns the store."""
    if getattr(planner, "_sbp_accelerated", False):
        return planner._sbp_store
    args = planner.args
    cc = args.engine.cc
    if store is None:
        store = NodeStore(args.engine.get_dimension(), capacity=len(planner.poses),
                          device=getattr(cc, "_device", 0))
    n0 = len(planner.nodes)
    if n0:
        store.extend(planner.poses[:n0])
    index_of = {id(n): i for i, n in enumerate(planner.nodes)}
    orig_add = planner.add_newnode
    orig_nearest = planner.find_nearest_neighbour_idx
    orig_choose = planner.choose_least_cost_parent
    costs = _NodeCosts(planner.nodes)

    def add_newnode(node):
        index_of[id(node)] = len(planner.nodes)
        orig_add(node)
        store.append(node.pos)

    def find_nearest_neighbour_idx(pos, poses):
        # the planner passes its own prefix `self.poses[:len(self.nodes)]` (rrtPlanner.py:77)
        if getattr(poses, "base", None) is planner.poses and len(poses) == len(store):
            return store.nearest(pos)
        return orig_nearest(pos, poses)

    def choose_least_cost_parent(newnode, nn=None, nodes=None, skip_optimality=False, use_rtree=False,
                                 poses=None):
        if skip_optimality or use_rtree or poses is not None or nodes is not planner.nodes \
                or len(nodes) != len(store):
            return orig_choose(newnode, nn, nodes, skip_optimality=skip_optimality, use_rtree=use_rtree,
                               poses=poses)
        planner._new_node_dist_to_all_others = _LazyDistances(args.engine.dist)
        nn_idx = None if nn is None else index_of[id(nn)]
        parent, cost, _ = planner_ops.choose_least_cost_parent(
            cc, store, costs, newnode.pos, nn_idx, args.radius, dist=args.engine.dist, poses=planner.poses)
        nn = nodes[parent]
        newnode.cost = cost
        assert newnode is not nn
        newnode.parent = nn
        return newnode, nn

    planner.add_newnode = add_newnode
    planner.find_nearest_neighbour_idx = find_nearest_neighbour_idx
    planner.choose_least_cost_parent = choose_least_cost_parent
    planner._sbp_accelerated = True
    planner._sbp_store = store
    return store


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
        if len(store) != n:                          # (re)fill: PRM only appends between builds
            if len(store) > n:
                store.truncate(0)
            store.extend(poses[len(store):])
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
