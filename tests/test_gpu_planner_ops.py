"""-m gpu: the planners' per-node loops (planner_ops) against line-by-line Python
restatements of the reference loops, with the oracle as the collision checker."""
import numpy as np
import pytest

import c_oracle as orc
from _util import load_map, map_png

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


def ref_dist(p1, p2):
    # engine.euclidean_dist, engine.py:13-24
    import math

    p = p1 - p2
    return math.sqrt(p[0] ** 2 + p[1] ** 2)


def ref_choose_parent(visible, poses, costs, newpos, nn, radius):
    # rrtPlanner.py:173-196, restated
    if nn is not None:
        c_nn = ref_dist(newpos, poses[nn])
    for p in range(len(poses)):
        c = ref_dist(newpos, poses[p])
        if c <= radius and visible(newpos, poses[p]):
            if nn is None or costs[p] + c < costs[nn] + c_nn:
                nn, c_nn = p, c
    if nn is None:
        raise LookupError
    return nn, costs[nn] + ref_dist(poses[nn], newpos)


def ref_rewire(visible, poses, costs, parents, new, radius):
    # rrtPlanner.py:250-266, restated (linear variant)
    out = []
    for n in range(len(poses)):
        if n == new:
            continue
        c = ref_dist(poses[new], poses[n])
        if n != parents[new] and c <= radius and visible(poses[n], poses[new]) and costs[new] + c < costs[n]:
            parents[n] = new
            costs[n] = costs[new] + c
            out.append(n)
    return out


def _grow_tree(sg, cc, om_visible, om_feasible, W, H, d, n_nodes, radius, eps, rng, arm_hi=None):
    """A seeded RRT*-style growth driven twice: GPU planner_ops vs the restated loops."""
    from sbp_env_b200 import planner_ops as ops

    tree = sg.NodeStore(d)
    poses, costs, parents = [], [], []
    rposes, rcosts, rparents = [], [], []
    start = None
    while start is None:
        c = rng.random(d) * ([W, H] + [2 * np.pi] * (d - 2)) - ([0, 0] + [np.pi] * (d - 2))
        if om_feasible(c):
            start = c
    tree.append(start); poses.append(start); costs.append(0.0); parents.append(-1)
    rposes.append(start); rcosts.append(0.0); rparents.append(-1)
    rewired_total = 0
    tries = 0
    while len(poses) < n_nodes and tries < 40 * n_nodes:
        tries += 1
        r = rng.random(d) * ([W, H] + [2 * np.pi] * (d - 2)) - ([0, 0] + [np.pi] * (d - 2))
        nn = tree.nearest(r)
        assert nn == int(np.argmin(np.linalg.norm(np.array(rposes) - r, axis=1)))
        v = r - poses[nn]
        step = min(ref_dist(poses[nn], r), eps)
        newpos = poses[nn] + step * v / np.linalg.norm(v)
        if not om_visible(poses[nn], newpos):
            continue
        parent, ncost, cand = ops.choose_least_cost_parent(cc, tree, costs, newpos, nn, radius,
                                                           poses=np.array(poses))
        eparent, ecost = ref_choose_parent(om_visible, np.array(rposes), rcosts, newpos, nn, radius)
        assert parent == eparent and ncost == ecost
        new = tree.append(newpos)
        poses.append(newpos); costs.append(ncost); parents.append(parent)
        rposes.append(newpos); rcosts.append(ecost); rparents.append(eparent)
        got = ops.rewire(cc, tree, costs, parents, new, radius, poses=np.array(poses), candidates=cand)
        exp = ref_rewire(om_visible, np.array(rposes), rcosts, rparents, new, radius)
        assert got == exp
        assert costs == rcosts and parents == rparents
        rewired_total += len(got)
    assert len(poses) == n_nodes
    return rewired_total


def test_rrt_star_growth_2d(sg):
    free = load_map("room1")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("room1"))
    om = orc.Map(free)
    vis = lambda a, b: bool(om.visible2d(np.array([a[:2]]), np.array([b[:2]]))[0])
    feas = lambda p: bool(om.feasible2d(np.array([p[:2]]))[0])
    n_rewired = _grow_tree(sg, cc, vis, feas, W, H, 2, 600, 11.0, 10.0, np.random.default_rng(0))
    assert n_rewired > 0            # the linear rewire variant does fire


def test_rrt_star_growth_arm_4d(sg):
    free = load_map("4d")
    W, H = free.shape
    L = [30.0, 30.0]
    cc = sg.RobotArm4dCollisionChecker(map_png("4d"), stick_robot_length_config=L)
    om = orc.Map(free)
    vis = lambda a, b: bool(om.arm_visible(np.array([a]), np.array([b]), *L)[0])
    feas = lambda p: bool(om.arm_feasible(np.array([p]), *L)[0])
    _grow_tree(sg, cc, vis, feas, W, H, 4, 250, 11.0, 10.0, np.random.default_rng(1))


def test_choose_parent_no_nn_raises(sg):
    from sbp_env_b200 import planner_ops as ops

    cc = sg.ImgCollisionChecker(map_png("room1"))
    t = sg.NodeStore(2)
    t.append([100.0, 100.0])
    with pytest.raises(LookupError):
        ops.choose_least_cost_parent(cc, t, [0.0], np.array([300.0, 300.0]), None, 11.0)


def test_near_visible_capacity_growth_and_stats(sg):
    free = load_map("blank")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("blank"))
    om = orc.Map(free)
    rng = np.random.default_rng(3)
    pts = rng.random((20000, 2)) * [W, H]
    t = sg.NodeStore(2)
    t.extend(pts)
    x = np.array([W / 2, H / 2])
    before = cc.stats.visible_cnt
    idx, vis = t.near_visible(cc, x, 60.0, metric="full", closed=False, cap=16)   # forces a retry
    rp, eidx = orc.near(pts, x[None], 60.0, "full", False)
    assert np.array_equal(idx, eidx) and len(idx) > 16
    assert np.array_equal(vis, om.visible2d(np.repeat(x[None], len(idx), 0), pts[idx]))
    assert cc.stats.visible_cnt - before == len(idx)


def test_birrt_join_and_prm_nearest_free(sg):
    from sbp_env_b200 import planner_ops as ops

    free = load_map("maze1")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("maze1"))
    om = orc.Map(free)
    rng = np.random.default_rng(5)
    pts = rng.random((4000, 2)) * [W, H]
    pts = pts[om.feasible2d(pts)]
    t = sg.NodeStore(2)
    t.extend(pts)
    for _ in range(200):
        x = rng.random(2) * [W, H]
        # birrtPlanner.py:82-87
        d = np.linalg.norm(pts - x, axis=1)
        exp = None
        if d.min() < 10.0:
            i = int(np.argmin(d))
            if om.visible2d(pts[i][None], x[None])[0]:
                exp = i
        assert ops.birrt_join(cc, t, x, 10.0) == exp
        # prmPlanner.py:103-126 over nearest_neighbours(radius)
        near = [i for i, v in enumerate(d) if v < 25.0]
        nn, best = None, 999999
        for i in near:
            if i in (0, 1) or not om.visible2d(x[None], pts[i][None])[0]:
                continue
            c = ref_dist(x, pts[i])
            if nn is None or c < best:
                nn, best = i, c
        assert ops.prm_nearest_free(cc, t, x, 25.0, exclude=(0, 1)) == nn
    trees = [sg.NodeStore(2) for _ in range(3)]
    for k, tr in enumerate(trees):
        tr.extend(pts[k::3])
    x = pts[7] + 0.25
    got = ops.nearest_in_trees(trees, x, 6.0, own=1)
    exp = []
    for k in (0, 2):
        sub = pts[k::3]
        i = int(np.argmin(np.linalg.norm(sub - x, axis=1)))
        if ref_dist(x, sub[i]) < 6.0:
            exp.append((k, i))
    assert got == exp


def test_accelerate_rrt_hooks_on_a_planner_look_alike(sg):
    """planner_hooks.accelerate_rrt with the REAL device node store and checker, on a planner
    object that restates the reference's RRTPlanner members the hook touches (poses, nodes,
    add_newnode :106-113, find_nearest_neighbour_idx :268-279, the linear
    choose_least_cost_parent :173-196).  The hooked and the plain object grow identical trees
    from the same seeded samples.  (The unmodified reference classes are hooked in
    tests/test_reference_dropin.py, where the device is replaced by the oracle.)"""
    import types

    free = load_map("room1")
    W, H = free.shape

    class Node:
        def __init__(self, pos):
            self.pos, self.cost, self.parent = np.array(pos, dtype=np.float64), 0.0, None

    class Planner:
        def __init__(self, cc):
            eng = types.SimpleNamespace(cc=cc, dist=ref_dist, get_dimension=lambda: 2)
            self.args = types.SimpleNamespace(engine=eng, radius=11.0)
            self.poses, self.nodes = np.empty((3000, 2)), []

        def add_newnode(self, node):
            self.poses[len(self.nodes)] = node.pos
            self.nodes.append(node)

        def find_nearest_neighbour_idx(self, pos, poses):
            return int(np.argmin(np.linalg.norm(poses - pos, axis=1)))

        def choose_least_cost_parent(self, newnode, nn=None, nodes=None, skip_optimality=False,
                                     use_rtree=False, poses=None):
            cc = self.args.engine.cc
            idx, cost = ref_choose_parent(lambda a, b: cc.visible(a, b), self.poses[: len(nodes)],
                                          [n.cost for n in nodes], newnode.pos,
                                          None if nn is None else nodes.index(nn), self.args.radius)
            newnode.cost, newnode.parent = cost, nodes[idx]
            return newnode, nodes[idx]

    def grow(hooked):
        cc = sg.ImgCollisionChecker(map_png("room1"))
        vis0 = cc.stats.visible_cnt                 # process-global Stats: count the delta
        pl = Planner(cc)
        pl.add_newnode(Node((100.0, 100.0)))
        store = sg.planner_hooks.accelerate_rrt(pl) if hooked else None
        rng = np.random.default_rng(5)
        while len(pl.nodes) < 700:
            r = rng.random(2) * [W, H]
            if not cc.feasible(r):
                continue
            nn = pl.nodes[pl.find_nearest_neighbour_idx(r, pl.poses[: len(pl.nodes)])]
            v = r - nn.pos
            newpos = nn.pos + min(ref_dist(nn.pos, r), 10.0) * v / np.linalg.norm(v)
            if not cc.visible(nn.pos, newpos):
                continue
            newnode, _ = pl.choose_least_cost_parent(Node(newpos), nn, nodes=pl.nodes)
            pl.add_newnode(newnode)
        return pl, store, cc.stats.visible_cnt - vis0

    a, _, vis_a = grow(False)
    b, store, vis_b = grow(True)
    assert np.array_equal(a.poses[:700], b.poses[:700])
    assert [n.cost for n in a.nodes] == [n.cost for n in b.nodes]
    assert [None if n.parent is None else a.nodes.index(n.parent) for n in a.nodes] == \
           [None if n.parent is None else b.nodes.index(n.parent) for n in b.nodes]
    b._sbp_trees[0].sync()                         # the mirror is brought up to date lazily
    assert len(store) == 700 and np.array_equal(store.poses, b.poses[:700])
    assert vis_a == vis_b                          # same Stats accounting of the edge checks


def test_rrdt_array_mirrors_follow_their_arrays(sg):
    """planner_hooks.accelerate_rrdt with real device stores: arrays are mirrored on first
    sight above the size threshold, grow with the planner's prefix, are mirrored again when
    rewritten in place, and answer like np.argmin(np.linalg.norm(...)) (first minimum)."""
    import types

    rng = np.random.default_rng(8)
    cc = sg.ImgCollisionChecker(map_png("room1"))
    eng = types.SimpleNamespace(cc=cc, get_dimension=lambda: 2)

    class P:
        args = types.SimpleNamespace(engine=eng)

        @staticmethod
        def find_nearest_neighbour_idx(pos, poses):
            return int(np.argmin(np.linalg.norm(poses - pos, axis=1)))

    pl = P()
    ref = P.find_nearest_neighbour_idx
    mirrors = sg.planner_hooks.accelerate_rrdt(pl, min_nodes=64)
    big, small = rng.random((5000, 2)) * 500, rng.random((40, 2)) * 500
    big[10] = big[3]                                             # duplicate: first minimum wins
    for n in (64, 700, 701, 5000):
        for _ in range(20):
            x = rng.random(2) * 500
            assert pl.find_nearest_neighbour_idx(x, big[:n]) == ref(x, big[:n])
        assert pl.find_nearest_neighbour_idx(big[10], big[:n]) == 3
    assert len(mirrors) == 1 and len(next(iter(mirrors.values())).store) == 5000
    assert pl.find_nearest_neighbour_idx(small[5], small[:40]) == ref(small[5], small[:40])
    assert len(mirrors) == 1                                     # below the threshold: numpy
    big[:2000] = rng.random((2000, 2)) * 500                     # rewritten in place
    big[1999] += 1.0
    for n in (2000, 1500):
        x = rng.random(2) * 500
        assert pl.find_nearest_neighbour_idx(x, big[:n]) == ref(x, big[:n])
