"""-m gpu: the device-resident RRT* tree (sbp_tree_extend_host: choose parent +
append + rewire in one launch against device cost / parent arrays) -- replay of the reference's
recorded trajectories (tests/golden/planner.npz), and against the host path on a large random tree."""
import math
import time

import numpy as np
import pytest

from _util import golden, load_map, map_png, unpack

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


def dist(p1, p2):
    p = p1 - p2                                  # engine.euclidean_dist, engine.py:13-24
    return math.sqrt(p[0] ** 2 + p[1] ** 2)


def step_from_to(p1, p2, eps):
    if np.isclose(p1, p2).all():                 # env.py:135-151
        return p2
    unit = p2 - p1
    unit = unit / np.linalg.norm(unit)
    return p1 + min(dist(p1, p2), eps) * unit


@pytest.mark.parametrize("tag", ["rrt", "rrtrw"])
def test_reference_trajectory_replay_on_the_device_tree(sg, tag):
    """RRTPlanner.run_once (rrtPlanner.py:68-104) driven by the reference's recorded sample stream:
    nearest -> step_from_to -> visible -> DeviceTree.extend (choose parent + append [+ rewire]) ->
    goal test; every parent, cost, c_max and the visible() counter equal the reference's run, and
    the DEVICE arrays hold the same costs / parents bit for bit."""
    g = golden("planner")
    eps, radius, goal_radius = g[f"{tag}_params"]
    goal = g[f"{tag}_goal"]
    T = int(g[f"{tag}_steps"])
    accept = unpack(g[f"{tag}_accept"], T)
    cc = sg.ImgCollisionChecker(map_png("room1"))
    store = sg.NodeStore(2)
    tree = sg.DeviceTree(cc, store)
    tree.add_root(g[f"{tag}_poses"][0])
    c_max = float("inf")
    v0 = cc.stats.visible_cnt
    for t in range(T):
        rand = g[f"{tag}_rand"][t]
        nn = store.nearest(rand)
        assert nn == g[f"{tag}_nn"][t], t
        newpos = step_from_to(tree.poses[nn], rand, eps)
        if not cc.visible(tree.poses[nn], newpos):
            assert not accept[t], t
            continue
        assert accept[t], t
        parent, cost, rewired = tree.extend(newpos, nn, radius, rewire=(tag == "rrtrw"))
        if cc.visible(newpos, goal) and dist(newpos, goal) < goal_radius and cost < c_max:
            c_max = cost                          # rrtPlanner.py:96-104
    n = len(tree)
    assert n == len(g[f"{tag}_poses"]) == len(store)
    assert np.array_equal(np.asarray(tree.poses), g[f"{tag}_poses"])
    assert tree.parents == g[f"{tag}_parent"].tolist()
    assert tree.costs == g[f"{tag}_cost"].tolist()
    assert c_max == float(g[f"{tag}_c_max"])
    assert cc.stats.visible_cnt - v0 == int(g[f"{tag}_counters"][0])
    dc, dp = tree.read_device()
    assert dc.tolist() == tree.costs and dp.tolist() == tree.parents
    assert np.array_equal(store.poses_rows(np.arange(n)), g[f"{tag}_poses"])
    assert tree.ambiguous <= n // 50               # the device decided (nearly) every sample itself


def test_device_tree_equals_host_path_on_a_large_tree(sg):
    """20 000 nodes grown on intel_lab with radius 14 (dozens of candidates per sample, rewire on),
    the device tree against planner_ops on a second store: same parents, costs, rewired sets."""
    from sbp_env_b200 import planner_ops as ops

    free = load_map("intel_lab")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("intel_lab"))
    rng = np.random.default_rng(12)
    eps, radius, n_nodes = 6.0, 14.0, 6000
    sa, sb = sg.NodeStore(2), sg.NodeStore(2)
    tree = sg.DeviceTree(cc, sa)
    root = np.array([300.0, 300.0])
    while not cc.feasible(root):
        root = rng.random(2) * [W, H]
    tree.add_root(root)
    sb.append(root)
    costs, parents, poses = [0.0], [-1], [root.copy()]
    t_dev = t_host = 0.0
    while len(poses) < n_nodes:
        rand = rng.random(2) * [W, H]
        nn = sa.nearest(rand)
        newpos = step_from_to(poses[nn], rand, eps)
        if not cc.feasible(newpos) or not cc.visible(poses[nn], newpos):
            continue
        t0 = time.perf_counter()
        pa, ca, ra = tree.extend(newpos, nn, radius, rewire=True)
        t1 = time.perf_counter()
        P = np.asarray(poses)
        pb, cb, cand = ops.choose_least_cost_parent(cc, sb, costs, newpos, nn, radius, poses=P)
        new = sb.append(newpos)
        poses.append(newpos.copy())
        costs.append(cb)
        parents.append(pb)
        rb = ops.rewire(cc, sb, costs, parents, new, radius, poses=np.asarray(poses), candidates=cand)
        t2 = time.perf_counter()
        t_dev += t1 - t0
        t_host += t2 - t1
        assert (pa, ca, ra) == (pb, cb, sorted(rb)), len(poses)
    assert tree.costs == costs and tree.parents == parents
    dc, dp = tree.read_device()
    assert dc.tolist() == costs and dp.tolist() == parents
    assert sum(1 for p in parents if p >= 0) == n_nodes - 1
    print(f"device tree {1e6 * t_dev / n_nodes:.1f} us / sample, host path {1e6 * t_host / n_nodes:.1f} us / sample, "
          f"{tree.ambiguous} samples settled on the host")


def test_device_tree_no_parent_and_growth(sg):
    cc = sg.ImgCollisionChecker(map_png("room1"))
    store = sg.NodeStore(2)
    tree = sg.DeviceTree(cc, store)
    tree.add_root([100.0, 100.0])
    with pytest.raises(LookupError):
        tree.extend(np.array([300.0, 300.0]), None, 11.0)
    assert len(tree) == 1 and len(store) == 1
    # grows past the initial capacity (1024 nodes) of the store and of the cost / parent arrays
    rng = np.random.default_rng(1)
    pos = np.array([100.0, 100.0])
    for i in range(1500):
        nxt = pos + rng.standard_normal(2) * 0.5
        if not cc.feasible(nxt) or not cc.visible(pos, nxt):
            continue
        tree.extend(nxt, len(tree) - 1, 3.0)
        pos = nxt
    assert len(tree) == len(store) > 1100
    dc, dp = tree.read_device()
    assert dc.tolist() == tree.costs and dp.tolist() == tree.parents
