"""-m gpu: rows a13-a17 against REFERENCE-GENERATED trajectories (tests/golden/planner.npz,
written by oracle/gen_golden.py from whole seeded runs of the unmodified planners): the sample
stream of a run is replayed through planner_ops on real CUDA and every intermediate and final
quantity the fixture holds must come out identical.  Needs no reference on the box."""
import math

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
def test_rrt_star_trajectory_replay(sg, tag):
    """RRTPlanner.run_once (rrtPlanner.py:68-104) driven by the recorded sample stream:
    find_nearest_neighbour_idx -> step_from_to -> visible -> choose_least_cost_parent -> append
    (-> linear rewire, rrtPlanner.py:250-266, for "rrtrw") -> goal test."""
    from sbp_env_b200 import planner_ops as ops

    g = golden("planner")
    eps, radius, goal_radius = g[f"{tag}_params"]
    goal = g[f"{tag}_goal"]
    T = int(g[f"{tag}_steps"])
    accept = unpack(g[f"{tag}_accept"], T)
    cc = sg.ImgCollisionChecker(map_png("room1"))
    tree = sg.NodeStore(2)
    poses = np.empty((len(g[f"{tag}_poses"]) + 8, 2))
    poses[0] = g[f"{tag}_poses"][0]
    tree.append(poses[0])
    costs, parents, n, c_max = [0.0], [-1], 1, float("inf")
    v0 = cc.stats.visible_cnt
    for t in range(T):
        rand = g[f"{tag}_rand"][t]
        nn = tree.nearest(rand)
        assert nn == g[f"{tag}_nn"][t], t
        newpos = step_from_to(poses[nn], rand, eps)
        if not cc.visible(poses[nn], newpos):
            assert not accept[t], t
            continue
        assert accept[t], t
        parent, cost, cand = ops.choose_least_cost_parent(cc, tree, costs, newpos, nn, radius, poses=poses[:n])
        new = tree.append(newpos)
        poses[n] = newpos
        costs.append(cost)
        parents.append(parent)
        n += 1
        if tag == "rrtrw":
            ops.rewire(cc, tree, costs, parents, new, radius, poses=poses[:n], candidates=cand)
        if cc.visible(newpos, goal) and dist(newpos, goal) < goal_radius and cost < c_max:
            c_max = cost                          # rrtPlanner.py:96-104
    assert n == len(g[f"{tag}_poses"])
    assert np.array_equal(poses[:n], g[f"{tag}_poses"])
    assert parents == g[f"{tag}_parent"].tolist()
    assert costs == g[f"{tag}_cost"].tolist()
    assert c_max == float(g[f"{tag}_c_max"])
    # Stats: the reference's visible() count of the whole run (rrtPlanner.py:82, :96, :179, :260)
    assert cc.stats.visible_cnt - v0 == int(g[f"{tag}_counters"][0])


def test_birrt_join_replay(sg):
    """birtPlanner.py:82-87 for every accepted node of a seeded Bi-RRT* run up to the join."""
    from sbp_env_b200 import planner_ops as ops

    g = golden("planner")
    cc = sg.ImgCollisionChecker(map_png("room1"))
    stores = [sg.NodeStore(2), sg.NodeStore(2)]
    stores[0].append(g["bi_start"])
    stores[1].append(g["bi_goal"])
    eps = float(g["bi_eps"])
    joined = 0
    for tree_id, newpos, want in zip(g["bi_tree"], g["bi_newpos"], g["bi_join"]):
        stores[int(tree_id)].append(newpos)
        got = ops.birrt_join(cc, stores[1 - int(tree_id)], newpos, eps)
        assert (-1 if got is None else got) == int(want)
        joined += got is not None
    assert joined == 1 and bool(g["bi_found"])


def test_rrdt_neighbour_search_replay(sg):
    """rrdtPlanner.py:685-721 on a snapshot of a seeded RRdT* run's trees."""
    from sbp_env_b200 import planner_ops as ops

    g = golden("planner")
    off = g["rrdt_tree_off"]
    stores = []
    for a, b in zip(off[:-1], off[1:]):
        s = sg.NodeStore(2)
        s.extend(g["rrdt_tree_poses"][a:b])
        stores.append(s)
    eps = float(g["rrdt_eps"])
    ro, res = g["rrdt_res_off"], g["rrdt_res"]
    hits = 0
    for i, (q, own) in enumerate(zip(g["rrdt_Q"], g["rrdt_own"])):
        got = ops.nearest_in_trees(stores, q, eps, own=None if own < 0 else int(own))
        assert [list(x) for x in got] == res[ro[i]:ro[i + 1]].tolist(), i
        hits += len(got)
    assert hits == len(res) > 50


def test_prm_build_graph_replay(sg):
    """PRMPlanner.build_graph (prmPlanner.py:80-101) and get_solution (:128-154) of a seeded run:
    the candidate rule (radius, strict <, index-ascending), the visible pairs, engine.dist weights,
    hop counts of 200 node pairs as networkx computed them on the reference's DiGraph."""
    g = golden("planner")
    poses = g["prm_poses"]
    n = len(poses)
    cc = sg.ImgCollisionChecker(map_png("room1"))
    tree = sg.NodeStore(2)
    tree.extend(poses)
    src, dst, vis = sg.prm_connect_radius(cc, tree, float(g["prm_radius"]))
    s, d, v = src.cpu().numpy(), dst.cpu().numpy(), vis.cpu().numpy()
    # networkx hashes a Node by its position (utils/common.py Node.__hash__ / __eq__): the goal-biased
    # sampler draws the goal position several times, and the DiGraph folds those nodes into the
    # first one.  The product returns the candidate pairs of the reference's double loop itself
    # (prmPlanner.py:91-97, one per pair of Node objects); fold them the same way to compare.
    first = {}
    canon = np.array([first.setdefault(tuple(p), i) for i, p in enumerate(poses)])
    assert (canon != np.arange(n)).any()                       # the fixture does hold duplicates
    got = sorted(set(zip(canon[s[v]].tolist(), canon[d[v]].tolist())))
    # (which of the equal Node objects the DiGraph keeps as its key depends on insertion order: the
    # fixture's pairs are folded onto the first index of every position too)
    want_pairs = [tuple(e) for e in canon[g["prm_edges"]].tolist()]
    assert len(set(want_pairs)) == len(want_pairs)
    order = sorted(range(len(want_pairs)), key=lambda i: want_pairs[i])
    assert got == [want_pairs[i] for i in order]
    w = [sg.euclidean_dist(poses[a], poses[b]) for a, b in got]
    assert w == [float(g["prm_weights"][i]) for i in order]
    road = sg.Roadmap(n)
    road.build_from_edges(src, dst, vis)
    assert road.edges().tolist() == [list(e) for e in sorted(zip(s[v].tolist(), d[v].tolist()))]
    # hop counts: nodes of equal position are one graph node for networkx (0 hops between them)
    pa, pb = canon[g["prm_pairs"][:, 0]], canon[g["prm_pairs"][:, 1]]
    hops = np.asarray(road.shortest_paths(pa.copy(), pb.copy()))
    assert hops.tolist() == g["prm_pair_hops"].tolist()
    # node 0 is env.start_pt itself, which get_nearest_free skips by identity (prmPlanner.py:113)
    c_max, path = sg.prm_get_solution(cc, tree, road, g["prm_query"][0], g["prm_query"][1], float(g["prm_eps"]),
                                      poses=poses, exclude=(0,))
    assert c_max == float(g["prm_c_max"])
    assert (path is None) == (int(g["prm_start_goal"][0]) < 0 or int(g["prm_hops"]) < 0)
