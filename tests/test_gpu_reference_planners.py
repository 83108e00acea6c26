"""-m gpu: the UNMODIFIED reference planners on real CUDA.

Every test runs a whole seeded planning problem (``Env(args, fixed_seed=0)``) twice with the
reference's own planner / sampler / env code, imported from the staged copy ``oracle/_ref``
(``oracle/stage_ref.py``; ``/root/reference`` in the build container):

  a) with the reference's NumPy / Python collision checker, and
  b) with the ``sbp_env_b200`` checker -- every ``feasible`` / ``visible`` answered by the
     sm_100a kernels through the C ABI -- and, where stated, the ``planner_hooks.accelerate_*``
     instance hooks (node store on the device, nearest / choose-parent / PRM connection
     through the fused kernels),

and asserts identical node arrays, parents, costs, ``c_max`` and Stats counters.  This closes
rows a13-a17 in ONE hop: reference planner <-> CUDA path, no fakes in between.

cfg1 (RRT*, maps/room1.png, start 100,100, goal 350,350, 5 000 nodes; rrtPlanner.py:68-104)
and cfg3 (4-D arm Bi-RRT* on maps/4d.png, L = (30, 30); birrtPlanner.py:36-124) are
BASELINE.json configs[0] and configs[2].
"""
import numpy as np
import pytest

import ref_shims
from _refrun import reference_available, run, same_run, silence_tqdm, tree_state

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_available(),
                                 reason="needs the staged reference (oracle/stage_ref.py)")]

START, GOAL = "100,100", "350,350"
ARM_START, ARM_GOAL = "652.249,276.262,-2.884,-3.038", "832.789,934.662,0.670,1.442"


@pytest.fixture()
def ref(monkeypatch):
    sbp_env = ref_shims.import_reference()
    silence_tqdm(monkeypatch)
    from sbp_env.utils.common import Stats

    Stats.clear_instance()
    yield sbp_env
    Stats.clear_instance()


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


def _img_engine(name):
    from sbp_env import engine

    img = ref_shims.reference_map_path(name)
    return lambda args: engine.ImageEngine(args, img)


def _same_trees(a, b, goal_tree=False):
    pa, pb = a["env"].planner, b["env"].planner
    assert tree_state(pa) == tree_state(pb)
    if goal_tree:
        assert tree_state(pa, pa.goal_tree_nodes) == tree_state(pb, pb.goal_tree_nodes)
        na, nb = len(pa.goal_tree_nodes), len(pb.goal_tree_nodes)
        assert na == nb and np.array_equal(pa.goal_tree_poses[:na], pb.goal_tree_poses[:nb])
        assert pa.found_solution == pb.found_solution


@pytest.mark.parametrize("planner_id,nodes", [("rrt", 1200), ("birrt", 1200), ("informedrrt", 1000)])
def test_seeded_2d_runs_identical_with_cuda_checker(ref, sg, planner_id, nodes):
    fac = _img_engine("room1.png")
    a = run(planner_id, fac, None, START, GOAL, nodes)
    b = run(planner_id, fac, lambda cc: sg.ImgCollisionChecker(cc.image_fname), START, GOAL, nodes)
    same_run(a, b)
    _same_trees(a, b, goal_tree=planner_id == "birrt")
    assert a["counters"][0] > 5 * nodes           # visible() really went through the checker


@pytest.mark.parametrize("planner_id,nodes", [("rrt", 1500), ("informedrrt", 1500)])
def test_accelerate_rrt_hooks_on_cuda(ref, sg, planner_id, nodes):
    """reference planner + reference checker  ==  reference planner + CUDA checker + hooks
    (device node store, NodeStore.nearest, fused near+visible choose-parent)."""
    fac = _img_engine("room1.png")
    a = run(planner_id, fac, None, START, GOAL, nodes)
    stores = []
    b = run(planner_id, fac, lambda cc: sg.ImgCollisionChecker(cc.image_fname), START, GOAL, nodes,
            before_run=lambda env: stores.append(sg.planner_hooks.accelerate_rrt(env.planner)))
    same_run(a, b)
    _same_trees(a, b)
    have = len(stores[0])                          # the device store mirrors the tree (synced lazily:
    assert have >= b["n"] - 1                      # the last appended node may still be pending)
    assert np.array_equal(stores[0].poses, b["poses"][:have])


def test_cfg1_rrt_star_5000_nodes_identical(ref, sg):
    """BASELINE.json configs[0] at full size."""
    import time

    fac = _img_engine("room1.png")
    t0 = time.perf_counter()
    a = run("rrt", fac, None, START, GOAL, 5000)
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = run("rrt", fac, lambda cc: sg.ImgCollisionChecker(cc.image_fname), START, GOAL, 5000,
            before_run=lambda env: sg.planner_hooks.accelerate_rrt(env.planner))
    t_gpu = time.perf_counter() - t0
    same_run(a, b)
    _same_trees(a, b)
    assert a["counters"][2] >= 5000
    print(f"\ncfg1: reference {t_ref:.1f} s, CUDA checker + hooks {t_gpu:.1f} s, c_max {a['c_max']:.4f}")


def test_accelerate_birrt_hooks_on_cuda(ref, sg):
    fac = _img_engine("room1.png")
    a = run("birrt", fac, None, START, GOAL, 1500)
    b = run("birrt", fac, lambda cc: sg.ImgCollisionChecker(cc.image_fname), START, GOAL, 1500,
            before_run=lambda env: sg.planner_hooks.accelerate_birrt(env.planner))
    same_run(a, b)
    _same_trees(a, b, goal_tree=True)


def test_accelerate_rrdt_hook_on_cuda(ref, sg):
    fac = _img_engine("room1.png")
    a = run("rrdt", fac, None, START, GOAL, 500)
    mirrors = []
    b = run("rrdt", fac, lambda cc: sg.ImgCollisionChecker(cc.image_fname), START, GOAL, 500,
            before_run=lambda env: mirrors.append(sg.planner_hooks.accelerate_rrdt(env.planner, min_nodes=8)))
    same_run(a, b)
    assert len(mirrors[0]) >= 1


def test_cfg3_arm_birrt_identical(ref, sg):
    """BASELINE.json configs[2]: 4-D arm Bi-RRT* on maps/4d.png (scaled to 1500 nodes), plain
    CUDA checker and CUDA checker + hooks against the reference's own arm checker."""
    from sbp_env import engine

    img = ref_shims.reference_map_path("4d.png")
    kw = dict(rover_arm_robot_lengths=(30.0, 30.0))
    fac = lambda args: engine.RobotArmEngine(args, img)
    swap = lambda cc: sg.RobotArm4dCollisionChecker(cc.image_fname,
                                                    stick_robot_length_config=cc.stick_robot_length_config)
    a = run("birrt", fac, None, ARM_START, ARM_GOAL, 1500, **kw)
    b = run("birrt", fac, swap, ARM_START, ARM_GOAL, 1500, **kw)
    same_run(a, b)
    _same_trees(a, b, goal_tree=True)
    c = run("birrt", fac, swap, ARM_START, ARM_GOAL, 1500,
            before_run=lambda env: sg.planner_hooks.accelerate_birrt(env.planner), **kw)
    same_run(a, c)
    _same_trees(a, c, goal_tree=True)
    assert a["counters"][0] > 10000


@pytest.mark.parametrize("map_name,start,goal,nodes", [("maze1.png", "25,25", "290,290", 250),
                                                      ("room1.png", START, GOAL, 300)])
def test_prm_build_graph_and_solution_identical(ref, sg, map_name, start, goal, nodes):
    """PRMPlanner.build_graph (prmPlanner.py:80-101) + get_solution (:128-154): the reference's
    double loop vs accelerate_prm (batched radius query + visible_batch on CUDA) -- the same
    DiGraph edge for edge in insertion order with the same weights, the same c_max; and the
    device roadmap (Roadmap + prm_get_solution) agrees with networkx on that graph."""
    fac = _img_engine(map_name)
    outs = []
    for hooked in (False, True):
        r = run("prm", fac, (lambda cc: sg.ImgCollisionChecker(cc.image_fname)) if hooked else None,
                start, goal, nodes)
        pl = r["env"].planner
        store = sg.planner_hooks.accelerate_prm(pl) if hooked else None
        pl.build_graph()
        edges = [(tuple(u.pos), tuple(v.pos), w["weight"]) for u, v, w in pl.graph.edges(data=True)]
        outs.append(dict(run=r, edges=edges, sol=pl.get_solution(), store=store, planner=pl,
                         vis=r["env"].args.engine.cc.stats.visible_cnt))
    same_run(outs[0]["run"], outs[1]["run"])
    assert outs[0]["edges"] == outs[1]["edges"] and len(outs[0]["edges"]) > nodes
    assert outs[0]["sol"] == outs[1]["sol"]
    assert outs[0]["vis"] == outs[1]["vis"]
    # device roadmap on the same candidate rule (radius, strict <) and the same query
    pl, store = outs[1]["planner"], outs[1]["store"]
    cc = pl.args.engine.cc
    n = len(pl.nodes)
    radius = pl.gamma * np.power(np.log(n + 1) / (n + 1), 1 / 2)
    src, dst, vis = sg.prm_connect_radius(cc, store, float(radius))
    g = sg.Roadmap(n)
    g.build_from_edges(src, dst, vis)
    index = {id(nd): i for i, nd in enumerate(pl.nodes)}
    want = sorted((index[id(u)], index[id(v)]) for u, v in pl.graph.edges()
                  if id(u) in index and id(v) in index)
    # networkx folds Node objects of equal position (Node.__hash__ / __eq__ are position based; the
    # goal-biased sampler draws the goal several times) into the first one: fold ours the same way
    first = {}
    canon = np.array([first.setdefault(tuple(p), i) for i, p in enumerate(pl.poses[:n])])
    # (which of the equal Node objects the DiGraph keeps as its key depends on insertion order: fold
    # the reference's pairs onto the first index of every position too)
    want = sorted(set(map(tuple, canon[np.asarray(want, dtype=np.int64).reshape(-1, 2)].tolist())))
    assert sorted(set(map(tuple, canon[np.asarray(g.edges())].tolist()))) == want
    c_max, path = sg.prm_get_solution(cc, store, g, pl.args.sampler.start_pos, pl.args.sampler.goal_pos,
                                      pl.args.epsilon, dist=pl.args.engine.dist, poses=pl.poses[:n],
                                      exclude=(0,))          # node 0 is env.start_pt (prmPlanner.py:113)
    import networkx as nx

    if outs[0]["sol"] == float("inf"):
        assert c_max == float("inf") and path is None
    else:
        # the same end nodes as the reference picked, the same number of hops as networkx, and --
        # when our deterministic fewest-hop path is the one networkx happened to return -- the
        # same c_max (networkx returns an adjacency-order dependent path among equals)
        ref_goal = pl.args.env.goal_pt.parent
        ref_start = next(nd for nd in pl.nodes if nd.parent is pl.args.env.start_pt)
        assert (int(path[0]), int(path[-1])) == (index[id(ref_start)], index[id(ref_goal)])
        ref_path = nx.shortest_path(pl.graph, ref_start, ref_goal)
        assert len(ref_path) == len(path)
        if [index[id(nd)] for nd in ref_path] == [int(i) for i in path]:
            assert c_max == outs[0]["sol"]
