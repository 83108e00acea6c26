"""Build-container only (needs /root/reference; skipped on the GPU box): drives the
UNMODIFIED reference planners (Env, RRT*, Bi-RRT*, Informed RRT*, PRM, the 4-D arm)
through the Python boundary classes of sbp_env_b200 and checks that whole seeded runs
are identical to runs with the reference's own checker -- node arrays, c_max and the
Stats counters.

There is no CUDA device here, so the device layer (libsbpgpu through ctypes) is
replaced by the oracle for these tests: what is being tested is the drop-in contract of
the classes (signatures, Stats accounting, deepcopy, attribute surface, error
behaviour) inside the reference's own control flow.  The kernels themselves are tested
against the same oracle by the -m gpu suite.
"""
import copy
import os
from functools import partialmethod

import numpy as np
import pytest

import c_oracle as orc
import ref_shims
from _util import reference_present

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_present(), reason="needs /root/reference")]


class FakeDeviceMap:
    """Stands in for runtime.DeviceMap: keeps the occupancy grid for the oracle."""

    def __init__(self, free_xy, device=0):
        self.om = orc.Map(np.asarray(free_xy) == 1)
        self.W, self.H = self.om.W, self.om.H


@pytest.fixture()
def ref(monkeypatch):
    sbp_env = ref_shims.import_reference()
    from tqdm import tqdm

    monkeypatch.setattr(tqdm, "__init__", partialmethod(tqdm.__init__, disable=True))
    import sbp_env_b200.collision_checker as gcc
    from sbp_env_b200.runtime import raise_for_flags

    monkeypatch.setattr(gcc, "DeviceMap", FakeDeviceMap)

    def img_feasible(self, P):
        out, st = self.device_map.om.feasible2d(P, return_status=True)
        raise_for_flags(int(np.bitwise_or.reduce(st)) if len(st) else 0, "feasible")
        return out

    def img_visible(self, P, Q):
        out, st = self.device_map.om.visible2d(P, Q, return_status=True)
        raise_for_flags(int(np.bitwise_or.reduce(st)) if len(st) else 0, "visible", nan_is_error=False)
        return out

    def img_first(self, P, Q, with_counts=False):
        out = self.device_map.om.coor_before_collision(P, Q).astype(np.int32)
        if not with_counts:
            return out
        ex = np.maximum(np.abs(out[:, 0] - np.trunc(P[:, 0])), np.abs(out[:, 1] - np.trunc(P[:, 1]))) + 1
        return out, ex.astype(np.int64)

    def arm_feasible(self, C4):
        L0, L1 = self._lengths()
        out, st = self.device_map.om.arm_feasible(C4, L0, L1, return_status=True)
        raise_for_flags(int(np.bitwise_or.reduce(st)) if len(st) else 0, "feasible")
        return out

    def arm_visible(self, C1, C2):
        L0, L1 = self._lengths()
        out, st = self.device_map.om.arm_visible(C1, C2, L0, L1, return_status=True)
        raise_for_flags(int(np.bitwise_or.reduce(st)) if len(st) else 0, "visible")
        return out

    monkeypatch.setattr(gcc.ImgCollisionChecker, "_feasible_host", img_feasible)
    monkeypatch.setattr(gcc.ImgCollisionChecker, "_visible_host", img_visible)
    monkeypatch.setattr(gcc.ImgCollisionChecker, "_first_blocked_host", img_first)
    monkeypatch.setattr(gcc.RobotArm4dCollisionChecker, "_feasible_host", arm_feasible)
    monkeypatch.setattr(gcc.RobotArm4dCollisionChecker, "_visible_host", arm_visible)
    from sbp_env.utils.common import Stats

    Stats.clear_instance()
    yield sbp_env
    Stats.clear_instance()


from _refrun import make_args, run, same_run as _same  # noqa: E402,F401


@pytest.mark.parametrize("planner_id", ["rrt", "birrt", "informedrrt"])
def test_seeded_2d_runs_identical(ref, planner_id):
    from sbp_env import engine
    from sbp_env_b200 import ImgCollisionChecker

    img = ref_shims.reference_map_path("room1.png")
    fac = lambda args: engine.ImageEngine(args, img)
    a = run(planner_id, fac, None, "100,100", "350,350", 250)
    b = run(planner_id, fac, lambda cc: ImgCollisionChecker(cc.image_fname), "100,100", "350,350", 250)
    _same(a, b)
    assert a["counters"][2] >= 250            # valid samples (Bi-RRT splits them over two trees)


def test_sampler_prescreen_keeps_the_seeded_run_and_saves_launches(ref, monkeypatch):
    """pre-screening each 10 000-draw bucket of the reference's
    RandomnessManager (one feasible_batch launch) leaves the seeded run -- draws, accepted
    nodes, c_max, Stats counters -- identical, while the per-state feasible() launches of the
    sampling loop disappear."""
    from sbp_env import engine
    import sbp_env_b200 as sg
    import sbp_env_b200.collision_checker as gcc

    img = ref_shims.reference_map_path("room1.png")
    fac = lambda args: engine.ImageEngine(args, img)
    swap = lambda cc: sg.ImgCollisionChecker(cc.image_fname)
    launches = []
    inner = gcc.ImgCollisionChecker._feasible_host

    def counting(self, P):
        launches.append(len(P))
        return inner(self, P)

    monkeypatch.setattr(gcc.ImgCollisionChecker, "_feasible_host", counting)
    a = run("rrt", fac, swap, "100,100", "350,350", 250)
    plain = list(launches)
    del launches[:]
    b = run("rrt", fac, swap, "100,100", "350,350", 250,
            before_run=lambda env: sg.samplers.attach_prescreen_to_sampler(env.args.sampler))
    _same(a, b)
    assert max(plain) == 1 and len(plain) > 250               # one launch per tested state
    assert launches.count(10000) >= 1                          # the bucket, in one launch
    # what is left are the goal-biased draws (goal_pos is not a bucket row) and env checks
    assert len(launches) < len(plain) // 5


class FakeNodeStore:
    """Stands in for NodeStore (no CUDA device in the build container): same methods, answers
    from numpy / the oracle.  The real store is checked against the same oracle by -m gpu."""

    def __init__(self, dim, om):
        self.dim, self.om, self.rows = dim, om, []
        self.nearest_calls = self.near_visible_calls = 0

    def __len__(self):
        return len(self.rows)

    def extend(self, X):
        self.rows.extend(np.asarray(X, dtype=np.float64).reshape(-1, self.dim).copy())

    def append(self, x):
        self.extend(np.asarray(x).reshape(1, self.dim))
        return len(self.rows) - 1

    def poses_rows(self, idx):
        return np.array([self.rows[int(i)] for i in idx]).reshape(-1, self.dim)

    def nearest(self, x):
        self.nearest_calls += 1
        return int(orc.knn(np.array(self.rows), np.asarray(x, dtype=np.float64).reshape(1, self.dim), 1)[0][0, 0])

    def near_visible(self, cc, x, r, metric="xy", closed=True, direction=0, cap=256):
        self.near_visible_calls += 1
        P = np.array(self.rows)
        x = np.asarray(x, dtype=np.float64).reshape(1, self.dim)
        rp, idx = orc.near(P, x, r, metric, closed)
        X = np.repeat(x[:, :2], len(idx), 0)
        vis = self.om.visible2d(X, P[idx][:, :2]) if direction == 0 else self.om.visible2d(P[idx][:, :2], X)
        cc.stats.visible_cnt += len(idx)
        return idx, vis


@pytest.mark.parametrize("planner_id", ["rrt", "informedrrt"])
def test_accelerate_rrt_hooks_keep_the_seeded_run(ref, planner_id):
    """planner_hooks.accelerate_rrt on an unmodified reference planner object
    -- nearest through the node store, choose_least_cost_parent through one fused
    near + visible query -- gives the identical seeded run (nodes, c_max, Stats), and the
    planner's own O(N) loop over engine.dist no longer runs."""
    from sbp_env import engine
    import sbp_env_b200 as sg

    img = ref_shims.reference_map_path("room1.png")
    fac = lambda args: engine.ImageEngine(args, img)
    swap = lambda cc: sg.ImgCollisionChecker(cc.image_fname)
    dist_calls = []

    def counting_engine(args):
        e = engine.ImageEngine(args, img)
        inner = e.dist
        e.dist = lambda *a: (dist_calls.append(1), inner(*a))[1]
        return e

    a = run(planner_id, counting_engine, swap, "100,100", "350,350", 300)
    plain_calls = len(dist_calls)
    del dist_calls[:]
    stores = []

    def hook(env):
        om = orc.Map(np.asarray(env.args.engine.cc._img) == 1)
        stores.append(sg.planner_hooks.accelerate_rrt(env.planner, store=FakeNodeStore(2, om)))

    b = run(planner_id, counting_engine, swap, "100,100", "350,350", 300, before_run=hook)
    _same(a, b)
    pa, pb = a["env"].planner, b["env"].planner
    assert [n.cost for n in pa.nodes] == [n.cost for n in pb.nodes]
    assert [None if n.parent is None else tuple(n.parent.pos) for n in pa.nodes] == \
           [None if n.parent is None else tuple(n.parent.pos) for n in pb.nodes]
    st = stores[0]
    assert len(st) >= len(pb.nodes) - 1 and st.near_visible_calls >= 300 and st.nearest_calls >= 300
    assert len(dist_calls) < plain_calls // 4          # the O(N) engine.dist loop is gone


def test_accelerate_birrt_hooks_keep_the_seeded_run(ref):
    """Bi-RRT*: two trees, each mirrored into its own store (picked by array identity, synced
    lazily because run_once appends to the arrays directly); the join and the re-parenting walk
    after it run unchanged.  Identical seeded run."""
    from sbp_env import engine
    import sbp_env_b200 as sg

    img = ref_shims.reference_map_path("room1.png")
    fac = lambda args: engine.ImageEngine(args, img)
    swap = lambda cc: sg.ImgCollisionChecker(cc.image_fname)
    a = run("birrt", fac, swap, "100,100", "350,350", 400)
    stores = []

    def hook(env):
        om = orc.Map(np.asarray(env.args.engine.cc._img) == 1)
        stores.extend(sg.planner_hooks.accelerate_birrt(env.planner, stores=(FakeNodeStore(2, om), FakeNodeStore(2, om))))

    b = run("birrt", fac, swap, "100,100", "350,350", 400, before_run=hook)
    _same(a, b)
    pa, pb = a["env"].planner, b["env"].planner
    assert [n.cost for n in pa.nodes] == [n.cost for n in pb.nodes]
    assert np.array_equal(pa.goal_tree_poses[: len(pa.goal_tree_nodes)], pb.goal_tree_poses[: len(pb.goal_tree_nodes)])
    assert pb.found_solution == pa.found_solution
    assert all(s.near_visible_calls > 50 and s.nearest_calls > 50 for s in stores)


def test_accelerate_rrdt_hook_keeps_the_seeded_run(ref):
    """RRdT*: every tree's pose array that is searched gets a mirror on first sight (the
    threshold is lowered to 8 rows here so that the disjoint trees are covered too); the
    seeded run is identical."""
    from sbp_env import engine
    import sbp_env_b200 as sg

    img = ref_shims.reference_map_path("room1.png")
    fac = lambda args: engine.ImageEngine(args, img)
    swap = lambda cc: sg.ImgCollisionChecker(cc.image_fname)
    a = run("rrdt", fac, swap, "100,100", "350,350", 350)
    made = []

    class FakeTruncatable(FakeNodeStore):
        def truncate(self, n):
            del self.rows[n:]

    def hook(env):
        om = orc.Map(np.asarray(env.args.engine.cc._img) == 1)

        def factory(base):
            made.append(FakeTruncatable(2, om))
            return made[-1]

        sg.planner_hooks.accelerate_rrdt(env.planner, min_nodes=8, store_factory=factory)

    b = run("rrdt", fac, swap, "100,100", "350,350", 350, before_run=hook)
    _same(a, b)
    assert len(made) >= 1 and sum(s.nearest_calls for s in made) > 100


def test_accelerate_prm_builds_the_identical_graph(ref, monkeypatch):
    """planner_hooks.accelerate_prm: build_graph through the batched radius query and one
    visible_batch launch per block gives the reference's DiGraph -- same edges in the same
    insertion order, same weights -- and therefore the same get_solution; the per-pair
    cc.visible calls of the reference's double loop are gone."""
    from sbp_env import engine
    import sbp_env_b200 as sg
    import sbp_env_b200.collision_checker as gcc

    img = ref_shims.reference_map_path("maze1.png")
    fac = lambda args: sg.ImageEngine(args, img)
    single, batched = [], []
    inner = gcc.ImgCollisionChecker._visible_host

    def counting(self, P, Q):
        (single if len(P) == 1 else batched).append(len(P))
        return inner(self, P, Q)

    monkeypatch.setattr(gcc.ImgCollisionChecker, "_visible_host", counting)

    class FakeNear(FakeNodeStore):
        def truncate(self, n):
            del self.rows[n:]

        def near(self, X, r, metric="full", closed=False):
            return orc.near(np.array(self.rows), np.asarray(X, dtype=np.float64), r, metric, closed)

    outs = []
    for hooked in (False, True):
        del single[:], batched[:]
        r = run("prm", fac, None, "25,25", "290,290", 70)
        pl = r["env"].planner
        if hooked:
            om = orc.Map(np.asarray(pl.args.engine.cc._img) == 1)
            sg.planner_hooks.accelerate_prm(pl, store=FakeNear(2, om), chunk_rows=16)
        before = len(single)
        pl.build_graph()
        in_build = len(single) - before
        outs.append(dict(run=r, edges=[(tuple(u.pos), tuple(v.pos), w["weight"]) for u, v, w in pl.graph.edges(data=True)],
                         sol=pl.get_solution(), singles=in_build, batches=list(batched),
                         vis_cnt=r["env"].args.engine.cc.stats.visible_cnt))
    _same(outs[0]["run"], outs[1]["run"])
    assert outs[0]["edges"] == outs[1]["edges"] and len(outs[0]["edges"]) > 50
    assert outs[0]["sol"] == outs[1]["sol"]
    assert outs[0]["singles"] > 1000 and outs[1]["singles"] == 0 and len(outs[1]["batches"]) == 5
    assert outs[0]["vis_cnt"] == outs[1]["vis_cnt"]


def test_seeded_arm_run_identical(ref):
    from sbp_env import engine
    from sbp_env_b200 import RobotArm4dCollisionChecker

    img = ref_shims.reference_map_path("4d.png")
    kw = dict(rover_arm_robot_lengths=(30.0, 30.0))
    fac = lambda args: engine.RobotArmEngine(args, img)
    start, goal = "652.249,276.262,-2.884,-3.038", "832.789,934.662,0.670,1.442"
    swap = lambda cc: RobotArm4dCollisionChecker(cc.image_fname, stick_robot_length_config=cc.stick_robot_length_config)
    a = run("birrt", fac, None, start, goal, 120, **kw)
    b = run("birrt", fac, swap, start, goal, 120, **kw)
    _same(a, b)


def test_engine_mirrors_and_prm(ref):
    """The Engine look-alikes carry an unmodified reference PRM run: sampling, build_graph
    (radius PRM*, prmPlanner.py:80-101) and get_solution."""
    from sbp_env import engine
    import sbp_env_b200 as sg

    img = ref_shims.reference_map_path("maze1.png")
    ref_engine = lambda args: engine.ImageEngine(args, img)
    my_engine = lambda args: sg.ImageEngine(args, img)
    outs = []
    for fac in (ref_engine, my_engine):
        r = run("prm", fac, None, "25,25", "290,290", 60)
        pl = r["env"].planner
        pl.build_graph()
        sol = pl.get_solution()
        outs.append((r, sorted((tuple(u.pos), tuple(v.pos)) for u, v in pl.graph.edges()), sol))
        e = r["env"].args.engine
        assert e.get_dimension() == 2 and e.lower.tolist() == [0, 0] and e.upper.tolist() == [322, 322]
        assert e.dist(np.array([0.0, 0.0]), np.array([3.0, 4.0])) == 5.0
    _same(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1] and len(outs[0][1]) > 0
    assert outs[0][2] == outs[1][2]


def test_checker_surface_matches_reference(ref):
    """Attribute surface, Stats binding, deepcopy and error behaviour side by side."""
    from sbp_env.collisionChecker import ImgCollisionChecker as RefCC
    from sbp_env.utils.common import Stats
    from sbp_env_b200 import ImgCollisionChecker as GpuCC

    img = ref_shims.reference_map_path("room1.png")
    Stats.clear_instance()
    r, g = RefCC(img), GpuCC(img)
    assert g.stats is r.stats is Stats.get_instance()             # same singleton
    assert np.array_equal(r._img, g._img) and r._img.dtype == g._img.dtype
    assert np.array_equal(r.image, g.image) and r.image_fname == g.image_fname
    rng = np.random.default_rng(0)
    for _ in range(300):
        p, q = rng.random(2) * [700, 500] - 60, rng.random(2) * [700, 500] - 60
        assert bool(r.feasible(p)) == bool(g.feasible(p))
        assert bool(r.visible(p, q)) == bool(g.visible(p, q))
        assert tuple(r.get_coor_before_collision(p, q)) == g.get_coor_before_collision(p, q)
        assert r.get_line(p, q) == g.get_line(p, q)
    assert Stats.get_instance().visible_cnt == 600
    for bad, exc in (((float("nan"), 1.0), ValueError), ((float("inf"), 1.0), OverflowError)):
        for cc in (r, g):
            with pytest.raises(exc):
                cc.feasible(bad)
    assert r.visible((float("nan"), 1.0), (5.0, 5.0)) is False and g.visible((float("nan"), 1.0), (5.0, 5.0)) is False
    g2 = copy.deepcopy(g)
    assert g2.visible((100, 100), (105, 104)) == r.visible((100, 100), (105, 104))
