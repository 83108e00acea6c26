"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

Every expected value written here is produced by calling the reference's own
functions (imported through oracle/ref_shims.py, nothing copied):

  maps.npz      `_img` of ImgCollisionChecker for the 8 shipped maps, bit-packed
                (collisionChecker.py:88-108) -- derived occupancy data, not images
  img_cc.npz    ImgCollisionChecker.feasible / visible / get_coor_before_collision /
                get_line on seeded random + adversarial inputs (:118-215)
  arm_cc.npz    RobotArm4dCollisionChecker.feasible / visible / create_ranges /
                _interpolate_configs (:346-426)
  nn.npz        RRTPlanner.find_nearest_neighbour_idx (rrtPlanner.py:268-279),
                prmPlanner.nearest_neighbours (prmPlanner.py:28-44),
                engine.euclidean_dist (engine.py:13-24), np.linalg.norm rows

  planner.npz   whole seeded planner runs of the UNMODIFIED planners (rows a13-a17): RRT* on
                room1 -- the sample stream, per step nearest index / accepted flag, and the final
                parent / cost arrays (rrtPlanner.py:68-104, :173-196); the same run with the
                LINEAR rewire variant switched on (rrtPlanner.py:250-266, `use_rtree=False`);
                the Bi-RRT join test per accepted node (birrtPlanner.py:82-87); RRdT*'s
                find_nearest_node_from_neighbour on a snapshot of its trees (rrdtPlanner.py:
                685-721); PRMPlanner.build_graph edge list + get_solution (prmPlanner.py:80-154)

The GPU box has no /root/reference; the fixture tests there read only these files.
"""
from __future__ import annotations

import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
MAPS = ["room1", "intel_lab", "maze1", "maze2", "4d", "noise", "test", "blank"]


def test_image_png():
    """The 4x6 fixture of the reference's tests (tests/common_vars.py:13-29)."""
    from PIL import Image

    arr = np.array([[255, 255, 0, 0, 100, 200]] * 4, dtype=np.uint8)
    f = io.BytesIO()
    Image.fromarray(arr).save(f, "png")
    f.name = "test.png"
    f.seek(0)
    return f


def edge_sets(rng, W, H, n_uniform, n_short, n_adv):
    """Seeded endpoint pairs: uniform, RRT-like short, and adversarial."""
    P_u = rng.random((n_uniform, 2)) * [W, H]
    Q_u = rng.random((n_uniform, 2)) * [W, H]
    P_s = rng.random((n_short, 2)) * [W, H]
    ang = rng.random(n_short) * 2 * np.pi
    ln = rng.random(n_short) * 14.0
    Q_s = P_s + np.stack([np.cos(ang), np.sin(ang)], 1) * ln[:, None]
    # adversarial: wrap-around negatives, out of range, integer-exact, (-1,0) truncation,
    # same pixel, axis aligned, exact diagonals, NaN endpoints
    lo, hi = -1.6, 1.6
    P_a = (rng.random((n_adv, 2)) * (hi - lo) + lo) * [W, H]
    Q_a = (rng.random((n_adv, 2)) * (hi - lo) + lo) * [W, H]
    k = n_adv // 8
    P_a[:k] = np.floor(P_a[:k])
    Q_a[:k] = np.floor(Q_a[:k])
    P_a[k:2 * k] = -rng.random((k, 2))                    # int(-0.x) == 0
    Q_a[2 * k:3 * k] = P_a[2 * k:3 * k] + rng.random((k, 2)) * 0.9  # mostly one pixel
    Q_a[3 * k:4 * k, 0] = P_a[3 * k:4 * k, 0]             # vertical
    Q_a[4 * k:5 * k, 1] = P_a[4 * k:5 * k, 1]             # horizontal
    dgl = np.floor(rng.random(k) * 60) * rng.choice([-1, 1], k)
    P_a[5 * k:6 * k] = np.floor(rng.random((k, 2)) * [W, H])
    Q_a[5 * k:6 * k] = P_a[5 * k:6 * k] + np.stack([dgl, dgl * rng.choice([-1, 1], k)], 1)
    P_a[6 * k:6 * k + 4] = np.nan
    Q_a[6 * k + 4:6 * k + 8, 1] = np.nan
    P_a[6 * k + 8] = [-0.0, -0.0]
    P_a[6 * k + 9] = [W - 1e-9, H - 1e-9]
    P_a[6 * k + 10] = [W, 0]
    P_a[6 * k + 11] = [-W, -H]
    P_a[6 * k + 12] = [-W - 1e-9, 0]
    P_a[6 * k + 13] = [-W - 1, 0]
    return (np.concatenate([P_u, P_s, P_a]), np.concatenate([Q_u, Q_s, Q_a]))


def ref_visible(cc, P, Q):
    return np.array([bool(cc.visible(p, q)) for p, q in zip(P, Q)], dtype=bool)


def gen_maps(ref):
    from sbp_env.collisionChecker import ImgCollisionChecker

    out = {}
    for name in MAPS:
        cc = ImgCollisionChecker(ref_shims.reference_map_path(name + ".png"))
        free = cc._img == 1                      # [W][H]
        out[name + "_shape"] = np.array(free.shape, np.int64)
        out[name + "_bits"] = np.packbits(free.reshape(-1))
    cc = ImgCollisionChecker(test_image_png())
    free = cc._img == 1
    out["kat4x6_shape"] = np.array(free.shape, np.int64)
    out["kat4x6_bits"] = np.packbits(free.reshape(-1))
    np.savez_compressed(os.path.join(GOLDEN, "maps.npz"), **out)
    print("maps.npz", {k: v.tolist() for k, v in out.items() if k.endswith("_shape")})


def gen_img_cc(ref):
    from sbp_env.collisionChecker import ImgCollisionChecker

    out = {}
    for mi, name in enumerate(["room1", "intel_lab", "maze1", "noise", "kat4x6"]):
        src = test_image_png() if name == "kat4x6" else ref_shims.reference_map_path(name + ".png")
        cc = ImgCollisionChecker(src)
        W, H = cc._img.shape
        rng = np.random.default_rng(1000 + mi)
        small = name == "kat4x6"
        P, Q = edge_sets(rng, W, H, 300 if small else 3000, 300 if small else 3000, 800)
        vis = ref_visible(cc, P, Q)
        # feasible: reference raises on NaN -> keep finite rows, add huge magnitudes
        pts = np.concatenate([P, Q])
        pts = pts[np.isfinite(pts).all(1)]
        huge = np.array([[1e6, 1.0], [1.0, -1e12], [1e19, 0.0], [-1e300, 5.0], [1e300, 1e300],
                         [2.0 ** 31, 0.0], [-(2.0 ** 31) - 1, 0.0], [2.0 ** 63, 1.0],
                         [-(2.0 ** 63), 1.0], [np.nextafter(2.0 ** 63, 0), 1.0],
                         [np.nextafter(-(2.0 ** 63), -np.inf), 1.0]])
        pts = np.concatenate([pts, huge])
        feas = np.zeros(len(pts), bool)
        fstat = np.zeros(len(pts), np.uint8)      # 2: the reference raised OverflowError
        for i, p in enumerate(pts):
            try:
                feas[i] = bool(cc.feasible(p))
            except OverflowError:                 # |int(v)| >= 2**63: numpy cannot index
                fstat[i] = 2
        out[f"{name}_feasible_status"] = fstat
        finite = np.isfinite(P).all(1) & np.isfinite(Q).all(1)
        sel = np.flatnonzero(finite)[:: max(1, int(finite.sum()) // 1500)]
        cbc = np.array([cc.get_coor_before_collision(P[i], Q[i]) for i in sel], dtype=np.float64)
        # get_coor_before_collision returns pos2 itself (floats) for an empty walk; never
        # happens (lines have >= 1 pixel), so the values are integral
        assert np.all(cbc == np.floor(cbc))
        out[f"{name}_P"], out[f"{name}_Q"] = P, Q
        out[f"{name}_visible"] = np.packbits(vis)
        out[f"{name}_pts"] = pts
        out[f"{name}_feasible"] = np.packbits(feas)
        out[f"{name}_cbc_sel"] = sel.astype(np.int64)
        out[f"{name}_cbc"] = cbc.astype(np.int64)
        print(name, "edges", len(P), "visible", int(vis.sum()), "pts", len(pts), "feasible", int(feas.sum()))
    # get_line samples (start->end pixel lists), ragged -> offsets
    rng = np.random.default_rng(77)
    S = np.floor((rng.random((400, 2)) - 0.3) * 90)
    E = np.floor((rng.random((400, 2)) - 0.3) * 90)
    S[:20] = E[:20]                                        # single pixel
    E[20:40, 0] = S[20:40, 0]
    E[40:60, 1] = S[40:60, 1]
    S = np.concatenate([S, rng.random((100, 2)) * 50 - 10])   # fractional -> int() truncation
    E = np.concatenate([E, rng.random((100, 2)) * 50 - 10])
    lines = [np.array(ImgCollisionChecker.get_line(s, e), dtype=np.int64).reshape(-1, 2) for s, e in zip(S, E)]
    out["line_S"], out["line_E"] = S, E
    out["line_off"] = np.cumsum([0] + [len(l) for l in lines]).astype(np.int64)
    out["line_xy"] = np.concatenate(lines)
    np.savez_compressed(os.path.join(GOLDEN, "img_cc.npz"), **out)


def gen_arm_cc(ref):
    from sbp_env.collisionChecker import RobotArm4dCollisionChecker

    out = {}
    cases = [("4d", (30.0, 30.0), 4000, 250, 1500), ("4d", (5.5, 7.25), 1500, 100, 500),
             ("kat4x6", (0.1, 0.1), 600, 200, 400), ("maze1", (12.0, 3.0), 1500, 100, 500)]
    for ci, (name, L, n_feas, n_uni, n_short) in enumerate(cases):
        src = test_image_png() if name == "kat4x6" else ref_shims.reference_map_path(name + ".png")
        cc = RobotArm4dCollisionChecker(src, stick_robot_length_config=list(L))
        W, H = cc._img.shape
        rng = np.random.default_rng(2000 + ci)
        lo = np.array([0, 0, -np.pi, -np.pi])
        hi = np.array([W, H, np.pi, np.pi])
        C = rng.random((n_feas, 4)) * (hi - lo) + lo
        # adversarial configs: axis-aligned angles (FK lands on exact integers), base
        # outside / negative, integer bases
        A = rng.random((200, 4)) * (hi - lo) + lo
        A[:40, 2] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi], 40)
        A[20:60, 3] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi, -np.pi], 40)
        A[60:100, :2] = np.floor(A[60:100, :2])
        A[100:140, :2] = (rng.random((40, 2)) * 3.2 - 1.6) * [W, H]
        A[140:160, :2] = -rng.random((20, 2))
        A[160:180, 2:] = rng.random((20, 2)) * 200 - 100      # many turns
        C = np.concatenate([C, A])
        feas = np.array([bool(cc.feasible(c)) for c in C], dtype=bool)
        C1 = rng.random((n_uni, 4)) * (hi - lo) + lo
        C2 = rng.random((n_uni, 4)) * (hi - lo) + lo
        S1 = rng.random((n_short, 4)) * (hi - lo) + lo
        # prefer feasible starts so that short edges are not all trivially blocked
        S1f = np.array([bool(cc.feasible(c)) for c in S1])
        S1 = np.concatenate([S1[S1f], S1[~S1f][: max(10, n_short // 10)]])
        d = rng.standard_normal((len(S1), 4))
        d /= np.linalg.norm(d, axis=1)[:, None]
        S2 = S1 + d * (rng.random((len(S1), 1)) * [12.0, 12.0, 1.5, 1.5])
        S2[:30, :2] = S1[:30, :2]                             # N == 1 -> duplicated config
        E1 = np.concatenate([C1, S1, A[:100]])
        E2 = np.concatenate([C2, S2, A[100:200]])
        vis = np.array([bool(cc.visible(a, b)) for a, b in zip(E1, E2)], dtype=bool)
        tag = f"c{ci}"
        out[tag + "_map"] = np.array(name)
        out[tag + "_L"] = np.array(L)
        out[tag + "_C"] = C
        out[tag + "_feasible"] = np.packbits(feas)
        out[tag + "_E1"], out[tag + "_E2"] = E1, E2
        out[tag + "_visible"] = np.packbits(vis)
        print(tag, name, L, "configs", len(C), "feasible", int(feas.sum()), "edges", len(E1), "visible", int(vis.sum()))
    out["n_cases"] = np.array(len(cases))
    # create_ranges / _interpolate_configs samples
    rng = np.random.default_rng(31)
    cc = RobotArm4dCollisionChecker(test_image_png(), stick_robot_length_config=[0.1, 0.1])
    st = rng.random((50, 2)) * 8 - 4
    en = rng.random((50, 2)) * 8 - 4
    Ns = rng.integers(2, 40, 50)
    cr = [cc.create_ranges(s, e, int(N)) for s, e, N in zip(st, en, Ns)]
    out["cr_start"], out["cr_stop"], out["cr_N"] = st, en, Ns.astype(np.int64)
    out["cr_off"] = np.cumsum([0] + [c.size for c in cr]).astype(np.int64)
    out["cr_val"] = np.concatenate([c.reshape(-1) for c in cr])
    a = rng.random((60, 4)) * [40, 40, 6, 6] - [5, 5, 3, 3]
    b = rng.random((60, 4)) * [40, 40, 6, 6] - [5, 5, 3, 3]
    b[:5, :2] = a[:5, :2]
    ic = [cc._interpolate_configs(x, y) for x, y in zip(a, b)]
    out["ic_a"], out["ic_b"] = a, b
    out["ic_off"] = np.cumsum([0] + [len(c) for c in ic]).astype(np.int64)
    out["ic_val"] = np.concatenate(ic)
    np.savez_compressed(os.path.join(GOLDEN, "arm_cc.npz"), **out)


def gen_nn(ref):
    from sbp_env.engine import euclidean_dist
    from sbp_env.planners import prmPlanner
    from sbp_env.planners.rrtPlanner import RRTPlanner

    out = {}
    for d in (2, 3, 4):
        rng = np.random.default_rng(3000 + d)
        N, m = 2500, 120
        poses = rng.random((N, d)) * 500
        # exact ties: lattice points and duplicated rows
        poses[:400] = np.floor(poses[:400] / 25) * 25
        poses[400:450] = poses[350:400]
        X = rng.random((m, d)) * 500
        X[:30] = np.floor(X[:30] / 25) * 25 + 12.5           # equidistant to lattice rows
        X[30:40] = poses[100:110]                            # zero distance, duplicates
        nodes = list(range(N))
        out[f"d{d}_poses"], out[f"d{d}_X"] = poses, X
        out[f"d{d}_nearest"] = np.array(
            [RRTPlanner.find_nearest_neighbour_idx(x, poses) for x in X], np.int64)
        dist = np.stack([np.linalg.norm(poses - x, axis=1) for x in X])
        out[f"d{d}_dist8"] = dist[:8]                        # full rows for 8 queries only (size)
        knn20 = np.stack([np.argsort(r, kind="stable")[:20] for r in dist]).astype(np.int64)
        out[f"d{d}_knn20"] = knn20
        out[f"d{d}_knn20_dist"] = np.take_along_axis(dist, knn20, 1)
        for ri, r in enumerate((18.0, 60.0)):
            rows = [np.array(prmPlanner.nearest_neighbours(nodes, poses, x, r), np.int64) for x in X]
            out[f"d{d}_prm_r{ri}"] = np.array(r)
            out[f"d{d}_prm_r{ri}_off"] = np.cumsum([0] + [len(q) for q in rows]).astype(np.int64)
            out[f"d{d}_prm_r{ri}_idx"] = np.concatenate(rows) if rows else np.zeros(0, np.int64)
        # engine.dist gate of choose_least_cost_parent (rrtPlanner.py:177-181): first two
        # dims, scalar pow, `<=`
        mm = 40
        xyd = np.array([[euclidean_dist(x, p) for p in poses] for x in X[:mm]])
        out[f"d{d}_xy_dist8"] = xyd[:8]
        r = 25.0
        rows = [np.flatnonzero(row <= r).astype(np.int64) for row in xyd]
        out[f"d{d}_xy_r"] = np.array(r)
        out[f"d{d}_xy_off"] = np.cumsum([0] + [len(q) for q in rows]).astype(np.int64)
        out[f"d{d}_xy_idx"] = np.concatenate(rows)
        print("nn d", d, "nearest", out[f"d{d}_nearest"][:5], "ties in first row", int((dist[0] == dist[0].min()).sum()))
    np.savez_compressed(os.path.join(GOLDEN, "nn.npz"), **out)


def _quiet_tqdm():
    from functools import partialmethod

    from tqdm import tqdm

    tqdm.__init__ = partialmethod(tqdm.__init__, disable=True)


def gen_planner(ref):
    """Trajectory fixtures: everything below is produced by the reference's own planner code."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
    import _refrun
    from sbp_env import engine
    from sbp_env.planners.rrtPlanner import RRTPlanner
    from sbp_env.utils.common import Node

    _quiet_tqdm()
    out = {}
    room1 = ref_shims.reference_map_path("room1.png")
    fac = lambda args: engine.ImageEngine(args, room1)

    def index_map(nodes):
        return {id(n): i for i, n in enumerate(nodes)}

    def parents_of(nodes, extra=()):
        ix = index_map(nodes)
        for i, n in enumerate(extra):
            ix.setdefault(id(n), -2 - i)
        return np.array([-1 if n.parent is None else ix.get(id(n.parent), -9) for n in nodes], np.int64)

    # ---- (A) RRT*, shipped behaviour (rewire is a no-op); (B) linear rewire on
    for tag, linear_rewire in (("rrt", False), ("rrtrw", True)):
        rec = dict(rand=[], nn=[], accept=[])

        def hook(env, rec=rec, linear_rewire=linear_rewire):
            pl = env.planner
            smp = env.args.sampler
            inner_next = smp.get_valid_next_pos
            inner_nn = pl.find_nearest_neighbour_idx

            def get_valid_next_pos():
                r = inner_next()
                rec["rand"].append(np.array(r[0], dtype=np.float64))
                rec["accept"].append(False)
                return r

            def find_nn(pos, poses):
                i = inner_nn(pos, poses)
                rec["nn"].append(int(i))
                return i

            inner_add = pl.add_newnode

            def add_newnode(node):
                if rec["accept"]:
                    rec["accept"][-1] = True
                return inner_add(node)

            smp.get_valid_next_pos = get_valid_next_pos
            pl.find_nearest_neighbour_idx = find_nn
            pl.add_newnode = add_newnode
            if linear_rewire:
                pl.rewire = lambda newnode, nodes, **kw: RRTPlanner.rewire(pl, newnode, nodes, use_rtree=False)

        r = _refrun.run("rrt", fac, None, "100,100", "350,350", 500, before_run=hook)
        pl = r["env"].planner
        T = len(rec["rand"])
        assert len(rec["nn"]) == T
        out[f"{tag}_rand"] = np.array(rec["rand"])
        out[f"{tag}_nn"] = np.array(rec["nn"], np.int64)
        out[f"{tag}_accept"] = np.packbits(np.array(rec["accept"], bool))
        out[f"{tag}_steps"] = np.array(T)
        out[f"{tag}_poses"] = r["poses"]
        out[f"{tag}_parent"] = parents_of(pl.nodes)
        out[f"{tag}_cost"] = np.array([float(n.cost) for n in pl.nodes])
        out[f"{tag}_c_max"] = np.array(float(pl.c_max))
        out[f"{tag}_counters"] = np.array(r["counters"], np.int64)
        out[f"{tag}_params"] = np.array([pl.args.epsilon, pl.args.radius, pl.args.goal_radius], np.float64)
        out[f"{tag}_goal"] = np.array(pl.goal_pt.pos, np.float64)
        print(tag, "steps", T, "nodes", r["n"], "c_max", pl.c_max, "counters", r["counters"])

    # ---- (C) Bi-RRT join test per accepted node, up to and including the join
    rec = dict(tree=[], newpos=[], join=[])

    def hook_bi(env):
        pl = env.planner
        cc = pl.args.engine.cc
        inner_vis = cc.visible
        state = dict(last_two=[])

        inner_run = pl.run_once

        def run_once():
            was_found = pl.found_solution
            n0, n1 = len(pl.nodes), len(pl.goal_tree_nodes)
            turn_goal = pl.goal_tree_turn and not pl.found_solution
            inner_run()
            if was_found:
                return
            grew = len(pl.goal_tree_nodes) > n1 if turn_goal else len(pl.nodes) > n0
            if not grew:
                return
            own_nodes = pl.goal_tree_nodes if turn_goal else pl.nodes
            own_n0 = n1 if turn_goal else n0
            newpos = np.array(own_nodes[own_n0].pos, dtype=np.float64)
            other_poses = pl.poses[:n0] if turn_goal else pl.goal_tree_poses[:n1]
            d = np.linalg.norm(other_poses - newpos, axis=1)            # birrtPlanner.py:82-84
            j = -1
            if min(d) < pl.args.epsilon:
                i = int(np.argmin(d))
                if pl.found_solution:                                    # the planner joined here
                    j = i
            rec["tree"].append(1 if turn_goal else 0)
            rec["newpos"].append(newpos)
            rec["join"].append(j)

        pl.run_once = run_once

    r = _refrun.run("birrt", fac, None, "100,100", "350,350", 500, before_run=hook_bi)
    pl = r["env"].planner
    out["bi_tree"] = np.array(rec["tree"], np.int64)
    out["bi_newpos"] = np.array(rec["newpos"])
    out["bi_join"] = np.array(rec["join"], np.int64)
    out["bi_start"], out["bi_goal"] = np.array(pl.start_pt.pos, np.float64), np.array(pl.goal_pt.pos, np.float64)
    out["bi_eps"] = np.array(float(pl.args.epsilon))
    out["bi_found"] = np.array(bool(pl.found_solution))
    print("birrt accepted before/at join", len(rec["tree"]), "joined at", int(np.flatnonzero(np.array(rec["join"]) >= 0)[0])
          if (np.array(rec["join"]) >= 0).any() else None, "found", pl.found_solution)

    # ---- (D) RRdT*: neighbour search over a snapshot of the disjoint trees
    r = _refrun.run("rrdt", fac, None, "100,100", "350,350", 400)
    pl = r["env"].planner
    trees = [*pl._disjointed_trees, pl.root]
    sizes = [len(t.nodes) for t in trees]
    out["rrdt_tree_off"] = np.cumsum([0] + sizes).astype(np.int64)
    out["rrdt_tree_poses"] = np.concatenate([np.array(t.poses[: len(t.nodes)]) for t in trees])
    rng = np.random.default_rng(5)
    W, H = pl.args.engine.cc._img.shape
    Q = rng.random((300, 2)) * [W, H]
    Q[:100] = out["rrdt_tree_poses"][rng.integers(0, len(out["rrdt_tree_poses"]), 100)] + rng.standard_normal((100, 2)) * 4
    own = rng.integers(-1, len(trees), len(Q))
    res_off, res = [0], []
    for q, o in zip(Q, own):
        lst = pl.find_nearest_node_from_neighbour(Node(q), None if o < 0 else trees[o], pl.args.epsilon)
        for nd, tr in lst:
            ti = next(i for i, t in enumerate(trees) if t is tr)
            res.append((ti, next(i for i, n in enumerate(tr.nodes) if n is nd)))
        res_off.append(len(res))
    out["rrdt_Q"], out["rrdt_own"] = Q, own.astype(np.int64)
    out["rrdt_res_off"] = np.array(res_off, np.int64)
    out["rrdt_res"] = np.array(res, np.int64).reshape(-1, 2)
    out["rrdt_eps"] = np.array(float(pl.args.epsilon))
    print("rrdt trees", sizes, "hits", len(res))

    # ---- (E) PRM: build_graph edge list (insertion order) + get_solution
    r = _refrun.run("prm", fac, None, "100,100", "350,350", 700)
    pl = r["env"].planner
    pl.build_graph()
    ix = index_map(pl.nodes)
    e = [(ix[id(u)], ix[id(v)], w["weight"]) for u, v, w in pl.graph.edges(data=True) if id(u) in ix and id(v) in ix]
    n = len(pl.nodes)
    out["prm_poses"] = r["poses"]
    out["prm_radius"] = np.array(float(pl.gamma * np.power(np.log(n + 1) / (n + 1), 1 / 2)))
    out["prm_edges"] = np.array([(a, b) for a, b, _ in e], np.int64)
    out["prm_weights"] = np.array([w for _, _, w in e])
    sol = pl.get_solution()
    import networkx as nx

    goal = pl.args.env.goal_pt.parent
    start = next((nd for nd in pl.nodes if nd.parent is pl.args.env.start_pt), None)
    out["prm_c_max"] = np.array(float(sol))
    out["prm_start_goal"] = np.array([-1 if start is None else ix[id(start)], -1 if goal is None else ix[id(goal)]], np.int64)
    out["prm_hops"] = np.array(-1 if start is None or goal is None else nx.shortest_path_length(pl.graph, start, goal))
    out["prm_query"] = np.array([pl.args.sampler.start_pos, pl.args.sampler.goal_pos], np.float64)
    out["prm_eps"] = np.array(float(pl.args.epsilon))
    # hop counts of a batch of node pairs (nx.shortest_path_length, -1 = no path)
    rng = np.random.default_rng(6)
    pairs = rng.integers(0, n, (200, 2))
    hops = []
    for a, b in pairs:
        try:
            hops.append(nx.shortest_path_length(pl.graph, pl.nodes[a], pl.nodes[b]))
        except (nx.NetworkXNoPath, nx.NodeNotFound):      # an isolated node was never added
            hops.append(-1)
    out["prm_pairs"], out["prm_pair_hops"] = pairs.astype(np.int64), np.array(hops, np.int64)
    print("prm nodes", n, "edges", len(e), "c_max", sol, "hops", int(out["prm_hops"]))
    np.savez_compressed(os.path.join(GOLDEN, "planner.npz"), **out)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = ref_shims.import_reference()
    os.chdir(ref_shims.REFERENCE_ROOT)
    gen_maps(ref)
    gen_img_cc(ref)
    gen_arm_cc(ref)
    gen_nn(ref)
    gen_planner(ref)
    for f in sorted(os.listdir(GOLDEN)):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)))


if __name__ == "__main__":
    main()
