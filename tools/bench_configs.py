"""BASELINE.json configs 3-5 at full size on one B200 (run under gpurun): throughput of
the kernels in the large-map / large-tree regimes, with parity checked against the oracle
on subsamples and through size-independent properties.  Not a bench line (bench.py is
configs[1]); the JSON it prints is committed under profiles/ as evidence.

    python tools/bench_configs.py [cfg3] [cfg4] [cfg5]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import c_oracle as orc  # noqa: E402
import sbp_env_b200 as sg  # noqa: E402
from _util import load_map  # noqa: E402

dev = torch.device("cuda", 0)
HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)) / 1e3


def edge_stats(P, Q, t, n_px_tested=None):
    px = (torch.maximum((P[:, 0].trunc() - Q[:, 0].trunc()).abs(), (P[:, 1].trunc() - Q[:, 1].trunc()).abs()) + 1).sum().item()
    n = P.shape[0]
    ab = 33.0 * n + 0.125 * px
    return {"edges": n, "ms": t * 1e3, "edges_per_s": n / t, "interpolants_full_line": px,
            "interpolants_per_s": px / t, "algorithmic_GBps": ab / t / 1e9, "hbm_frac": ab / t / 1e9 / HBM}


def free_samples(free_dev, n, seed):
    """n uniform samples in free space (device), first accepted of a seeded stream."""
    W, H = free_dev.shape
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    out, have = [], 0
    while have < n:
        c = torch.rand((2 * n, 2), generator=g, device=dev, dtype=torch.float64) * torch.tensor([W, H], device=dev)
        ok = free_dev[c[:, 0].long(), c[:, 1].long()] != 0
        c = c[ok]
        out.append(c)
        have += c.shape[0]
    return torch.cat(out)[:n].contiguous()


def cfg1():
    """RRT*-style growth on maps/room1, start (100,100), 5000 accepted nodes, epsilon 10,
    radius 11: the per-node GPU path (nearest + visible + fused near_visible through
    planner_ops) against the reference-style O(N) Python loop on the host (oracle checker)."""
    import math

    from sbp_env_b200 import planner_ops as ops

    free = load_map("room1")
    W, H = free.shape
    cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8))
    om = orc.Map(free)

    def ref_dist(a, b):
        p = a - b
        return math.sqrt(p[0] ** 2 + p[1] ** 2)

    def grow(n_nodes, gpu):
        rng = np.random.default_rng(0)
        poses = np.empty((n_nodes + 1, 2))
        poses[0] = (100.0, 100.0)
        costs, parents, n = [0.0], [-1], 1
        tree = sg.NodeStore(2, capacity=n_nodes + 1) if gpu else None
        if gpu:
            tree.append(poses[0])
        calls = 0
        while n < n_nodes:
            r = rng.random(2) * [W, H]
            if gpu:
                if not cc.feasible(r):
                    continue
                nn = tree.nearest(r)
            else:
                if not om.feasible2d(r[None])[0]:
                    continue
                nn = int(np.argmin(np.linalg.norm(poses[:n] - r, axis=1)))
            v = r - poses[nn]
            newpos = poses[nn] + min(ref_dist(poses[nn], r), 10.0) * v / np.linalg.norm(v)
            if gpu:
                if not cc.visible(poses[nn], newpos):
                    continue
                parent, cost, cand = ops.choose_least_cost_parent(cc, tree, costs, newpos, nn, 11.0, poses=poses)
                tree.append(newpos)
                calls += len(cand[0]) + 1
            else:
                if not om.visible2d(poses[nn][None], newpos[None])[0]:
                    continue
                parent, c_nn = nn, ref_dist(newpos, poses[nn])
                for p in range(n):                       # rrtPlanner.py:176-187
                    c = ref_dist(newpos, poses[p])
                    if c <= 11.0 and om.visible2d(newpos[None], poses[p][None])[0]:
                        calls += 1
                        if costs[p] + c < costs[parent] + c_nn:
                            parent, c_nn = p, c
                cost = costs[parent] + ref_dist(poses[parent], newpos)
            poses[n] = newpos
            costs.append(cost)
            parents.append(parent)
            n += 1
        return poses[:n], costs, parents, calls

    grow(200, True)
    t0 = time.perf_counter()
    gp, gc, gpar, gcalls = grow(5000, True)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    hp, hc, hpar, _ = grow(1500, False)
    t_host = time.perf_counter() - t0
    return {"gpu_5000_nodes_s": t_gpu, "gpu_us_per_node": t_gpu / 5000 * 1e6, "visible_checks": gcalls,
            "host_python_loop_1500_nodes_s": t_host,
            "identical_first_1500_nodes": bool(np.array_equal(gp[:1500], hp) and gc[:1500] == hc and gpar[:1500] == hpar),
            "note": "host loop = the reference's O(N) per-node Python loop restated with the C oracle as checker"}


def cfg3():
    """4-D arm on maps/4d (1024^2), L = (30, 30): batched link-sweep checks."""
    free = load_map("4d")
    cc = sg.RobotArm4dCollisionChecker.__new__(sg.RobotArm4dCollisionChecker)
    sg.CollisionChecker.__init__(cc)
    cc.image_fname, cc.stick_robot_length_config = None, [30.0, 30.0]
    cc._attach(free.astype(np.uint8), 0)
    om = orc.Map(free)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    lo = torch.tensor([0, 0, -np.pi, -np.pi], device=dev, dtype=torch.float64)
    hi = torch.tensor([1024, 1024, np.pi, np.pi], device=dev, dtype=torch.float64)
    out = {}
    n = 1 << 20
    C1 = torch.rand((n, 4), generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo
    step = torch.randn((n, 4), generator=g, device=dev, dtype=torch.float64)
    step = step / step[:, :2].norm(dim=1, keepdim=True) * 10.0          # epsilon = 10 px base step
    step[:, 2:] = step[:, 2:].clamp(-1, 1) * 0.5
    sets = {"rrt_like_10px": (C1, C1 + step),
            "uniform_pairs": (C1[: n // 8], torch.rand((n // 8, 4), generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo)}
    for name, (A, B) in sets.items():
        cnt = torch.zeros(1, dtype=torch.int64, device=dev)
        raw = cc._visible_device(A, B, cfg_tested=cnt, resolve=False)
        t = timeit(lambda: cc._visible_device(A, B, resolve=False))
        ncfg = torch.clamp(torch.maximum((A[:, 0].trunc() - B[:, 0].trunc()).abs(), (A[:, 1].trunc() - B[:, 1].trunc()).abs()) + 1, min=2).sum().item()
        sub = slice(0, 20000)
        exp = om.arm_visible(A[sub].cpu().numpy(), B[sub].cpu().numpy(), 30.0, 30.0)
        got = cc.visible_batch(A[sub], B[sub]).cpu().numpy()
        out[name] = {"edges": A.shape[0], "ms": t * 1e3, "edges_per_s": A.shape[0] / t,
                     "interpolated_configs_full": ncfg, "configs_per_s": ncfg / t,
                     "configs_evaluated": int(cnt.item()), "ambiguous": int((raw == 2).sum().item()),
                     "visible_fraction": float((raw == 1).float().mean().item()),
                     "algorithmic_GBps": 65.0 * A.shape[0] / t / 1e9,
                     "parity_vs_oracle_20000": bool(np.array_equal(got, exp))}
    Cf = C1
    t = timeit(lambda: cc.feasible_batch(Cf.cpu().numpy()[:1]))  # warm
    Cn = Cf.cpu().numpy()
    t0 = time.perf_counter()
    f = cc.feasible_batch(Cn)
    out["feasible_host_api_2^20"] = {"configs_per_s_e2e": n / (time.perf_counter() - t0),
                                     "parity_vs_oracle_20000": bool(np.array_equal(f[:20000], om.arm_feasible(Cn[:20000], 30.0, 30.0)))}
    return out


def cfg4():
    """16384^2 blocky-noise grid, 1M-node tree, near(r=64) + visible for a stream of new samples."""
    g = torch.Generator(device=dev)
    g.manual_seed(16384)
    coarse = (torch.rand((256, 256), generator=g, device=dev) < 0.75).to(torch.uint8)
    free = coarse.repeat_interleave(64, 0).repeat_interleave(64, 1).contiguous()
    W = H = 16384
    cc = sg.ImgCollisionChecker.from_free_mask(free)
    info = cc.device_map.info()
    nodes = free_samples(free, 1_000_000, 1)
    tree = sg.NodeStore(2, capacity=1 << 20)
    tree.extend(nodes)
    X = free_samples(free, 10_000, 7)
    out = {"map": {"W": W, "H": H, **info}}
    # batched: one radius query for the whole stream, then one visible launch
    rp, idx = tree.near(X, 64.0, "xy", True)
    t_near = timeit(lambda: tree.near(X, 64.0, "xy", True))
    rows = torch.repeat_interleave(torch.arange(X.shape[0], device=dev), rp[1:] - rp[:-1])
    P, Q = X[rows].contiguous(), nodes[idx].contiguous()
    vis = cc.visible_batch(P, Q)
    t_vis = timeit(lambda: cc.visible_batch(P, Q))
    out["near_batched_10000x1M_r64"] = {"ms": t_near * 1e3, "neighbours": int(idx.numel()),
                                        "dist_evals_per_s": 1e4 * 1e6 / t_near}
    out["visible_of_neighbours"] = {**edge_stats(P, Q, t_vis), "visible_fraction": float(vis.float().mean().item())}
    # parity on a subsample against the oracle (needs the 256 MiB byte grid on the host)
    fh = free.cpu().numpy()
    om = orc.Map(fh)
    sel = torch.randperm(P.shape[0], generator=g, device=dev)[:100_000]
    exp = om.visible2d(P[sel].cpu().numpy(), Q[sel].cpu().numpy())
    out["visible_of_neighbours"]["parity_vs_oracle_100000"] = bool(np.array_equal(vis[sel].cpu().numpy(), exp))
    nh = nodes.cpu().numpy()
    orp, oidx = orc.near(nh, X[:64].cpu().numpy(), 64.0, "xy", True)
    out["near_batched_10000x1M_r64"]["parity_vs_oracle_64_queries"] = bool(
        np.array_equal(rp[:65].cpu().numpy(), orp) and np.array_equal(idx[: int(orp[-1])].cpu().numpy(), oidx))
    # sequential: the fused one-launch query per new sample (RRT* insertion latency)
    Xh = X.cpu().numpy()
    tree.near_visible(cc, Xh[0], 64.0)
    t0 = time.perf_counter()
    nh_hits = 0
    for i in range(2000):
        ii, vv = tree.near_visible(cc, Xh[i], 64.0)
        nh_hits += len(ii)
    dt = time.perf_counter() - t0
    out["fused_near_visible_per_sample"] = {"us_per_sample": dt / 2000 * 1e6, "mean_neighbours": nh_hits / 2000,
                                            "samples_per_s": 2000 / dt}
    e0 = int(rp[0].item())
    ii, vv = tree.near_visible(cc, Xh[0], 64.0)
    out["fused_near_visible_per_sample"]["parity_vs_batched"] = bool(
        np.array_equal(ii, idx[e0:int(rp[1].item())].cpu().numpy()) and np.array_equal(vv, vis[e0:int(rp[1].item())].cpu().numpy()))
    # single nearest on the 1M tree
    q1 = X[:1].contiguous()
    t1 = timeit(lambda: tree.knn(q1, 1), reps=20)
    out["nearest_1q_1M"] = {"us": t1 * 1e6, "dist_evals_per_s": 1e6 / t1}
    return out


def make_maze(cells=256, pitch=128, wall=16, seed=32768):
    """Seeded recursive-backtracker maze; returns the free grid [W][H] uint8 on the device."""
    rng = np.random.default_rng(seed)
    east = np.zeros((cells, cells), bool)      # passage between (cx, cy) and (cx+1, cy)
    south = np.zeros((cells, cells), bool)     # passage between (cx, cy) and (cx, cy+1)
    seen = np.zeros((cells, cells), bool)
    stack = [(0, 0)]
    seen[0, 0] = True
    while stack:
        cx, cy = stack[-1]
        nb = [(cx + dx, cy + dy) for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1))
              if 0 <= cx + dx < cells and 0 <= cy + dy < cells and not seen[cx + dx, cy + dy]]
        if not nb:
            stack.pop()
            continue
        nx, ny = nb[rng.integers(len(nb))]
        if nx != cx:
            east[min(cx, nx), cy] = True
        else:
            south[cx, min(cy, ny)] = True
        seen[nx, ny] = True
        stack.append((nx, ny))
    W = cells * pitch
    h = wall // 2
    free = torch.empty((W, W), dtype=torch.uint8, device=dev)
    east_d, south_d = torch.from_numpy(east).to(dev), torch.from_numpy(south).to(dev)
    ys = torch.arange(W, device=dev)
    ly, cy = ys % pitch, ys // pitch
    in_y = (ly >= h) & (ly < pitch - h)
    for x0 in range(0, W, 1024):
        xs = torch.arange(x0, x0 + 1024, device=dev)
        lx, cx = xs % pitch, xs // pitch
        in_x = (lx >= h) & (lx < pitch - h)
        f = in_x[:, None] & in_y[None, :]
        e_here = east_d[cx][:, cy]
        e_left = east_d[(cx - 1).clamp(min=0)][:, cy] & (cx > 0)[:, None]
        f |= (lx >= pitch - h)[:, None] & in_y[None, :] & e_here
        f |= (lx < h)[:, None] & in_y[None, :] & e_left
        s_here = south_d[cx][:, cy]
        s_up = south_d[cx][:, (cy - 1).clamp(min=0)] & (cy > 0)[None, :]
        f |= in_x[:, None] & (ly >= pitch - h)[None, :] & s_here
        f |= in_x[:, None] & (ly < h)[None, :] & s_up
        free[x0:x0 + 1024] = f.to(torch.uint8)
    return free


def cfg5():
    """32768^2 maze, 1M free nodes, k = 10 -> 10M candidate edges; 4096 start/goal pairs."""
    t0 = time.perf_counter()
    free = make_maze()
    W = H = free.shape[0]
    cc = sg.ImgCollisionChecker.from_free_mask(free)
    torch.cuda.synchronize()
    out = {"map": {"W": W, "H": H, **cc.device_map.info(), "free_fraction": float(free.float().mean().item()),
                   "build_s": time.perf_counter() - t0}}
    nodes = free_samples(free, 1_000_000, 2)
    tree = sg.NodeStore(2, capacity=1 << 20)
    tree.extend(nodes)
    k = 10
    ev = {}

    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev[name] = e

    s = torch.cuda.Event(enable_timing=True)
    sg.prm_connect_knn(cc, tree, k)                    # warm-up
    torch.cuda.synchronize()
    flush.zero_()
    s.record()
    nbr, vis = sg.prm_connect_knn(cc, tree, k, mark=mark)
    torch.cuda.synchronize()
    t_knn = s.elapsed_time(ev["knn"]) / 1e3
    t_edges = ev["knn"].elapsed_time(ev["edges"]) / 1e3
    t_vis = ev["edges"].elapsed_time(ev["visible"]) / 1e3
    P = nodes[nbr.reshape(-1)].contiguous()
    Q = nodes.repeat_interleave(k, 0).contiguous()
    out["prm_1M_nodes_k10"] = {"build_ms": (t_knn + t_edges + t_vis) * 1e3, "knn_ms": t_knn * 1e3,
                               "knn_dist_evals_per_s": 1e12 / t_knn, "edges_ms": t_edges * 1e3,
                               "visible": {**edge_stats(P, Q, t_vis), "visible_fraction": float(vis.float().mean().item())}}
    # properties at full size: direction symmetry, idempotence, visible => endpoints feasible
    v2 = cc.visible_batch(Q, P)
    fe = cc.feasible_batch(P) & cc.feasible_batch(Q)
    out["prm_1M_nodes_k10"]["properties"] = {
        "direction_symmetric": bool(torch.equal(v2.reshape(-1, k) & (nbr >= 0), vis)),
        "visible_implies_feasible_endpoints": bool((~(vis.reshape(-1) & ~fe)).all().item())}
    # oracle parity on subsamples (1 GiB byte grid on the host)
    fh = free.cpu().numpy()
    om = orc.Map(fh)
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    sel = torch.randperm(P.shape[0], generator=g, device=dev)[:200_000]
    exp = om.visible2d(P[sel].cpu().numpy(), Q[sel].cpu().numpy())
    out["prm_1M_nodes_k10"]["visible"]["parity_vs_oracle_200000"] = bool(np.array_equal(vis.reshape(-1)[sel].cpu().numpy(), exp))
    nh = nodes.cpu().numpy()
    rows = np.arange(0, 1_000_000, 1_000_000 // 256)[:256]
    oi, _ = orc.knn(nh, nh[rows], k + 1)
    ok = True
    nb = nbr.cpu().numpy()
    for r, row in zip(rows, oi):
        lst = row.tolist()
        lst.remove(r) if r in lst else lst.pop()
        ok &= nb[r].tolist() == lst
    out["prm_1M_nodes_k10"]["knn_parity_vs_oracle_256_rows"] = bool(ok)
    # uniform random long edges on the big map (L2 / HBM regime)
    n_u = 1 << 22
    Pu = torch.rand((n_u, 2), generator=g, device=dev, dtype=torch.float64) * W
    Qu = Pu + (torch.rand((n_u, 2), generator=g, device=dev, dtype=torch.float64) - 0.5) * 600
    vu = cc.visible_batch(Pu, Qu)
    out["uniform_600px_box_edges_2^22"] = {**edge_stats(Pu, Qu, timeit(lambda: cc.visible_batch(Pu, Qu))),
                                           "visible_fraction": float(vu.float().mean().item())}
    # 4096 independent start/goal pairs: connect each endpoint to its k nearest roadmap nodes
    SG = free_samples(free, 8192, 3)
    tq = timeit(lambda: sg.prm_connect_knn(cc, tree, k, queries=SG))
    out["connect_4096_start_goal_pairs"] = {"ms": tq * 1e3, "queries": 8192,
                                            "dist_evals_per_s": 8192 * 1e6 / tq}
    # roadmap graph on the device: CSR assembly of the 10M candidate edges and
    # the 4096 start/goal queries as fewest-hop searches (one CTA per query)
    gm = sg.Roadmap(len(tree))
    t_build = timeit(lambda: gm.build_from_knn(nbr, vis))
    nq, kq = 4096, 1
    nb_q, vis_q = sg.prm_connect_knn(cc, tree, kq, queries=SG)        # nearest roadmap node of each endpoint
    ends = torch.where(vis_q[:, 0], nb_q[:, 0], torch.full_like(nb_q[:, 0], -1))
    S, T = ends[:nq].contiguous(), ends[nq:].contiguous()
    t_bfs = timeit(lambda: gm.shortest_paths(S, T, max_len=2048), reps=2)
    hops, paths = gm.shortest_paths(S, T, max_len=2048)
    hh = hops.cpu().numpy()
    # spot check against networkx on the same edge list (a few queries: nx BFS on 8.7M edges is slow)
    import networkx as nx

    rp, col = gm.csr()
    rp, col = rp.cpu().numpy(), col.cpu().numpy()
    G = nx.DiGraph()
    G.add_nodes_from(range(len(tree)))
    G.add_edges_from(zip(np.repeat(np.arange(len(tree)), np.diff(rp)).tolist(), col.tolist()))
    ok = True
    Sh, Th = S.cpu().numpy(), T.cpu().numpy()
    for q in range(8):
        if Sh[q] < 0 or Th[q] < 0:
            ok &= hh[q] == -1
        elif nx.has_path(G, int(Sh[q]), int(Th[q])):
            ok &= hh[q] == nx.shortest_path_length(G, int(Sh[q]), int(Th[q]))
        else:
            ok &= hh[q] == -1
    out["roadmap_graph"] = {"nodes": len(tree), "edges": gm.n_edges, "csr_build_ms": t_build * 1e3,
                            "bfs_4096_queries_ms": t_bfs * 1e3, "queries_per_s": nq / t_bfs,
                            "reachable_fraction": float((hh >= 0).mean()),
                            "mean_hops_reachable": float(hh[hh >= 0].mean()) if (hh >= 0).any() else None,
                            "max_hops": int(hh.max()), "hops_parity_vs_networkx_8_queries": bool(ok)}
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
    res = {}
    for w in which:
        t0 = time.perf_counter()
        res[w] = {"cfg1": cfg1, "cfg3": cfg3, "cfg4": cfg4, "cfg5": cfg5}[w]()
        res[w]["wall_s"] = time.perf_counter() - t0
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))
