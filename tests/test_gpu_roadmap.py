"""-m gpu: the device roadmap against networkx, the library the reference's
PRMPlanner uses (prmPlanner.py:99-101 DiGraph edges m_g -> v; :139-143 has_path /
shortest_path, unweighted)."""
import networkx as nx
import numpy as np
import pytest

import c_oracle as orc
from _util import load_map, map_png

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


def _levels_and_min_parent_path(adj_in, adj_out, n, s, t):
    """Oracle of the deterministic rule: BFS levels from s; walking back from t every node takes
    its lowest-index predecessor on the previous level."""
    level = np.full(n, -1, np.int64)
    level[s] = 0
    frontier = [s]
    while frontier and level[t] < 0:
        nxt = []
        for u in frontier:
            for v in adj_out[u]:
                if level[v] < 0:
                    level[v] = level[u] + 1
                    nxt.append(v)
        frontier = nxt
    if level[t] < 0:
        return -1, None
    path = [t]
    while path[-1] != s:
        v = path[-1]
        path.append(min(u for u in adj_in[v] if level[u] == level[v] - 1))
    return int(level[t]), path[::-1]


def _build(sg, name, n, k, seed):
    free = load_map(name)
    cc = sg.ImgCollisionChecker(map_png(name))
    rng = np.random.default_rng(seed)
    pts = rng.random((4 * n, 2)) * free.shape
    pts = pts[cc.feasible_batch(pts)][:n]
    tree = sg.NodeStore(2)
    tree.extend(pts)
    nbr, vis = sg.prm_connect_knn(cc, tree, k)
    return cc, tree, pts, nbr, vis


@pytest.mark.parametrize("name,n,k", [("maze1", 4000, 12), ("intel_lab", 6000, 10)])
def test_roadmap_matches_networkx(sg, name, n, k):
    cc, tree, pts, nbr, vis = _build(sg, name, n, k, 3)
    g = sg.Roadmap.from_knn(nbr, vis)
    nb, vs = nbr.cpu().numpy(), vis.cpu().numpy()
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    for v in range(n):                               # prmPlanner.py:91-101
        for j in range(k):
            if vs[v, j] and nb[v, j] >= 0:
                G.add_edge(int(nb[v, j]), v)
    want = np.array(sorted(G.edges()), np.int64).reshape(-1, 2)
    assert g.n_edges == len(want)
    assert np.array_equal(g.edges(), want)

    adj_out = [sorted(G.successors(u)) for u in range(n)]
    adj_in = [sorted(G.predecessors(u)) for u in range(n)]
    rng = np.random.default_rng(0)
    S, T = rng.integers(0, n, 400), rng.integers(0, n, 400)
    S[:5] = T[:5]                                    # start == goal
    hops, paths = g.shortest_paths(S, T, max_len=n)
    reach = 0
    for q in range(len(S)):
        s, t = int(S[q]), int(T[q])
        if nx.has_path(G, s, t):
            reach += 1
            assert hops[q] == nx.shortest_path_length(G, s, t), (q, s, t)
            h, p = _levels_and_min_parent_path(adj_in, adj_out, n, s, t)
            got = paths[q, : hops[q] + 1]
            assert h == hops[q] and list(got) == p, (q, s, t)
            assert all(G.has_edge(int(a), int(b)) for a, b in zip(got[:-1], got[1:]))
            assert (paths[q, hops[q] + 1:] == -1).all()
        else:
            assert hops[q] == -1 and (paths[q] == -1).all()
    assert reach > 20
    # hop counts alone, device tensors in and out, and a path buffer that is too short
    import torch

    h2 = g.shortest_paths(torch.from_numpy(S).cuda(), torch.from_numpy(T).cuda())
    assert np.array_equal(h2.cpu().numpy(), hops)
    h3, p3 = g.shortest_paths(S, T, max_len=3)
    assert np.array_equal(h3, hops)
    assert ((p3 == -1).all(1) == ((hops < 0) | (hops + 1 > 3))).all()
    assert np.array_equal(g.has_path(S, T), hops >= 0)
    # out-of-range indices are "no path", never a fault
    hb = g.shortest_paths(np.array([-1, 0, n]), np.array([0, n + 5, 1]))
    assert (hb == -1).all()


def test_roadmap_from_edge_lists_and_rebuild(sg):
    n = 50
    src = np.array([0, 1, 2, 3, 3, 10, 0, 7, 7], np.int64)
    dst = np.array([1, 2, 3, 4, 0, 11, 2, 7, 99], np.int64)         # 7->7 self loop, 7->99 out of range
    g = sg.Roadmap(n)
    g.build_from_edges(src, dst)
    assert g.n_edges == 8
    assert np.array_equal(g.edges(), np.array(sorted(zip(src[:-1], dst[:-1])), np.int64))
    hops, paths = g.shortest_paths(np.array([0, 0, 4, 10, 3]), np.array([4, 3, 0, 11, 2]), max_len=8)
    assert hops.tolist() == [3, 2, -1, 1, 2]
    assert paths[0, :4].tolist() == [0, 2, 3, 4] and paths[1, :3].tolist() == [0, 2, 3]
    assert paths[4, :4].tolist() == [3, 0, 2, -1]
    keep = np.ones(len(src), np.uint8)
    keep[6] = 0                                                     # drop the shortcut 0 -> 2
    g.build_from_edges(src, dst, keep)
    assert g.n_edges == 7
    hops, paths = g.shortest_paths(np.array([0]), np.array([4]), max_len=8)
    assert hops.tolist() == [4] and paths[0, :5].tolist() == [0, 1, 2, 3, 4]
    g.build_from_edges(np.zeros(0, np.int64), np.zeros(0, np.int64))
    assert g.n_edges == 0 and g.shortest_paths(np.array([0, 5]), np.array([1, 5])).tolist() == [-1, 0]


def test_prm_get_solution_on_device_roadmap(sg):
    """get_solution (prmPlanner.py:128-154): nearest visible roadmap node within epsilon for
    start and goal, fewest-hop path, cost accumulated with engine.dist."""
    cc, tree, pts, nbr, vis = _build(sg, "maze1", 4000, 12, 9)
    g = sg.Roadmap.from_knn(nbr, vis)
    om = orc.Map(load_map("maze1"))
    # start next to node 0, goal next to the node farthest (in hops) from it
    n = len(pts)
    reach = g.shortest_paths(np.zeros(n, np.int64), np.arange(n))
    far = int(np.argmax(reach))
    assert reach[far] >= 5
    start, goal, eps = pts[0] + [0.25, -0.25], pts[far] + [-0.25, 0.25], 9.0

    def nearest_free(pos):                           # prmPlanner.py:103-126 with the oracle
        d = np.linalg.norm(pts - pos, axis=1)
        best, bi = None, None
        for i in np.flatnonzero(d < eps):
            if om.visible2d(pos.reshape(1, 2), pts[i].reshape(1, 2))[0]:
                c = sg.euclidean_dist(pos, pts[i])
                if best is None or c < best:
                    best, bi = c, int(i)
        return bi

    s, t = nearest_free(start), nearest_free(goal)
    c_max, path = sg.prm_get_solution(cc, tree, g, start, goal, eps)
    assert s is not None and t is not None and path is not None
    assert path[0] == s and path[-1] == t
    want = 0.0
    for a, b in zip(path[:-1], path[1:]):
        want = want + sg.euclidean_dist(pts[b], pts[a])
    assert c_max == want
    hops = g.shortest_paths(np.array([s]), np.array([t]))
    assert len(path) == hops[0] + 1
    # a goal no roadmap node is within epsilon of
    assert sg.prm_get_solution(cc, tree, g, start, np.array([-500.0, -500.0]), eps) == (float("inf"), None)


@pytest.mark.parametrize("name,n,k", [("maze1", 5000, 6), ("room1", 4000, 8)])
def test_undirected_roadmap_matches_networkx(sg, name, n, k):
    """The undirected form -- both directions of every visible candidate, what the reference's
    symmetric radius rule puts into its DiGraph (prmPlanner.py:91-101) -- against networkx on 512
    queries: the arc set (no arc twice), reachability, hop counts, component labels; the packed
    adjacency builds the identical graph."""
    import torch

    cc, tree, pts, nbr, vis = _build(sg, name, n, k, 11)
    g = sg.Roadmap.from_knn(nbr, vis, undirected=True)
    nb, vs = nbr.cpu().numpy(), vis.cpu().numpy()
    G = nx.DiGraph()
    G.add_nodes_from(range(n))
    for v in range(n):
        for j in range(k):
            if vs[v, j] and nb[v, j] >= 0:
                G.add_edge(int(nb[v, j]), v)
                G.add_edge(v, int(nb[v, j]))
    want = np.array(sorted(G.edges()), np.int64).reshape(-1, 2)
    assert g.n_edges == len(want)                                 # mirrored arcs are not stored twice
    assert np.array_equal(g.edges(), want)
    gp = sg.Roadmap.from_packed(sg.distributed.pack_adjacency(nbr, vis), undirected=True)
    assert np.array_equal(gp.edges(), want)
    gd = sg.Roadmap.from_packed(sg.distributed.pack_adjacency(nbr, vis))
    assert np.array_equal(gd.edges(), sg.Roadmap.from_knn(nbr, vis).edges())
    # components: label = smallest node index of the component
    labels, n_comp = g.components()
    comps = list(nx.weakly_connected_components(G))
    assert n_comp == len(comps)
    want_label = np.empty(n, np.int64)
    for c in comps:
        want_label[list(c)] = min(c)
    assert np.array_equal(labels.cpu().numpy(), want_label)
    rng = np.random.default_rng(1)
    S, T = rng.integers(0, n, 512), rng.integers(0, n, 512)
    S[:4] = T[:4]
    hops = g.shortest_paths(S, T)
    hops2, paths = g.shortest_paths(torch.from_numpy(S).cuda(), torch.from_numpy(T).cuda(), max_len=n)
    assert np.array_equal(hops2.cpu().numpy(), hops)
    paths = paths.cpu().numpy()
    reach = 0
    for q in range(len(S)):
        s, t = int(S[q]), int(T[q])
        if nx.has_path(G, s, t):
            reach += 1
            assert hops[q] == nx.shortest_path_length(G, s, t), (q, s, t)
            got = paths[q, : hops[q] + 1]
            assert got[0] == s and got[-1] == t
            assert all(G.has_edge(int(a), int(b)) for a, b in zip(got[:-1], got[1:]))
        else:
            assert hops[q] == -1 and (paths[q] == -1).all()
    assert 20 < reach
    # the weak-component filter on a DIRECTED graph is only a necessary condition: pairs inside one
    # weak component without a directed path must still come out as -1
    gdir = sg.Roadmap.from_knn(nbr, vis)
    Gd = nx.DiGraph()
    Gd.add_nodes_from(range(n))
    Gd.add_edges_from(gdir.edges().tolist())
    hd = gdir.shortest_paths(S, T)
    for q in range(128):
        s, t = int(S[q]), int(T[q])
        assert hd[q] == (nx.shortest_path_length(Gd, s, t) if nx.has_path(Gd, s, t) else -1)


def test_undirected_edge_lists(sg):
    n = 12
    src = np.array([0, 1, 2, 5, 5], np.int64)
    dst = np.array([1, 2, 0, 6, 5], np.int64)                      # 5 -> 5: a self loop is not mirrored
    g = sg.Roadmap(n)
    g.build_from_edges(src, dst, undirected=True)
    want = sorted({(a, b) for a, b in zip(src, dst)} | {(b, a) for a, b in zip(src, dst)})
    assert g.edges().tolist() == [list(e) for e in want]
    labels, n_comp = g.components()
    assert labels.cpu().tolist() == [0, 0, 0, 3, 4, 5, 5, 7, 8, 9, 10, 11] and n_comp == 9
    assert g.shortest_paths(np.array([2, 6, 0, 3]), np.array([1, 5, 5, 3])).tolist() == [1, 1, -1, 0]
