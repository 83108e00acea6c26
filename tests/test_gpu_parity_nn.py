"""-m gpu: node store, nearest / kNN / radius queries against golden fixtures and the
oracle.  Index sets bit-exact, fp64 distances bit-exact."""
import numpy as np
import pytest

import c_oracle as orc
from _util import golden, load_map, map_png
from sbp_env_b200.runtime import tune  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


def test_reference_kat_nearest(sg):
    # tests/test_rrtPlanner.py:136-148
    poses = np.array([[0, 0], [1, 1], [2, 2], [3, 3.0]])
    t = sg.NodeStore(2)
    for p in poses:
        t.append(p)
    assert len(t) == 4 and np.array_equal(t.poses, poses)
    for i, p in enumerate(poses):
        assert t.nearest(p + 1e-2) == i


@pytest.mark.parametrize("d", [2, 3, 4])
def test_golden_nn(sg, d):
    g = golden("nn")
    poses, X = g[f"d{d}_poses"], g[f"d{d}_X"]
    t = sg.NodeStore(d, capacity=16)          # forces capacity growth
    t.extend(poses[:1000])
    t.extend(poses[1000:])
    assert np.array_equal(t.poses, poses)
    idx, dist = t.knn(X, 1)
    assert np.array_equal(idx[:, 0], g[f"d{d}_nearest"])
    for k in (20, 7, 32):
        idx, dist = t.knn(X, k)
        if k <= 20:
            assert np.array_equal(idx, g[f"d{d}_knn20"][:, :k])
            assert np.array_equal(dist, g[f"d{d}_knn20_dist"][:, :k])
        else:
            oi, od = orc.knn(poses, X, k)
            assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    for ri in (0, 1):
        rp, ix = t.near(X, float(g[f"d{d}_prm_r{ri}"]), "full", closed=False)
        assert np.array_equal(rp, g[f"d{d}_prm_r{ri}_off"])
        assert np.array_equal(ix, g[f"d{d}_prm_r{ri}_idx"])
    mm = len(g[f"d{d}_xy_off"]) - 1
    rp, ix = t.near(X[:mm], float(g[f"d{d}_xy_r"]), "xy", closed=True)
    # engine.dist squares through libm pow: <= 1 ulp from the product, so
    # membership can only differ for a distance within 1 ulp of r -- none in this fixture
    assert np.array_equal(rp, g[f"d{d}_xy_off"])
    assert np.array_equal(ix, g[f"d{d}_xy_idx"])


@pytest.mark.parametrize("d,n,m,k", [(2, 50_000, 3000, 11), (2, 300_000, 5, 20), (4, 40_000, 700, 4),
                                     (2, 1, 10, 3), (3, 7, 300, 16), (6, 5000, 40, 2)])
def test_knn_vs_oracle(sg, d, n, m, k):
    """Both kernels (thread-per-query batched with chunk merge; CTA-per-query wide)."""
    rng = np.random.default_rng(d * 1000 + k)
    poses = rng.random((n, d)) * 300
    poses[: n // 10] = np.floor(poses[: n // 10] / 10) * 10      # exact ties
    X = rng.random((m, d)) * 300
    X[: m // 3] = np.floor(X[: m // 3] / 10) * 10 + 5
    t = sg.NodeStore(d)
    t.extend(poses)
    idx, dist = t.knn(X, k)
    oi, od = orc.knn(poses, X, k)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist, od)


@pytest.mark.parametrize("metric,closed", [("full", False), ("full", True), ("xy", True), ("xy", False)])
@pytest.mark.parametrize("d,n,m", [(2, 60_000, 1500), (2, 400_000, 3), (4, 30_000, 2)])
def test_near_vs_oracle(sg, metric, closed, d, n, m):
    rng = np.random.default_rng(n + m)
    poses = np.floor(rng.random((n, d)) * 400)                    # integer lattice: d == r happens
    X = np.floor(rng.random((m, d)) * 400)
    t = sg.NodeStore(d)
    t.extend(poses)
    for r in (5.0, 13.0, 0.0):                                    # 5 and 13 are hypotenuses
        rp, ix = t.near(X, r, metric, closed)
        orp, oix = orc.near(poses, X, r, metric, closed)
        assert np.array_equal(rp, orp), (r, metric, closed)
        assert np.array_equal(ix, oix)


def test_threshold_boundaries(sg):
    """d < r vs d <= r decided on d^2, at radii straddling rounded sqrts."""
    rng = np.random.default_rng(2)
    poses = rng.random((5000, 2)) * 50
    q = rng.random((1, 2)) * 50
    t = sg.NodeStore(2)
    t.extend(poses)
    d = np.sort(np.linalg.norm(poses - q, axis=1))
    for r in [d[10], np.nextafter(d[10], 0), np.nextafter(d[10], 99), d[777], 1e-300, 1e300, float("inf")]:
        for closed in (False, True):
            rp, ix = t.near(q, float(r), "full", closed)
            orp, oix = orc.near(poses, q, float(r), "full", closed)
            assert np.array_equal(ix, oix), (r, closed)


def test_device_tensors_and_prm_connect(sg):
    import torch

    free = load_map("intel_lab")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("intel_lab"))
    om = orc.Map(free)
    rng = np.random.default_rng(0)
    pts = rng.random((30000, 2)) * [W, H]
    pts = pts[om.feasible2d(pts)][:8000]
    pts[100] = pts[50]                                           # a duplicated node
    t = sg.NodeStore(2)
    t.extend(torch.from_numpy(pts).cuda())
    k = 10
    nbr, vis = sg.prm_connect_knn(cc, t, k)
    nbr, vis = nbr.cpu().numpy(), vis.cpu().numpy()
    oi, od = orc.knn(pts, pts, k + 1)
    for v in range(len(pts)):
        row = oi[v].tolist()
        row.remove(v) if v in row else row.pop()
        assert nbr[v].tolist() == row, v
    exp = om.visible2d(pts[nbr.reshape(-1)], np.repeat(pts, k, axis=0)).reshape(-1, k)
    exp &= nbr != np.arange(len(pts))[:, None]
    assert np.array_equal(vis, exp)
    # a shard of the rows gives the same rows
    nbr2, vis2 = sg.prm_connect_knn(cc, t, k, rows=(1234, 500))
    assert np.array_equal(nbr2.cpu().numpy(), nbr[1234:1734])
    assert np.array_equal(vis2.cpu().numpy(), vis[1234:1734])
    src, dst = sg.roadmap_edges_to_host(nbr, vis)
    assert len(src) == int(vis.sum())


@pytest.mark.parametrize("map_name,n_nodes,k", [("intel_lab", 6000, 10), ("maze1", 5, 9), ("4d", 3000, 4)])
def test_fused_knn_edges_visible_equals_the_two_launch_form(sg, map_name, n_nodes, k):
    """sbp_knn_edges_visible2d == sbp_nodes_knn_edges + sbp_visible2d (+ `& valid`): stored rows,
    a row slice, external queries, empty neighbour slots (k > n) and the visible counter."""
    import torch

    free = load_map(map_name)
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png(map_name))
    rng = np.random.default_rng(3)
    pts = rng.random((n_nodes, 2)) * [W, H]
    if n_nodes > 100:
        pts[77] = pts[76] = pts[75]                              # duplicates ahead of the row node
    t = sg.NodeStore(2)
    t.extend(pts)
    Xq = torch.from_numpy(rng.random((257, 2)) * [W, H]).cuda()
    for first, count, queries in [(0, n_nodes, None), (n_nodes // 3, n_nodes // 2, None), (0, 257, Xq)]:
        X = queries if queries is not None else t.rows_device(first, count)
        idx, _ = t.knn(X, k if queries is not None else k + 1)
        nbr_a, P, Q, valid = t.knn_edges(idx, first_row=first, queries=queries)
        vis_a = cc.visible_batch(P, Q).reshape(count, k) & valid.view(torch.bool).reshape(count, k)
        before = cc.stats.visible_cnt
        nbr_b, vis_b = cc.knn_edges_visible(t, idx, first_row=first, queries=queries)
        assert cc.stats.visible_cnt - before == count * k
        assert torch.equal(nbr_a, nbr_b)
        assert torch.equal(vis_a, vis_b)
        if n_nodes <= k:
            assert int((nbr_b < 0).sum()) > 0 and not bool(vis_b[nbr_b < 0].any())
        # the packed adjacency written by the kernel itself == pack_adjacency of its results
        code = torch.full((count, k), -7, dtype=torch.int32, device=nbr_b.device)
        nbr_p, vis_p = cc.knn_edges_visible(t, idx, first_row=first, queries=queries, packed_out=code)
        assert torch.equal(nbr_p, nbr_b) and torch.equal(vis_p, vis_b)
        assert torch.equal(code, sg.distributed.pack_adjacency(nbr_b, vis_b))
        # ... also when the destination is pinned host memory (zero-copy stores, bench.py e2e)
        code_h = torch.full((count, k), -7, dtype=torch.int32).pin_memory()
        cc.knn_edges_visible(t, idx, first_row=first, queries=queries, packed_out=(code_h.data_ptr(), False))
        torch.cuda.synchronize()
        assert torch.equal(code_h, code.cpu())
    # forced global-memory map path gives the same answer
    from sbp_env_b200 import _lib

    lib = _lib.load()
    tune("force_global", 1)
    try:
        idx, _ = t.knn(t.rows_device(0, n_nodes), k + 1)
        nbr_c, vis_c = cc.knn_edges_visible(t, idx)
    finally:
        tune("force_global", 0)
    nbr_d, vis_d = cc.knn_edges_visible(t, idx)
    assert torch.equal(nbr_c, nbr_d) and torch.equal(vis_c, vis_d)


def test_store_extend_from_pinned_and_pageable_host_tensors(sg):
    """A pinned host tensor is read by the append kernel itself (zero-copy), a pageable one goes
    through the host path: same store either way."""
    import torch

    rng = np.random.default_rng(12)
    pts = rng.random((5000, 2)) * 300
    stores = []
    for src in (torch.from_numpy(pts).pin_memory(), torch.from_numpy(pts.copy()), torch.from_numpy(pts).cuda(), pts):
        t = sg.NodeStore(2)
        t.extend(src)
        torch.cuda.synchronize()
        stores.append(t.rows_device(0, len(t)).cpu().numpy())
    for a in stores:
        assert np.array_equal(a, pts)


def test_empty_store_and_truncate(sg):
    t = sg.NodeStore(2)
    idx, dist = t.knn(np.zeros((3, 2)), 2)
    assert (idx == -1).all() and np.isinf(dist).all()
    rp, ix = t.near(np.zeros((3, 2)), 5.0)
    assert rp.tolist() == [0, 0, 0, 0] and len(ix) == 0
    t.extend(np.arange(20, dtype=float).reshape(10, 2))
    t.truncate(4)
    assert len(t) == 4 and t.nearest([100.0, 100.0]) == 3


def test_prm_connect_external_queries(sg):
    import torch

    free = load_map("maze1")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("maze1"))
    om = orc.Map(free)
    rng = np.random.default_rng(9)
    nodes = rng.random((5000, 2)) * [W, H]
    X = rng.random((700, 2)) * [W, H]
    t = sg.NodeStore(2)
    t.extend(nodes)
    k = 6
    nbr, vis, dist = sg.prm_connect_knn(cc, t, k, queries=torch.from_numpy(X).cuda(), return_dist=True)
    oi, od = orc.knn(nodes, X, k)
    assert np.array_equal(nbr.cpu().numpy(), oi) and np.array_equal(dist.cpu().numpy(), od)
    exp = om.visible2d(nodes[oi.reshape(-1)], np.repeat(X, k, axis=0)).reshape(-1, k)
    assert np.array_equal(vis.cpu().numpy(), exp)


def test_prm_connect_radius_is_the_reference_double_loop(sg):
    """prmPlanner.py:91-101 restated: for v in nodes: for m_g in near(v, r): skip self;
    visible(m_g, v)."""
    free = load_map("maze2")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("maze2"))
    om = orc.Map(free)
    rng = np.random.default_rng(21)
    pts = rng.random((3000, 2)) * [W, H]
    pts = pts[om.feasible2d(pts)][:1200]
    pts[40] = pts[10]                                     # duplicate of another node: NOT skipped
    t = sg.NodeStore(2)
    t.extend(pts)
    r = 30.0
    src, dst, vis = sg.prm_connect_radius(cc, t, r)
    es, ed = [], []
    for v in range(len(pts)):
        d = np.linalg.norm(pts - pts[v], axis=1)
        for m_g in np.flatnonzero(d < r):
            if m_g != v:
                es.append(m_g); ed.append(v)
    assert src.cpu().numpy().tolist() == es and dst.cpu().numpy().tolist() == ed
    assert np.array_equal(vis.cpu().numpy(), om.visible2d(pts[es], pts[ed]))


def test_arm_checker_map_mat_path(sg):
    """RobotArm4dCollisionChecker(map_mat=...) (collisionChecker.py:333-334)."""
    free = load_map("maze1")
    mat = free.T.astype(np.float64) * 255.0               # an image-like [H][W] matrix
    cc = sg.RobotArm4dCollisionChecker(None, map_mat=mat, stick_robot_length_config=[9.0, 4.0])
    assert cc._img.shape == free.shape and cc._img.dtype.kind == "i"
    om = orc.Map(free)
    rng = np.random.default_rng(8)
    C = rng.random((5000, 4)) * [free.shape[0], free.shape[1], 6, 6] - [0, 0, 3, 3]
    assert np.array_equal(cc.feasible_batch(C), om.arm_feasible(C, 9.0, 4.0))


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e3, 0.0), (1e6, 0.0), (1e9, 0.0), (1e14, 0.0),
                                          (1e16, 0.0), (1e200, 0.0), (1.0, 1e6), (1e-3, 4096.0),
                                          (1e-9, 1.0)])
def test_knn_fp32_prefilter_never_changes_results(sg, scale, offset):
    """The fp32 prefilter of the batched kNN scan is conservative: at every coordinate
    magnitude -- including clouds that fp32 cannot resolve at all (offset >> spread) and
    magnitudes where it is bypassed (>= 1e15) -- indices and fp64 distances are those of the
    oracle, and identical with the prefilter switched off."""
    from sbp_env_b200 import _lib

    rng = np.random.default_rng(int(abs(np.log10(scale)) * 7 + offset) % 2 ** 31)
    n, m, k = 6000, 800, 11
    poses = rng.random((n, 2)) * scale + offset
    poses[:300] = np.round(poses[:300] / scale * 8) / 8 * scale          # exact ties
    X = rng.random((m, 2)) * scale + offset
    t = sg.NodeStore(2)
    t.extend(poses)
    oi, od = orc.knn(poses, X, k)
    lib = _lib.load()
    try:
        for pf in (1, 0):
            tune("knn_prefilter", pf)
            idx, dist = t.knn(X, k)
            assert np.array_equal(idx, oi), (scale, offset, pf)
            assert np.array_equal(dist, od), (scale, offset, pf)
    finally:
        tune("knn_prefilter", 1)


def test_knn_prefilter_with_queries_far_outside_the_cloud(sg):
    rng = np.random.default_rng(77)
    poses = rng.random((5000, 4)) * 100
    X = np.concatenate([rng.random((300, 4)) * 100, rng.random((300, 4)) * 1e7, -rng.random((300, 4)) * 1e17,
                        rng.random((100, 4)) * 1e300])
    t = sg.NodeStore(4)
    t.extend(poses)
    idx, dist = t.knn(X, 5)
    oi, od = orc.knn(poses, X, 5)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)


@pytest.mark.parametrize("layout", ["uniform", "clustered", "duplicates", "line", "outliers"])
def test_knn_grid_seed_never_changes_results(sg, layout):
    """The grid-seeded a-priori bound (d = 2) only prunes insertions: results equal the
    oracle's and the unseeded scan's on benign and on adversarial point layouts."""
    from sbp_env_b200 import _lib

    rng = np.random.default_rng(len(layout))
    n, m, k = 20000, 1500, 11
    if layout == "uniform":
        poses = rng.random((n, 2)) * 600
    elif layout == "clustered":
        c = rng.random((20, 2)) * 600
        poses = c[rng.integers(0, 20, n)] + rng.standard_normal((n, 2)) * 0.5
    elif layout == "duplicates":
        poses = np.floor(rng.random((n, 2)) * 12) * 50.0          # 144 distinct positions
    elif layout == "line":
        poses = np.stack([rng.random(n) * 600, np.full(n, 123.25)], 1)   # zero extent in y
    else:
        poses = rng.random((n, 2)) * 10
        poses[:5] = [[1e9, 1e9], [-1e9, 3.0], [5.0, 1e12], [0.0, 0.0], [10.0, 10.0]]
    X = np.concatenate([poses[rng.integers(0, n, m // 2)] + rng.standard_normal((m // 2, 2)) * 0.3,
                        rng.random((m - m // 2, 2)) * 900 - 150])
    t = sg.NodeStore(2)
    t.extend(poses)
    oi, od = orc.knn(poses, X, k)
    lib = _lib.load()
    try:
        for seed in (1, 0):
            tune("knn_seed", seed)
            for kk in (k, 10, 1, 20):
                idx, dist = t.knn(X, kk)
                if kk == k:
                    assert np.array_equal(idx, oi) and np.array_equal(dist, od), (layout, seed)
                else:
                    ei, ed = orc.knn(poses, X, kk)
                    assert np.array_equal(idx, ei) and np.array_equal(dist, ed), (layout, seed, kk)
    finally:
        tune("knn_seed", 1)


@pytest.mark.parametrize("layout", ["uniform", "duplicates", "clustered", "nan_inf"])
@pytest.mark.parametrize("n", [2048, 2055, 20001, 66000])
def test_knn_filter_pipeline_never_changes_results(sg, layout, n):
    """Seeded filter (packed fp32 scan of 8-node groups, candidate lists) + exact fp64 refine:
    identical to the oracle and to the one-kernel scan for every chunking, for node counts
    that are not multiples of the group / tile size, for layouts whose candidate lists
    overflow (duplicates: the flagged blocks are redone by the one-kernel scan) and for
    queries / nodes with non-finite coordinates."""
    from sbp_env_b200 import _lib

    rng = np.random.default_rng(n + len(layout))
    m, k = 1300, 11
    if layout == "duplicates":
        poses = np.floor(rng.random((n, 2)) * 9) * 64.0              # 81 distinct positions
    elif layout == "clustered":
        c = rng.random((30, 2)) * 600
        poses = c[rng.integers(0, 30, n)] + rng.standard_normal((n, 2)) * 0.25
    else:
        poses = rng.random((n, 2)) * 600
    X = np.concatenate([poses[rng.integers(0, n, m // 2)] + rng.standard_normal((m // 2, 2)) * 0.3,
                        rng.random((m - m // 2, 2)) * 900 - 150])
    if layout == "nan_inf":
        poses[5] = [np.nan, 3.0]
        poses[77] = [np.inf, 10.0]
        poses[n - 1] = [250.0, -np.inf]
        X[3] = [np.nan, np.nan]
        X[40] = [np.inf, 0.0]
        X[41] = [100.0, np.nan]
    t = sg.NodeStore(2)
    t.extend(poses)
    lib = _lib.load()
    try:
        for kk in (k, 1, 20, 32):
            ei, ed = orc.knn(poses, X, kk)
            for filt, chunks, cells in ((1, 0, 1), (1, 0, 3), (1, 0, 0), (1, 1, 0), (1, 7, 0), (0, 0, 0)):
                tune("knn_filter", filt)
                tune("knn_filter_chunks", chunks)
                tune("knn_cells", cells)
                idx, dist = t.knn(X, kk)
                assert np.array_equal(idx, ei), (layout, n, kk, filt, chunks, cells)
                assert np.array_equal(dist, ed, equal_nan=True), (layout, n, kk, filt, chunks, cells)
    finally:
        tune("knn_filter", 1)
        tune("knn_filter_chunks", 0)
        tune("knn_cells", 2)


@pytest.mark.parametrize("layout", ["uniform", "clustered", "duplicates", "line", "outliers", "lattice", "nan_inf"])
@pytest.mark.parametrize("d", [2, 3])
def test_knn_cell_grid_answers_equal_the_oracle(sg, layout, d):
    """k_knn_cells answers a query exactly from the rectangle of grid cells that covers its
    K-th seed distance; queries whose rectangle holds more than the limit stay with the
    brute-force filter.  Every limit (0: only rectangles inside the seed block, small: a mix of
    both paths in one batch, huge: everything from the grid) gives the oracle's rows -- also
    with exact distance ties (lattice: many equidistant nodes, the lower index must win)."""
    from sbp_env_b200 import _lib

    rng = np.random.default_rng(len(layout) + d)
    n, m = 30000, 2100
    if layout == "uniform":
        poses = rng.random((n, d)) * 600
    elif layout == "clustered":
        c = rng.random((20, d)) * 600
        poses = c[rng.integers(0, 20, n)] + rng.standard_normal((n, d)) * 0.5
    elif layout == "duplicates":
        poses = np.floor(rng.random((n, d)) * 12) * 50.0
    elif layout == "line":
        poses = rng.random((n, d)) * 600
        poses[:, 1] = 123.25                                       # zero extent in y
    elif layout == "lattice":
        poses = np.floor(rng.random((n, d)) * 200) * 3.0          # integer lattice with repeats
    elif layout == "outliers":
        poses = rng.random((n, d)) * 10
        poses[:3, :2] = [[1e9, 1e9], [-1e9, 3.0], [5.0, 1e12]]
    else:
        poses = rng.random((n, d)) * 600
        poses[5, 0] = np.nan
        poses[77, 0] = np.inf
        poses[n - 1, 1] = -np.inf
    X = np.concatenate([poses[rng.integers(0, n, m // 3)],
                        poses[rng.integers(0, n, m // 3)] + rng.standard_normal((m // 3, d)) * 0.3,
                        rng.random((m - 2 * (m // 3), d)) * 900 - 150])
    if layout == "nan_inf":
        X[3] = np.nan
        X[40, 0] = np.inf
        X[41, 1] = np.nan
    t = sg.NodeStore(d)
    t.extend(poses)
    lib = _lib.load()
    try:
        for kk in (11, 1, 4, 21, 32):
            ei, ed = orc.knn(poses, X, kk)
            for mode, limit, density in ((1, 0, 50), (1, 48, 50), (1, 1024, 50), (1, 1 << 30, 50),
                                         (3, 0, 50), (3, 200, 50), (3, 1024, 300), (3, 1 << 30, 50),
                                         (3, 1024, 5)):
                tune("knn_cells", mode)                  # 1 = warp, 3 = thread per query
                tune("knn_cells_limit", limit)
                tune("knn_cell_density", density)
                idx, dist = t.knn(X, kk)
                assert np.array_equal(idx, ei), (layout, d, kk, mode, limit, density)
                assert np.array_equal(dist, ed, equal_nan=True), (layout, d, kk, mode, limit, density)
        # as many queries as nodes (the PRM build: the stored rows themselves, and an unrelated
        # batch of the same size): the queries are then walked in cell-sorted order
        for Xs in (poses, np.roll(X, 7, axis=0)[np.arange(n) % len(X)]):
            ei, ed = orc.knn(poses, Xs, 11)
            for mode, order in ((3, 1), (3, 0), (2, 1)):
                tune("knn_cells", mode)
                tune("knn_cells_limit", 1024)
                tune("knn_cell_density", 150)
                tune("knn_cells_order", order)
                idx, dist = t.knn(Xs, 11)
                assert np.array_equal(idx, ei), (layout, d, mode, order)
                assert np.array_equal(dist, ed, equal_nan=True), (layout, d, mode, order)
    finally:
        tune("knn_cells", 2)
        tune("knn_cells_limit", 1024)
        tune("knn_cell_density", 150)
        tune("knn_cells_order", 1)


def test_prm_connect_graph_replay_equals_eager(sg):
    """PRMConnectGraph / CapturedStep: the captured step replays to the same neighbour and
    visibility blocks as the eager call, follows in-place changes of the query tensor, keeps
    the Stats accounting of visible_batch and refuses a store that changed size."""
    import torch

    free = load_map("intel_lab")
    cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8))
    rng = np.random.default_rng(5)
    pts = rng.random((9000, 2)) * free.shape
    pts = pts[cc.feasible_batch(pts)][:4000]
    tree = sg.NodeStore(2, capacity=8192)
    tree.extend(pts)
    nbr_e, vis_e = sg.prm_connect_knn(cc, tree, 10)
    g = sg.PRMConnectGraph(cc, tree, 10)
    before = cc.stats.visible_cnt
    for _ in range(2):
        nbr, vis = g.replay()
    torch.cuda.synchronize()
    assert torch.equal(nbr, nbr_e) and torch.equal(vis, vis_e)
    assert cc.stats.visible_cnt - before == 2 * 4000 * 10
    # same edges as the oracle
    oi, _ = orc.knn(pts, pts, 11)
    assert np.array_equal(np.sort(nbr.cpu().numpy(), 1), np.sort(oi[:, 1:], 1))
    # external queries, updated in place between replays
    Xq = torch.from_numpy(rng.random((1500, 2)) * free.shape).cuda()
    gq = sg.PRMConnectGraph(cc, tree, 7, queries=Xq)
    for _ in range(2):
        Xq.copy_(torch.from_numpy(rng.random((1500, 2)) * free.shape))
        nbr, vis = gq.replay()
        ne, ve = sg.prm_connect_knn(cc, tree, 7, queries=Xq)
        assert torch.equal(nbr, ne) and torch.equal(vis, ve)
    tree.append(np.array([1.0, 2.0]))
    with pytest.raises(RuntimeError):
        g.replay()


def test_knn_seeds_in_sparse_pockets_and_far_queries(sg):
    """Grid seeds: isolated pockets with fewer than k nodes nearby, queries far outside the
    cloud, and very dense clumps -- the block of cells grows until it holds k nodes and the
    bound uses all of its nodes; results equal the oracle's."""
    rng = np.random.default_rng(99)
    clumps = np.concatenate([c + rng.standard_normal((3000, 2)) * 0.02 for c in ([10.0, 10.0], [500.0, 20.0])])
    lonely = np.array([[250.0, 250.0], [251.0, 250.0], [900.0, 900.0], [-300.0, 40.0]])
    sparse = rng.random((3000, 2)) * 1000
    poses = np.concatenate([clumps, lonely, sparse])
    X = np.concatenate([lonely + 0.25, [[5000.0, 5000.0], [-1e6, 0.0], [10.0, 10.0], [500.0, 20.0]],
                        rng.random((600, 2)) * 1200 - 100, clumps[::50] + 1e-3])
    t = sg.NodeStore(2)
    t.extend(poses)
    for k in (11, 1, 32):
        idx, dist = t.knn(X, k)
        oi, od = orc.knn(poses, X, k)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), k


@pytest.mark.parametrize("d", [2, 3, 4, 6])
def test_near_filtered_scan_never_changes_results(sg, d):
    """Filtered radius query (centred fp32 shadow + packed scan, exact fp64 decision of the
    groups that pass): row pointers and index lists equal the oracle's and the unfiltered
    kernels' for both metrics and comparisons, for radii from 0 to larger than the cloud, on
    lattice data where d == r happens, with an unfilterable outlier node and with queries
    that are NaN, far away or beyond fp32 range."""
    from sbp_env_b200 import _lib

    rng = np.random.default_rng(40 + d)
    n, m = 23_456, 700
    poses = np.floor(rng.random((n, d)) * 250)
    X = np.floor(rng.random((m, d)) * 250)
    X[5] = np.nan
    X[6, 0] = 1e300
    X[7] = -4000.0
    X[8, 1] = np.inf
    lib = _lib.load()
    for variant in ("plain", "outlier"):
        if variant == "outlier":
            poses = poses.copy()
            poses[123] = 1e14            # the band E exceeds r^2: every query becomes unfilterable
        t = sg.NodeStore(d)
        t.extend(poses)
        try:
            for metric in ("full", "xy"):
                for closed in (False, True):
                    for r in (0.0, 5.0, 13.0, 1000.0) if variant == "plain" else (13.0,):
                        orp, oix = orc.near(poses, X, r, metric, closed)
                        for filt in (1, 0):
                            tune("near_filter", filt)
                            rp, ix = t.near(X, r, metric, closed)
                            assert np.array_equal(rp, orp), (variant, metric, closed, r, filt)
                            assert np.array_equal(ix, oix), (variant, metric, closed, r, filt)
        finally:
            tune("near_filter", 1)


def test_pack_adjacency_device_kernel_roundtrip(sg):
    """distributed.pack_adjacency on CUDA tensors (one launch of the pack kernel) equals the
    torch expression and unpacks to the inputs, for sizes that are not multiples of 4 and for
    empty neighbour slots (-1)."""
    import torch

    D = sg.distributed
    g = torch.Generator(device="cuda")
    g.manual_seed(0)
    for m, k in ((1, 1), (3, 7), (1001, 10), (4096, 4)):
        nbr = torch.randint(-1, 1 << 29, (m, k), generator=g, device="cuda", dtype=torch.int64)
        vis = torch.rand((m, k), generator=g, device="cuda") < 0.5
        code = D.pack_adjacency(nbr, vis)
        assert code.dtype == torch.int32 and code.shape == (m, k)
        assert torch.equal(code, D.pack_adjacency(nbr.cpu(), vis.cpu()).cuda())
        n2, v2 = D.unpack_adjacency(code)
        assert torch.equal(n2, nbr) and torch.equal(v2, vis)


def test_captured_step_host_to_host_follows_its_inputs(sg):
    """CapturedStep around a whole host-to-host step (H2D of a sample batch from pinned memory,
    node store rebuilt, connection, packed adjacency back to pinned memory): every replay works
    on the CURRENT contents of the pinned input and equals the eager computation."""
    import torch

    free = load_map("room1")
    cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8))
    rng = np.random.default_rng(21)
    m, k = 3000, 6

    def batch():
        pts = rng.random((3 * m, 2)) * free.shape
        return pts[cc.feasible_batch(pts)][:m]

    pinned = torch.from_numpy(batch()).pin_memory()
    x_dev = torch.empty((m, 2), dtype=torch.float64, device="cuda")
    code_h = torch.empty((m, k), dtype=torch.int32).pin_memory()
    tree = sg.NodeStore(2, capacity=4096)

    def body():
        x_dev.copy_(pinned, non_blocking=True)
        tree.truncate(0)
        tree.extend(x_dev)
        nb, vs = sg.prm_connect_knn(cc, tree, k)
        code_h.copy_(sg.distributed.pack_adjacency(nb, vs), non_blocking=True)

    step = sg.CapturedStep(body)
    for _ in range(3):
        pinned.copy_(torch.from_numpy(batch()))           # new samples in the same pinned buffer
        step.replay()
        torch.cuda.synchronize()
        nbr, vis = sg.distributed.unpack_adjacency(code_h)
        pts = pinned.numpy()
        oi, _ = orc.knn(pts, pts, k + 1)
        want = np.array([[j for j in r if j != i][:k] for i, r in enumerate(oi)])
        assert np.array_equal(nbr.numpy(), want)
        om = orc.Map(free)
        assert np.array_equal(vis.numpy(), om.visible2d(pts[want.reshape(-1)], np.repeat(pts, k, 0)).reshape(m, k))
        assert len(tree) == m and np.array_equal(tree.poses, pts)


@pytest.mark.parametrize("d,layout", [(4, "arm"), (3, "cube"), (6, "cube"), (4, "planar"), (3, "nan_inf")])
def test_knn_filter_pipeline_in_more_than_two_dimensions(sg, d, layout):
    """d > 2: the seeds come from the x-y cell grid, the bound is the exact d-dimensional K-th
    distance among the nodes found there.  Arm-like clouds (x, y in pixels, angles in radians)
    and clouds that are flat in the extra coordinates filter well; isotropic cubes give loose
    bounds, are recognised (expected count inside the bound) and go to the one-kernel scan.
    Either way the result equals the oracle's and the unfiltered kernels'."""
    from sbp_env_b200 import _lib

    rng = np.random.default_rng(10 * d + len(layout))
    n, m = 30_011, 2100
    if layout == "arm":
        poses = rng.random((n, 4)) * [1024, 1024, 2 * np.pi, 2 * np.pi] - [0, 0, np.pi, np.pi]
    elif layout == "planar":
        poses = np.concatenate([rng.random((n, 2)) * 500, rng.standard_normal((n, 2)) * 1e-3], 1)
    else:
        poses = rng.random((n, d)) * 300
    X = poses[rng.integers(0, n, m)] + rng.standard_normal((m, d)) * 0.2
    X[: m // 4] = rng.random((m // 4, d)) * (poses.max(0) - poses.min(0)) * 1.2 + poses.min(0)
    if layout == "nan_inf":
        poses[3] = np.nan
        poses[9, 2] = np.inf
        X[5, 1] = np.nan
        X[6] = 1e300
    t = sg.NodeStore(d)
    t.extend(poses)
    lib = _lib.load()
    try:
        for k in (11, 1, 32):
            ei, ed = orc.knn(poses, X, k)
            for filt, cells in ((1, 1), (1, 3), (1, 0), (0, 0)):
                tune("knn_filter", filt)
                tune("knn_cells", cells)
                idx, dist = t.knn(X, k)
                assert np.array_equal(idx, ei), (d, layout, k, filt, cells)
                assert np.array_equal(dist, ed, equal_nan=True), (d, layout, k, filt, cells)
    finally:
        tune("knn_filter", 1)
        tune("knn_cells", 2)


def test_cfg2_full_size_connection_equals_oracle(sg):
    """BASELINE.json configs[1] at FULL size -- intel_lab, 50 000 free samples, k = 10, i.e. the
    bench workload itself: every neighbour row and every visibility bit of the 500 000
    candidate edges equals the oracle's (the oracle needs ~0.5 s on 16 cores for this)."""
    from _util import draw_free_samples

    free = load_map("intel_lab")
    cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8))
    nodes = draw_free_samples(free, 50_000, 0, cc.feasible_batch)
    tree = sg.NodeStore(2, capacity=50_000)
    tree.extend(nodes)
    k = 10
    nbr, vis = sg.prm_connect_knn(cc, tree, k)
    oi, _ = orc.knn(nodes, nodes, k + 1)
    rows = np.arange(len(nodes))[:, None]
    self_col = (oi == rows)
    # drop the self entry (or the last column when a duplicate precedes it)
    drop = np.where(self_col.any(1), self_col.argmax(1), k)
    keep = np.ones_like(oi, bool)
    keep[np.arange(len(nodes)), drop] = False
    want = oi[keep].reshape(len(nodes), k)
    assert np.array_equal(nbr.cpu().numpy(), want)
    om = orc.Map(free)
    want_vis = om.visible2d(nodes[want.reshape(-1)], np.repeat(nodes, k, 0)).reshape(-1, k)
    assert np.array_equal(vis.cpu().numpy(), want_vis)


def test_million_node_knn_properties(sg):
    """cfg5-scale store (1M nodes): size-independent properties of the kNN answer for 20 000
    queries -- rows ordered by (distance, index), distances equal to the fp64 recomputation, no
    stored node closer than the k-th neighbour among 64 random probes per query -- plus exact
    equality with the oracle on 200 sampled rows."""
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(12)
    n, m, k = 1_000_000, 20_000, 10
    P = torch.rand((n, 2), generator=g, device="cuda", dtype=torch.float64) * 32768
    P[:1000] = torch.floor(P[:1000])                                   # exact ties on a lattice
    X = torch.rand((m, 2), generator=g, device="cuda", dtype=torch.float64) * 32768
    X[:500] = torch.floor(X[:500]) + 0.5
    tree = sg.NodeStore(2, capacity=n)
    tree.extend(P)
    idx, dist = tree.knn(X, k)
    assert int(idx.min()) >= 0 and int(idx.max()) < n
    d = (P[idx] - X[:, None, :])
    re = torch.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1])     # same operation order as the kernel
    assert torch.equal(re, dist)
    ordered = (dist[:, 1:] > dist[:, :-1]) | ((dist[:, 1:] == dist[:, :-1]) & (idx[:, 1:] > idx[:, :-1]))
    assert bool(ordered.all())
    probe = torch.randint(0, n, (m, 64), generator=g, device="cuda")
    dp = P[probe] - X[:, None, :]
    dpr = torch.sqrt(dp[..., 0] * dp[..., 0] + dp[..., 1] * dp[..., 1])
    in_row = (probe[:, :, None] == idx[:, None, :]).any(-1)
    assert bool(((dpr >= dist[:, -1:]) | in_row).all())
    rows = torch.randperm(m, generator=g, device="cuda")[:200].cpu().numpy()
    oi, od = orc.knn(P.cpu().numpy(), X.cpu().numpy()[rows], k)
    assert np.array_equal(idx.cpu().numpy()[rows], oi) and np.array_equal(dist.cpu().numpy()[rows], od)


def test_nearest_is_np_argmin_also_on_nan(sg):
    """find_nearest_neighbour_idx (rrtPlanner.py:278-279) is np.argmin: the FIRST node at a NaN
    distance wins as soon as there is one -- NaN / inf - inf node coordinates, NaN or inf queries --
    where kNN (stable argsort) ranks NaN last.  Against the oracle's orc_nearest and numpy."""
    import torch

    rng = np.random.default_rng(21)
    for d in (2, 4):
        n = 5000
        P = rng.random((n, d)) * 300
        X = rng.random((40, d)) * 300
        tree = sg.NodeStore(d)
        tree.extend(P)
        # finite store, finite queries: plain first minimum
        gi, gd = tree.nearest_batch(X)
        oi, od = orc.nearest(P, X)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
        # NaN / inf queries against a finite store
        Xb = X.copy()
        Xb[0, 0] = np.nan
        Xb[1, d - 1] = np.inf
        Xb[2] = -np.inf
        gi, gd = tree.nearest_batch(Xb)
        oi, od = orc.nearest(P, Xb)
        want = np.array([int(np.argmin(np.linalg.norm(P - x, axis=1))) for x in Xb])
        assert np.array_equal(oi, want)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od, equal_nan=True)
        assert tree.nearest(Xb[0]) == 0                      # every distance is NaN: index 0
        # non-finite node coordinates: the first NaN-distance node, not the true minimum
        P2 = P.copy()
        P2[1234, 0] = np.nan
        P2[77, d - 1] = np.inf                               # inf - inf = NaN only for the inf query (Xb[1])
        P2[4000] = np.nan
        t2 = sg.NodeStore(d)
        t2.extend(P2)
        gi, gd = t2.nearest_batch(Xb)
        oi, od = orc.nearest(P2, Xb)
        with np.errstate(invalid="ignore"):
            want = np.array([int(np.argmin(np.linalg.norm(P2 - x, axis=1))) for x in Xb])
        assert np.array_equal(oi, want)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od, equal_nan=True)
        assert (gi[3:] == 1234).all() and gi[1] == 77
        # device tensors in, device tensors out
        di, dd = t2.nearest_batch(torch.from_numpy(Xb).cuda())
        assert np.array_equal(di.cpu().numpy(), oi)
        # kNN keeps the argsort rule (NaN last)
        ki, _ = t2.knn(X[:5], 1)
        assert (ki[:, 0] != 1234).all()


def test_knn_precision_32_same_rows_fp32_distances(sg):
    """BASELINE.json north_star: "distances must agree within 1e-6 relative in fp32 (exact in fp64)".
    precision=32 returns the rows of the exact selection (index sets bit-exact, ties included) with
    float32 distances: the fp64 distance rounded once."""
    import torch

    rng = np.random.default_rng(32)
    for d, n, m, k in ((2, 30000, 5000, 10), (2, 1500, 40, 20), (4, 8000, 600, 11), (3, 4000, 3, 1)):
        P = rng.random((n, d)) * 30000
        P[: n // 10] = np.floor(P[: n // 10])                     # lattice points: exact ties
        X = rng.random((m, d)) * 30000
        X[: m // 4] = np.floor(X[: m // 4])
        tree = sg.NodeStore(d)
        tree.extend(P)
        oi, od = orc.knn(P, X, k)
        gi, g32 = tree.knn(X, k, precision=32)
        assert g32.dtype == np.float32
        assert np.array_equal(gi, oi)
        assert np.array_equal(g32, od.astype(np.float32))
        assert (np.abs(g32.astype(np.float64) - od) <= 1e-6 * od).all()
        di, d32 = tree.knn(torch.from_numpy(X).cuda(), k, precision=32)
        assert d32.dtype == torch.float32
        assert np.array_equal(di.cpu().numpy(), oi) and np.array_equal(d32.cpu().numpy(), g32)
    empty = sg.NodeStore(2)
    ei, ed = empty.knn(np.zeros((3, 2)), 2, precision=32)
    assert (ei == -1).all() and np.isinf(ed).all() and ed.dtype == np.float32
