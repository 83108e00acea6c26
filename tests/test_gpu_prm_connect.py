"""-m gpu: the one-call roadmap connection (sbp_prm_connect2d: exact kNN of the stored rows +
row-ordered visibility kernel on the rows8 map copy) against the oracle and against the generic
two-call path, bit for bit; its cell-sorted parts (the multi-GPU shards); the packed outputs."""
import os
import sys

import numpy as np
import pytest

import c_oracle as orc
from _util import load_map, map_png

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu
SENT = -(1 << 31)


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


def oracle_connect(om, nodes, k):
    """prmPlanner.py:91-101 with k-NN candidates: kNN (k+1), self dropped, visible(m_g, v)."""
    n = len(nodes)
    oi, _ = orc.knn(nodes, nodes, k + 1)
    rows = np.arange(n)[:, None]
    is_self = oi == rows
    drop = np.where(is_self.any(1), is_self.argmax(1), k)
    keep = np.ones_like(oi, bool)
    keep[np.arange(n), drop] = False
    nbr = oi[keep].reshape(n, k)
    ok = nbr >= 0
    vis = om.visible2d(nodes[np.where(ok, nbr, 0).reshape(-1)], np.repeat(nodes, k, 0)).reshape(n, k) & ok
    return nbr, vis


@pytest.mark.parametrize("name,n,k", [("maze1", 6000, 7), ("intel_lab", 20000, 10), ("room1", 3000, 1),
                                      ("room1", 2500, 31), ("maze1", 300, 5), ("maze1", 40, 10),
                                      ("intel_lab", 8000, 14), ("maze1", 5000, 19)])
def test_one_call_connection_equals_oracle(sg, name, n, k):
    import torch

    free = load_map(name)
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png(name))
    om = orc.Map(free)
    rng = np.random.default_rng(n + k)
    nodes = rng.random((n, 2)) * [W, H]
    nodes[: n // 20] = np.floor(nodes[: n // 20])               # lattice points: exact ties
    nodes[n // 20: n // 20 + 10] = nodes[:10]                    # duplicates of stored rows (self-drop rule)
    tree = sg.NodeStore(2)
    tree.extend(nodes)
    want_nbr, want_vis = oracle_connect(om, nodes, k)
    packed = torch.full((n, k), SENT, dtype=torch.int32, device="cuda")
    before = cc.stats.visible_cnt
    nbr, vis = sg.prm_connect_knn(cc, tree, k, packed_out=packed)
    assert cc.stats.visible_cnt - before == n * k
    assert np.array_equal(nbr.cpu().numpy(), want_nbr)
    assert np.array_equal(vis.cpu().numpy(), want_vis)
    pn, pv = sg.distributed.unpack_adjacency(packed)
    assert torch.equal(pn, nbr) and torch.equal(pv, vis)
    # the generic two-call path (kNN call + general visibility kernel) gives the same rows
    n2, v2 = sg.prm_connect_knn(cc, tree, k, rows=(0, n))
    assert torch.equal(n2, nbr) and torch.equal(v2, vis)


def test_wraparound_and_out_of_range_nodes(sg):
    """Nodes with negative coordinates (numpy wrap-around) and outside the map take
    the general walker inside the same kernel."""
    free = load_map("room1")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("room1"))
    om = orc.Map(free)
    rng = np.random.default_rng(3)
    n, k = 4000, 6
    nodes = (rng.random((n, 2)) * 2.4 - 1.2) * [W, H]            # [-1.2 W, 1.2 W): wraps and out of range
    nodes[:200] = rng.random((200, 2)) * [W, H] - [W, H]         # fully negative: wrapped copies
    tree = sg.NodeStore(2)
    tree.extend(nodes)
    want_nbr, want_vis = oracle_connect(om, nodes, k)
    nbr, vis = sg.prm_connect_knn(cc, tree, k)
    assert np.array_equal(nbr.cpu().numpy(), want_nbr)
    assert np.array_equal(vis.cpu().numpy(), want_vis)
    assert 0 < int(vis.sum()) < n * k


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_cell_sorted_parts_cover_every_row_once(sg, parts):
    import torch

    free = load_map("intel_lab")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("intel_lab"))
    rng = np.random.default_rng(parts)
    n, k = 30011, 10
    nodes = rng.random((n, 2)) * [W, H]
    nodes[:500] = nodes[500:1000]                                # several nodes per cell position
    tree = sg.NodeStore(2)
    tree.extend(nodes)
    full_nbr, full_vis = sg.prm_connect_knn(cc, tree, k)
    full = sg.distributed.pack_adjacency(full_nbr, full_vis)
    owners = torch.zeros(n, dtype=torch.int32, device="cuda")
    merged = torch.full((n, k), SENT, dtype=torch.int32, device="cuda")
    for i in range(parts):
        mine = torch.full((n, k), SENT, dtype=torch.int32, device="cuda")
        assert sg.prm_connect_knn(cc, tree, k, part=(i, parts), packed_out=mine, want_rows=False) is None
        rows = mine[:, 0] != SENT
        base, rem = divmod(n, parts)
        assert int(rows.sum()) == base + (1 if i < rem else 0)   # balanced runs of the cell-sorted order
        assert torch.equal(mine[rows], full[rows])
        owners += rows.to(torch.int32)
        merged = torch.maximum(merged, mine)
        # a part is spatially compact: its rows span a band of grid rows, not the whole map
        ys = torch.from_numpy(nodes[:, 1]).cuda()[rows]
        assert float(ys.max() - ys.min()) < 1.5 * H / parts + 40
    assert bool((owners == 1).all()) and torch.equal(merged, full)


def test_pinned_host_output_and_graph_replay(sg):
    import torch

    free = load_map("maze2")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("maze2"))
    rng = np.random.default_rng(9)
    n, k = 9000, 8
    nodes = rng.random((n, 2)) * [W, H]
    tree = sg.NodeStore(2, capacity=n)
    tree.extend(nodes)
    nbr, vis = sg.prm_connect_knn(cc, tree, k)
    want = sg.distributed.pack_adjacency(nbr, vis).cpu()
    host = torch.full((n, k), SENT, dtype=torch.int32).pin_memory()
    sg.prm_connect_knn(cc, tree, k, packed_out=(host.data_ptr(), False), want_rows=False)
    torch.cuda.synchronize()
    assert torch.equal(host, want)
    dev_out = torch.full((n, k), SENT, dtype=torch.int32, device="cuda")
    g = sg.PRMConnectGraph(cc, tree, k, packed_out=dev_out, want_rows=False)
    for _ in range(3):
        dev_out.fill_(SENT)
        assert g.replay() == (None, None)
        torch.cuda.synchronize()
        assert torch.equal(dev_out.cpu(), want)
    # a later, larger eager call makes the library's work buffers grow: the captured graph notices
    # (scratch generation) and recaptures instead of replaying on freed memory
    big = sg.NodeStore(2)
    big.extend(rng.random((60000, 2)) * [W, H])
    sg.prm_connect_knn(cc, big, k)
    dev_out.fill_(SENT)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(dev_out.cpu(), want)


def test_cfg5_reduced_scale_equals_oracle(sg):
    """BASELINE.json configs[4] shrunk to a 32 x 32-cell maze (4096^2 px, 15 625 nodes): the map is
    L2-resident (global-memory path of every kernel), every row and bit against the oracle."""
    import torch
    import workloads

    dev = torch.device("cuda", 0)
    mz, cc, tree, nd, S, T = workloads.cfg5_on_device(dev, scale=0.125)
    nodes = nd.cpu().numpy()
    om = orc.Map(mz.raster_host())
    k = 10
    want_nbr, want_vis = oracle_connect(om, nodes, k)
    try:
        # both walk kernels, with and without the free-run planes (bounding box of tiles all free)
        # fused form (the kNN kernel settles all-free boxes, k_prm_walk_todo the rest) and the separate
        # walk kernels, with and without the free-run planes
        for fuse, rect, srt in ((1, 1, 1), (1, 0, 1), (0, 1, 1), (0, 0, 1), (0, 1, 0), (0, 0, 0)):
            sg.runtime.tune("prm_fuse", fuse)
            sg.runtime.tune("tile_rect", rect)
            sg.runtime.tune("prm_walk_sorted", srt)
            nbr, vis = sg.prm_connect_knn(cc, tree, k)
            assert np.array_equal(nbr.cpu().numpy(), want_nbr), (fuse, rect, srt)
            assert np.array_equal(vis.cpu().numpy(), want_vis), (fuse, rect, srt)
            packed = torch.full((len(nodes), k), -1, dtype=torch.int32, device=dev)
            order = torch.full((len(nodes),), -1, dtype=torch.int32, device=dev)
            sg.prm_connect_knn(cc, tree, k, packed_out=packed, order_out=order, want_rows=False)
            assert torch.equal(packed, sg.distributed.pack_adjacency(nbr, vis)[order.long()]), (fuse, rect, srt)
    finally:
        sg.runtime.tune("prm_fuse", 1)
        sg.runtime.tune("tile_rect", 1)
        sg.runtime.tune("prm_walk_sorted", 1)
    assert 0.5 < float(vis.float().mean()) < 0.99
    # the generic batched entry on the same candidate edges (k_visible2d_tiles)
    Pn = nodes[want_nbr.reshape(-1)]
    Qn = np.repeat(nodes, k, axis=0)
    assert np.array_equal(cc.visible_batch(Pn, Qn), want_vis.reshape(-1).astype(bool))
    # the analytic occupancy predicate, the device raster and the host raster agree
    P = np.random.default_rng(0).random((200000, 2)) * [mz.W, mz.H]
    assert np.array_equal(cc.feasible_batch(P), mz.free_points(P))
    assert np.array_equal(om.feasible2d(P), mz.free_points(P))


def test_cell_sorted_rows_and_order(sg):
    """sbp_prm_connect2d_sorted: the same rows in cell-sorted order (row p = node order[p]), whole and
    in parts (a part = one contiguous block of rows), into device and pinned host memory; the
    roadmap built from them equals the one built from the node-order rows."""
    import torch

    free = load_map("intel_lab")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("intel_lab"))
    rng = np.random.default_rng(17)
    n, k = 25013, 10
    nodes = rng.random((n, 2)) * [W, H]
    nodes[:300] = np.floor(nodes[:300])
    tree = sg.NodeStore(2)
    tree.extend(nodes)
    nbr, vis = sg.prm_connect_knn(cc, tree, k)
    want = sg.distributed.pack_adjacency(nbr, vis)
    order = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    packed = torch.full((n, k), SENT, dtype=torch.int32, device="cuda")
    n2, v2 = sg.prm_connect_knn(cc, tree, k, packed_out=packed, order_out=order)
    o = order.long()
    assert torch.equal(torch.sort(o).values, torch.arange(n, device="cuda"))     # a permutation
    assert torch.equal(packed, want[o]) and torch.equal(n2, nbr[o]) and torch.equal(v2, vis[o])
    # cell-sorted: consecutive rows are spatial neighbours
    P = torch.from_numpy(nodes).cuda()[o]
    assert float((P[1:] - P[:-1]).norm(dim=1).median()) < 12.0
    # parts: contiguous blocks, straight into pinned host memory (coalesced zero-copy stores)
    host = torch.full((n, k), SENT, dtype=torch.int32).pin_memory()
    parts = 3
    for i in range(parts):
        order2 = torch.full_like(order, -1)
        sg.prm_connect_knn(cc, tree, k, part=(i, parts), packed_out=(host.data_ptr(), False), want_rows=False,
                           order_out=order2)
        torch.cuda.synchronize()
        lo, cnt = sg.distributed.shard_range(n, i, parts)
        # a part builds the grid for its band of cell rows and writes its slice of the permutation:
        # the same one whatever the split
        assert torch.equal(order2[lo:lo + cnt], order[lo:lo + cnt])
        assert bool((order2[:lo] == -1).all()) and bool((order2[lo + cnt:] == -1).all())
        assert torch.equal(host[lo:lo + cnt], want[o][lo:lo + cnt].cpu())
        assert bool((host[lo + cnt:] == SENT).all())
    # roadmap from the sorted rows == roadmap from the node-order rows
    for und in (False, True):
        a = sg.Roadmap(n).build_from_packed(want, undirected=und)
        b = sg.Roadmap(n).build_from_packed(packed, undirected=und, row_ids=order)
        assert a.n_edges == b.n_edges
        ra, ca = a.csr()
        rb, cb = b.csr()
        assert torch.equal(ra, rb)
        ea, eb = a.edges(), b.edges()
        assert np.array_equal(np.asarray(ea), np.asarray(eb))


def test_rows_the_fused_kernel_leaves_to_the_exact_path(sg):
    """Near ties (distinct d^2 whose roots round to one value: the lower index must win although its
    d^2 is larger), non-finite stores, isolated nodes whose rectangle exceeds the limit: rows the
    fused kernel does not settle come out of the exact pipeline, identical to the oracle."""
    free = load_map("room1")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("room1"))
    om = orc.Map(free)
    rng = np.random.default_rng(23)
    n, k = 6000, 4
    nodes = rng.random((n, 2)) * [W, H]
    # 400 triples c, c + (6e-8, 5), c + (0, 5): d^2 = 25 + 1 ulp and 25, both roots round to 5.0, and
    # the node with the larger d^2 has the lower index
    c = np.floor(rng.random((400, 2)) * [W - 20, H - 20]) + 5
    nodes[0:400] = c
    nodes[400:800] = c + [6e-8, 5.0]
    nodes[800:1200] = c + [0.0, 5.0]
    d2a = ((nodes[400:800] - c) ** 2).sum(1)
    d2b = ((nodes[800:1200] - c) ** 2).sum(1)
    assert (d2a > d2b).any() and (np.sqrt(d2a) == np.sqrt(d2b)).all()
    tree = sg.NodeStore(2)
    tree.extend(nodes)
    want_nbr, want_vis = oracle_connect(om, nodes, k)
    nbr, vis = sg.prm_connect_knn(cc, tree, k)
    assert np.array_equal(nbr.cpu().numpy(), want_nbr)
    assert np.array_equal(vis.cpu().numpy(), want_vis)
    # sparse pocket: a dense cloud plus a few far-away nodes (their rectangles hold the whole cloud)
    nodes2 = rng.random((n, 2)) * [60, 60] + [100, 100]
    nodes2[:6] = [[5, 5], [W - 5, 5], [5, H - 5], [W - 5, H - 5], [W / 2, 5], [5, H / 2]]
    t2 = sg.NodeStore(2)
    t2.extend(nodes2)
    want_nbr, want_vis = oracle_connect(om, nodes2, k)
    nbr, vis = sg.prm_connect_knn(cc, t2, k)
    assert np.array_equal(nbr.cpu().numpy(), want_nbr)
    assert np.array_equal(vis.cpu().numpy(), want_vis)
    # a non-finite coordinate in the store: every row takes the exact path (NaN distances rank last)
    nodes3 = nodes.copy()
    nodes3[1234, 0] = np.nan
    nodes3[77, 1] = np.inf
    t3 = sg.NodeStore(2)
    t3.extend(nodes3)
    with np.errstate(invalid="ignore"):
        want_nbr, _ = oracle_connect(om, nodes3, k)
    nbr, vis = sg.prm_connect_knn(cc, t3, k)
    assert np.array_equal(nbr.cpu().numpy(), want_nbr)
    fin = np.isfinite(nodes3).all(1)
    rows_fin = fin & fin[np.where(want_nbr >= 0, want_nbr, 0)].all(1)
    wv = om.visible2d(nodes3[np.where(want_nbr >= 0, want_nbr, 0)[rows_fin].reshape(-1)],
                      np.repeat(nodes3[rows_fin], k, 0)).reshape(-1, k)
    assert np.array_equal(vis.cpu().numpy()[rows_fin], wv)
