"""-m gpu, needs >= 2 GPUs (skipped otherwise): the peer-memory adjacency all-gather
(sbp_adjacency_peer_scatter + symmetric-memory barrier) against the NCCL all-gather of the
same blocks, one process per GPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist

        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        from sbp_env_b200 import distributed as D

        kinds = set()
        for m, k, mc in ((5000, 10, True), (777, 3, True), (5000, 10, False)):   # 777 * 3 is not a multiple of 4
            ag = D.PeerAdjacencyGather(m, k, device=rank, multicast=mc)
            kinds.add(ag.kind)
            for step in range(5):                               # both buffers, reuse
                g = torch.Generator(device=dev)
                g.manual_seed(1000 * step + rank)
                nbr = torch.randint(-1, 1 << 20, (m, k), generator=g, device=dev, dtype=torch.int64)
                vis = torch.rand((m, k), generator=g, device=dev) < 0.7
                packed = ag.gather(nbr, vis)
                want_n, want_v = D.all_gather_adjacency(nbr, vis)
                got_n, got_v = D.unpack_adjacency(packed)
                assert torch.equal(got_n, want_n) and torch.equal(got_v, want_v), (m, k, step)
        # fused form: the visibility kernel of the connection step stores the packed adjacency
        # straight through the multicast mapping; only a barrier follows
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import sbp_env_b200 as sg
        from _util import load_map, map_png

        free = load_map("maze1")
        W, H = free.shape
        cc = sg.ImgCollisionChecker(map_png("maze1"), device=rank)
        rng = np.random.default_rng(5)
        pts = rng.random((6000, 2)) * [W, H]
        tree = sg.NodeStore(2, device=rank)
        tree.extend(pts)
        m, k = 1500, 7
        ag = D.PeerAdjacencyGather(m, k, device=rank)
        if ag.fused:
            kinds.add("fused into the visibility kernel")
            for step in range(4):
                Xq = torch.from_numpy(np.random.default_rng(100 * step + rank).random((m, 2)) * [W, H]).to(dev)
                h = step & 1
                nbr, vis = sg.prm_connect_knn(cc, tree, k, queries=Xq, packed_out=ag.target(h))
                packed = ag.complete(h)
                want_n, want_v = D.all_gather_adjacency(nbr, vis)
                got_n, got_v = D.unpack_adjacency(packed)
                assert torch.equal(got_n, want_n) and torch.equal(got_v, want_v), ("fused", step)
            # ... and with the rendezvous inside the kernel: no barrier launch at all
            kinds.add("fused + in-kernel rendezvous")
            ag2 = D.PeerAdjacencyGather(m, k, device=rank)
            for step in range(6):
                Xq = torch.from_numpy(np.random.default_rng(200 * step + rank).random((m, 2)) * [W, H]).to(dev)
                h = step & 1
                if rank == step % world:
                    torch.cuda._sleep(20_000_000)               # a late rank: the others must wait for it
                nbr, vis = sg.prm_connect_knn(cc, tree, k, queries=Xq, packed_out=ag2.target_with_rendezvous(h))
                packed = ag2.view(h).clone()                    # same stream: after the kernel returned
                cc.check_device_flags("rendezvous")
                want_n, want_v = D.all_gather_adjacency(nbr, vis)
                got_n, got_v = D.unpack_adjacency(packed)
                assert torch.equal(got_n, want_n) and torch.equal(got_v, want_v), ("rendezvous", step)
        # the roadmap connection of the stored rows, split over the ranks: every rank builds the cell
        # grid for its band of cell rows, answers its block of the cell-sorted rows and stores rows
        # and permutation slice through the multicast mapping (PeerAdjacencyBuffer)
        n2, k2 = 40013, 10
        free2 = load_map("intel_lab")
        cc2 = sg.ImgCollisionChecker(map_png("intel_lab"), device=rank)
        nodes2 = np.random.default_rng(77).random((n2, 2)) * list(free2.shape)
        nodes2[:300] = np.floor(nodes2[:300])
        tree2 = sg.NodeStore(2, device=rank)
        tree2.extend(nodes2)
        whole = torch.empty((n2, k2), dtype=torch.int32, device=dev)
        owhole = torch.empty((n2,), dtype=torch.int32, device=dev)
        sg.prm_connect_knn(cc2, tree2, k2, packed_out=whole, order_out=owhole, want_rows=False)
        buf = D.PeerAdjacencyBuffer(n2, k2, device=rank)
        if buf.fused:
            kinds.add("banded rows + permutation through the multicast mapping")
            for step in range(4):
                h = step & 1
                sg.prm_connect_knn(cc2, tree2, k2, part=(rank, world), packed_out=buf.target(h), want_rows=False,
                                   order_out=buf.order_target(h))
                packed = buf.complete(h)
                assert torch.equal(packed, whole) and torch.equal(buf.order_view(h), owhole), ("banded", step)
        # the same through plain device buffers + NCCL (fabrics without the mapping)
        mine = torch.full((n2, k2), -(1 << 31), dtype=torch.int32, device=dev)
        omine = torch.full((n2,), -1, dtype=torch.int32, device=dev)
        sg.prm_connect_knn(cc2, tree2, k2, part=(rank, world), packed_out=mine, want_rows=False, order_out=omine)
        lo, cnt = D.shard_range(n2, rank, world)
        assert bool((omine[:lo] == -1).all()) and bool((omine[lo + cnt:] == -1).all())
        dist.all_reduce(mine, op=dist.ReduceOp.MAX)
        dist.all_reduce(omine, op=dist.ReduceOp.MAX)
        assert torch.equal(mine, whole) and torch.equal(omine, owhole)
        dist.barrier()
        dist.destroy_process_group()
        print("exchange kinds exercised:", sorted(kinds), flush=True)
        q.put((rank, "ok"))
    except Exception as e:          # noqa: BLE001 - reported to the parent
        import traceback

        q.put((rank, "".join(traceback.format_exception(e))))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_peer_adjacency_gather_equals_nccl_all_gather():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in ps]
    for p in ps:
        p.join(30)
    assert all(r[1] == "ok" for r in res), res
