"""CPU-only, world_size 2 over gloo: the multi-GPU partition / all-gather assembly
(sbp_env_b200/distributed.py).  The per-rank compute is the oracle here (tests may use
it as the checker); on the GPU box the same code runs cc.visible_batch over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import c_oracle as orc
        from _util import load_map
        from sbp_env_b200 import distributed as D

        free = load_map("maze1")
        W, H = free.shape
        om = orc.Map(free)
        rng = np.random.default_rng(0)                   # same inputs on every rank
        n = 20001
        P, Q = rng.random((n, 2)) * [W, H], rng.random((n, 2)) * [W, H]
        calls = []

        def visible_fn(p, q):
            calls.append(len(p))
            return om.visible2d(p.numpy(), q.numpy())

        full = om.visible2d(P, Q)
        tP, tQ = torch.from_numpy(P), torch.from_numpy(Q)
        got = D.sharded_visible_batch(visible_fn, tP, tQ)
        assert np.array_equal(got.numpy(), full)
        assert calls[-1] in (n // world, n // world + 1)  # only this rank's chunk was checked
        px = np.maximum(np.abs(np.trunc(P) - np.trunc(Q)).max(1), 0) + 1
        got = D.sharded_visible_batch(visible_fn, tP, tQ, costs=px)
        assert np.array_equal(got.numpy(), full)

        # PRM connect: rows sharded, neighbour blocks all-gathered
        nodes = P[:3001]
        k = 6

        def connect_fn(first, count):
            idx, _ = orc.knn(nodes, nodes[first:first + count], k)
            vis = om.visible2d(nodes[idx.reshape(-1)], np.repeat(nodes[first:first + count], k, 0))
            return torch.from_numpy(idx), torch.from_numpy(vis.reshape(count, k))

        nbr, vis = D.sharded_prm_connect(connect_fn, len(nodes))
        eidx, _ = orc.knn(nodes, nodes, k)
        assert np.array_equal(nbr.numpy(), eidx)
        evis = om.visible2d(nodes[eidx.reshape(-1)], np.repeat(nodes, k, 0)).reshape(-1, k)
        assert np.array_equal(vis.numpy(), evis)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
    for rank, msg in res:
        assert msg == "ok", f"rank {rank}: {msg}"
