#!/usr/bin/env python
"""bench.py -- the hot path on BASELINE.json configs[4] (cfg5): PRM roadmap construction on the
32 768^2 synthetic maze, 1 000 000 free nodes, k = 10 nearest neighbours -> 10 000 000 candidate
edges checked through visible_batch, then 4 096 independent start / goal queries.

One step = one pass of the hot path over the whole roadmap: exact fp64 kNN of every node among
all nodes (the cell grid is rebuilt inside every step) + integer-Bresenham visibility of the
10 M candidate edges (map = 128 MiB bit-packed, L2 / HBM resident), the packed adjacency
(4 bytes per candidate) complete on every rank at the end of the step.  At N GPUs the rows are
split N ways (STRONG scaling: the total work is fixed) and every rank's block is written
through the NVSwitch multicast mapping by the connection kernel itself as the rows are
decided; the only other synchronisation is the barrier at the end of the step.

After the timed loop the query stage is timed separately (`query_stage`): CSR assembly of the
undirected roadmap from the packed adjacency, connection of the 8 192 query endpoints, connected
components, fewest-hop search; the queries are sharded round-robin over the ranks and the hop
counts all-gathered.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the UNMODIFIED reference (oracle/_ref: prmPlanner.nearest_neighbours +
ImgCollisionChecker.visible, all host cores via multiprocessing) on bounded samples of the same
workload.  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import workloads  # noqa: E402  (synthetic problem generators; harness code)

SCALE = float(os.environ.get("BENCH_SCALE", "1.0"))       # < 1: a smaller maze (debugging only)
K_NN = workloads.CFG5["k"]
K_CONNECT = 8                                              # nearest roadmap nodes tried per query endpoint
METRIC = "edge collision checks/sec (PRM candidate edges: exact kNN connect + visible_batch)"
UNIT = "edges/s"
# One launch of the dominant kernel on this workload under `ncu --set full` (N = 1, full size;
# profiles/r2_prm_knn_fused_full.txt): dram__bytes_read.sum + dram__bytes_write.sum, issue-slot
# utilisation, active lanes per warp instruction.
NCU_FUSED = {"kernel": "k_prm_knn<10, true>", "dram_bytes": 119872256, "issue_active_pct": 59.5,
             "lanes_per_instruction": 24.6, "warp_instructions": 210052932, "warps_active_pct": 47.4,
             "alu_pipe_pct": 64.7, "l1_hit_pct": 61.6, "l2_hit_pct": 73.8,
             "source": "profiles/r2_prm_knn_fused_full.txt"}
NCU_BUILD = {"kernel": "k_prm_grid_build", "kernel_us": 68.1, "dram_bytes": 16486656, "issue_active_pct": 25.3,
             "source": "profiles/r2_prm_grid_build_full.txt"}

def problem():
    mz, nodes, S, T = workloads.cfg5_problem(SCALE)
    return mz, nodes, S, T


def config_dict(mz, n_nodes, n_queries):
    """The SAME dict from both arms (the driver compares them)."""
    wl = (f"cfg5: PRM roadmap on a {mz.W}x{mz.H} seeded recursive-backtracker maze "
          f"({mz.cells}x{mz.cells} cells of {mz.pitch} px, walls {mz.wall} px), {n_nodes} free nodes, k={K_NN} -> "
          f"{n_nodes * K_NN} candidate edges per step (exact kNN of every node among all nodes + Bresenham "
          f"visible_batch, adjacency assembled on every rank); {n_queries} start/goal queries timed separately")
    return {"workload": wl, "map": f"synthetic maze {mz.W}x{mz.H}, seed {mz.seed}, bit-packed {mz.W * mz.H // 8} bytes",
            "nodes": int(n_nodes), "k": K_NN, "edges_per_step": int(n_nodes * K_NN), "queries": int(n_queries),
            "l2": "flushed (256 MiB write) between timed steps; the 128 MiB map exceeds what L2 keeps anyway",
            "scale": SCALE}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, ~2 ms period;
    falls back to one `nvidia-smi` query when pynvml is unavailable)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, gpu_index=0):
        self.gpu, self.samples, self.stop_flag, self.t, self.h = gpu_index, [], False, None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; map through it when set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        except Exception:
            self.nv = None

    def start(self):
        if self.nv is None:
            return

        def pump():
            nv = self.nv
            while not self.stop_flag:
                try:
                    clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                        nv, "nvmlDeviceGetCurrentClocksEventReasons") else \
                        nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append((time.time(), clk, rs))
                except Exception:
                    pass
                time.sleep(0.002)

        self.t = threading.Thread(target=pump, daemon=True)
        self.t.start()

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.t is not None:
            self.t.join(1.0)
        if self.nv is None or not self.samples:
            try:
                o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm",
                                    "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                   capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(o[0]), "sm_max_mhz": float(o[1]), "reasons": [],
                        "samples": 0, "note": "single nvidia-smi query after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"]}
        nv = self.nv
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        inside = [(c, r) for ts, c, r in self.samples if t0 <= ts <= t1] or [(c, r) for _, c, r in self.samples[-3:]]
        reasons = set()
        for _, r in inside:
            for bit, name in self.REASONS.items():
                if r & bit:
                    reasons.add(name)
        return {"sm_mhz": float(np.median([c for c, _ in inside])), "sm_max_mhz": float(mx),
                "reasons": sorted(reasons), "samples": len(inside)}


# ------------------------------------------------------------------ ours -----
def run_ours(args):
    import torch
    import torch.distributed as dist

    import sbp_env_b200 as sg

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- the problem: identical on every rank (seeded host generators) ----------------
    mz, nodes_h, S_h, T_h = problem()
    n, k = len(nodes_h), K_NN
    nq = len(S_h)
    free = mz.raster_device(dev)
    cc = sg.ImgCollisionChecker.from_free_mask(free, device=local)
    del free
    ctx = cc.device_map.ctx
    nodes_d = torch.from_numpy(nodes_h).to(dev)
    tree = sg.NodeStore(2, capacity=n, device=local)
    tree.extend(nodes_d)
    first, count = sg.distributed.shard_range(n, rank, world)  # rows of this rank: a run of the cell-sorted order
    edges_total = n * k
    SENTINEL = -(1 << 31)

    # N > 1: the rows are split into `world` contiguous runs of the cell-sorted order (spatially
    # compact shards); every rank's connection kernel stores the packed adjacency of its rows
    # through the NVSwitch multicast mapping of a symmetric [n][k] buffer as the edges are decided,
    # so every rank's copy is complete after the barrier.  Without the mapping (or without torch's
    # symmetric memory) the ranks' rows are merged by an NCCL all-reduce(MAX) over a sentinel.
    peer, exchange = None, "none (N = 1): packed adjacency written to a local buffer"
    if world > 1:
        try:
            peer = sg.distributed.PeerAdjacencyBuffer(n, k, device=local)
        except Exception as e:      # noqa: BLE001
            exchange = f"NCCL all-reduce(MAX) of the packed rows (peer memory unavailable: {type(e).__name__})"
    fused = peer is not None and peer.fused
    if fused:
        exchange = ("fused: the connection kernel stores the packed adjacency of its rows through the NVSwitch "
                    "multicast mapping as edges are decided; symmetric-memory barrier")
    elif world > 1 and peer is not None:
        exchange = "NCCL all-reduce(MAX) of the packed rows (no multicast mapping on this fabric)"
    # The adjacency is kept in CELL-SORTED row order: row p belongs to node order[p] (the permutation
    # of the kNN's counting sort, a function of the poses only: identical on every rank), so that a
    # rank's rows are ONE contiguous block which the walk kernel stores with coalesced lines --
    # through the multicast mapping at N > 1, into pinned host memory in the e2e step.
    packed_local = torch.full((n, k), SENTINEL, dtype=torch.int32, device=dev)
    order = torch.full((n,), -1, dtype=torch.int32, device=dev)
    packed_full = {}                                  # filled by every step: int32 [n][k], cell-sorted rows
    peer_step = [0]
    part = (rank, world)

    order_local = order                               # this rank's slice (N > 1) / the whole permutation (N = 1)
    order_full = {"o": order}

    def nccl_merge():
        merged = packed_local.clone()
        dist.all_reduce(merged, op=dist.ReduceOp.MAX)  # every row is written by exactly one rank
        om = order_local.clone()
        dist.all_reduce(om, op=dist.ReduceOp.MAX)
        order_full["o"] = om
        return merged

    def eager_step(mark=None):
        if fused:
            h = peer_step[0] & 1
            peer_step[0] += 1
            sg.prm_connect_knn(cc, tree, k, part=part, mark=mark, packed_out=peer.target(h), want_rows=False,
                               order_out=peer.order_target(h))
            packed_full["p"] = peer.complete(h)
            order_full["o"] = peer.order_view(h)
        else:
            if world > 1:
                order_local.fill_(-1)
            sg.prm_connect_knn(cc, tree, k, part=part, mark=mark, packed_out=packed_local, want_rows=False,
                               order_out=order_local)
            packed_full["p"] = packed_local if world == 1 else nccl_merge()
        if mark:
            mark("exchange")

    # the timed loop replays the step from a CUDA graph captured by the library (one per buffer half
    # when the exchange is fused; the barrier is captured too where torch allows it)
    barrier_in_graph = False
    if fused:
        try:
            gsteps = [sg.PRMConnectGraph(cc, tree, k, part=part, packed_out=peer.target(h), want_rows=False,
                                         order_out=peer.order_target(h), after=(lambda h=h: peer.complete(h)))
                      for h in (0, 1)]
            barrier_in_graph = True
        except Exception as e:      # noqa: BLE001
            print(f"[bench] barrier not capturable ({type(e).__name__}: {e}); eager barrier", file=sys.stderr)
        ok = torch.tensor([1 if barrier_in_graph else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        barrier_in_graph = bool(ok.item())
        if not barrier_in_graph:
            gsteps = [sg.PRMConnectGraph(cc, tree, k, part=part, packed_out=peer.target(h), want_rows=False,
                                         order_out=peer.order_target(h)) for h in (0, 1)]
        exchange += " (captured in the step's graph)" if barrier_in_graph else " (separate launch)"
        dist.barrier()
    else:
        gstep = sg.PRMConnectGraph(cc, tree, k, part=part, packed_out=packed_local, want_rows=False,
                                   order_out=order_local)

    def step():
        if fused:
            h = peer_step[0] & 1
            peer_step[0] += 1
            gsteps[h].replay()
            packed_full["p"] = peer.view(h) if barrier_in_graph else peer.complete(h)
            order_full["o"] = peer.order_view(h)
        else:
            gstep.replay()
            packed_full["p"] = packed_local if world == 1 else nccl_merge()

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def align_ranks():
        # N > 1: the untimed L2 flush does not take the same time on every rank; without this the
        # timed step of an early rank would include waiting for a rank that is still flushing
        if world > 1:
            if peer is not None:
                peer.hdl.barrier(channel=2)
            else:
                dist.barrier()

    def timed_loop(fn, steps):
        a = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        b = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for s in range(steps):
            flush.zero_()                      # evict L2 between timed iterations (not timed)
            align_ranks()                      # (not timed) all ranks start the step together
            a[s].record()
            fn()
            b[s].record()
        torch.cuda.synchronize()
        ms = torch.tensor([sum(a[s].elapsed_time(b[s]) for s in range(steps))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    total_ms = timed_loop(step, args.steps)
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- phase breakdown and the dominant kernel's own duration: the same step launched eagerly
    # behind a spin kernel (the host runs ahead, so the events measure the GPU); the connection
    # kernel (kNN + visibility fused) is bracketed by its own CUDA events on the launching stream
    names = ("knn", "edges", "visible", "exchange")     # marks of prm_connect_knn + the exchange
    psteps = min(args.steps, 10)
    for _ in range(2):
        eager_step()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    ctx.tune("time_kernels", 3)
    p_start = [torch.cuda.Event(enable_timing=True) for _ in range(psteps)]
    phase_ev = [[torch.cuda.Event(enable_timing=True) for _ in names] for _ in range(psteps)]
    for s in range(psteps):
        flush.zero_()
        torch.cuda._sleep(6_000_000)       # ~3 ms head start for the host
        align_ranks()
        p_start[s].record()
        it = iter(phase_ev[s])
        eager_step(mark=lambda _name, it=it: next(it).record())
    torch.cuda.synchronize()
    ctx.tune("time_kernels", 0)
    vis_ms_total, vis_n = ctx.kernel_time()
    vis_ms = vis_ms_total / max(1, vis_n)
    launches_per_step = (ctx.launches - launches0) // psteps
    ph = {}
    for i, nme in enumerate(names):
        prev = [p_start[s] if i == 0 else phase_ev[s][i - 1] for s in range(psteps)]
        ph[nme] = float(np.mean([prev[s].elapsed_time(phase_ev[s][i]) for s in range(psteps)]))
    # (the one-call connection marks its three phases together: their sum is the call)
    ph = {"connect_call": ph["knn"] + ph["edges"] + ph["visible"], "exchange": ph["exchange"]}
    # the cell-grid kNN kernel the same way
    ctx.tune("time_kernels", 2)
    for s in range(3):
        flush.zero_()
        eager_step()
    torch.cuda.synchronize()
    ctx.tune("time_kernels", 0)
    knn_ms_total, knn_n = ctx.kernel_time()
    knn_kernel_ms = knn_ms_total / max(1, knn_n)

    # ---- parity of the timed configuration (outside the timed region) ----------------------
    step()
    torch.cuda.synchronize()
    replayed = packed_full["p"].clone()
    eager_step()
    torch.cuda.synchronize()
    full = packed_full["p"]
    assert tuple(full.shape) == (n, k)
    assert torch.equal(replayed, full), "graph replay differs from the eager step"
    order = order_full["o"].clone()                   # the whole permutation, as assembled by the step
    assert torch.equal(torch.sort(order.long()).values, torch.arange(n, device=dev)), "row order is not a permutation"
    # this rank's own rows, computed into a private sentinel buffer: exactly its block of positions
    own = torch.full((n, k), SENTINEL, dtype=torch.int32, device=dev)
    order2 = torch.full((n,), -1, dtype=torch.int32, device=dev)
    sg.prm_connect_knn(cc, tree, k, part=part, packed_out=own, want_rows=False, order_out=order2)
    assert torch.equal(order2[first:first + count], order[first:first + count])
    if world > 1:
        assert bool((order2[:first] == -1).all()) and bool((order2[first + count:] == -1).all())
    mine_rows = own[:, 0] != SENTINEL
    assert int(mine_rows.sum().item()) == count and bool(mine_rows[first:first + count].all()), \
        "this rank did not write exactly its block of the rows"
    assert torch.equal(own[mine_rows], full[mine_rows]), "assembled adjacency differs from the local rows"
    adjacency_check = "single rank"
    if world > 1:
        # the WHOLE assembled adjacency and permutation against an NCCL merge of the ranks' rows, on
        # every rank; and against the single-rank build of the whole roadmap on rank 0's device
        want = own.clone()
        dist.all_reduce(want, op=dist.ReduceOp.MAX)
        owant = order2.clone()
        dist.all_reduce(owant, op=dist.ReduceOp.MAX)
        owners = mine_rows.to(torch.int32)
        dist.all_reduce(owners)
        whole = torch.empty_like(own)
        owhole = torch.empty_like(order2)
        sg.prm_connect_knn(cc, tree, k, packed_out=whole, want_rows=False, order_out=owhole)
        same = torch.tensor([1 if (torch.equal(want, full) and bool((owners == 1).all().item())
                                   and torch.equal(owant, order) and torch.equal(whole, full)
                                   and torch.equal(owhole, order)) else 0], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        assert bool(same.item()), "adjacency assembled through peer memory differs from the NCCL merge"
        adjacency_check = (f"every row written by exactly one rank; rows and permutation equal to the NCCL merge of the "
                           f"{world} ranks' blocks and to the single-rank build, on every rank")
        del whole, owhole, want, owant
    # node-order view of the adjacency for the checks below
    full_node = torch.empty_like(full)
    full_node[order.long()] = full
    # the one-call path against the generic path (kNN call + general visibility kernel) on a
    # slice of the rows, bit for bit
    sl = (rank * 20_000 % max(1, n - 20_000), min(20_000, n))
    nbr_g, vis_g = sg.prm_connect_knn(cc, tree, k, rows=sl)
    gn, gv = sg.distributed.unpack_adjacency(full_node[sl[0]:sl[0] + sl[1]])
    assert torch.equal(gn, nbr_g) and torch.equal(gv, vis_g), "one-call connection differs from the generic path"
    n_vis = int((full & 1).sum().item())
    # size-independent properties at full size (this rank checks every world-th row): the 2-D
    # checker is direction symmetric, and a visible edge has feasible endpoints
    sel = torch.arange(rank, n, world, device=dev)
    nbr_s, vis_s = sg.distributed.unpack_adjacency(full_node[sel])
    P = nodes_d[nbr_s.reshape(-1).clamp(min=0)].contiguous()
    Q = nodes_d[sel.repeat_interleave(k)].contiguous()
    v_rev = cc.visible_batch(Q, P).reshape(-1, k) & (nbr_s >= 0)
    assert torch.equal(v_rev, vis_s), "visible(p, q) != visible(q, p)"
    assert bool((~vis_s.reshape(-1) | (cc.feasible_batch(P) & cc.feasible_batch(Q))).all().item())
    px_full = float((torch.maximum((P[:, 0].trunc() - Q[:, 0].trunc()).abs(),
                                   (P[:, 1].trunc() - Q[:, 1].trunc()).abs()) + 1).sum().item())
    del P, Q, v_rev, nbr_s, vis_s, own, full_node
    px_t = torch.tensor([px_full], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(px_t)
    px_total = float(px_t.item())

    # ---- query stage: undirected CSR + endpoint connection + components + fewest-hop search ----
    road = sg.Roadmap(n, device=local)
    SG_all = np.concatenate([S_h, T_h])
    mine = np.arange(rank, nq, world)                                  # round-robin shard of the queries
    SG_mine = torch.from_numpy(np.concatenate([S_h[mine], T_h[mine]])).to(dev)
    hops_all = torch.full((nq,), -2, dtype=torch.int32, device=dev)
    q_ev = {}

    def query_stage(record=False):
        def mk(name):
            if record:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                q_ev[name] = e
        mk("start")
        road.build_from_packed(packed_full["p"], undirected=True, row_ids=order_full["o"])
        mk("csr")
        ends = sg.connect_to_roadmap(cc, tree, SG_mine, K_CONNECT)
        mk("connect")
        m_q = len(mine)
        hops = road.shortest_paths(ends[:m_q].contiguous(), ends[m_q:].contiguous())
        mk("search")
        if world > 1:
            pad = (nq + world - 1) // world
            buf = torch.full((pad,), -2, dtype=torch.int32, device=dev)
            buf[:m_q] = hops
            allh = torch.empty((world * pad,), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(allh, buf)
            for r in range(world):
                cnt = len(range(r, nq, world))
                hops_all[r::world] = allh[r * pad: r * pad + cnt]
        else:
            hops_all.copy_(hops)
        mk("gather")
        return ends

    query_stage()
    torch.cuda.synchronize()
    q_reps = 3
    q_ms = timed_loop(query_stage, q_reps) / q_reps
    ends = query_stage(record=True)
    torch.cuda.synchronize()
    q_ph = {b: q_ev[a].elapsed_time(q_ev[b]) for a, b in zip(("start", "csr", "connect", "search"),
                                                             ("csr", "connect", "search", "gather"))}
    hh = hops_all.cpu().numpy()
    assert (hh >= -1).all()
    _, n_comp = road.components()
    # parity spot check of the search at full size (rank 0): scipy's BFS on the same CSR for a
    # few of this rank's queries, reachable ones first
    query_check = None
    if rank == 0:
        query_check = check_queries(road, ends, hh[mine], len(mine))

    # ---- end to end through the public API with host buffers -------------------------------
    pinned = torch.from_numpy(nodes_h).pin_memory()
    code_h = torch.full((n, k), SENTINEL, dtype=torch.int32).pin_memory()   # [n][k] cell-sorted rows; this rank fills its block
    order_h = torch.full((n,), -1, dtype=torch.int32).pin_memory()
    order_e = torch.empty_like(order)
    code_dev = torch.full((n, k), SENTINEL, dtype=torch.int32, device=dev)
    x_dev = torch.empty((n, 2), dtype=torch.float64, device=dev)
    tree2 = sg.NodeStore(2, capacity=n, device=local)
    # H2D: the node-store append kernel reads the pinned host rows itself (coalesced zero-copy reads;
    # BENCH_E2E_H2D_COPY=1: explicit copy first).  D2H: the call is handed the pinned host buffers; the
    # library computes this rank's block of the packed adjacency in slices and copies every slice out
    # (DMA engine, second stream) while the next ones are computed, the row ids as soon as the grid is
    # built (BENCH_E2E_D2H_COPY=1: device buffers + two copies after the call).
    h2d_copy = os.environ.get("BENCH_E2E_H2D_COPY", "0") == "1"
    d2h_copy = os.environ.get("BENCH_E2E_D2H_COPY", "0") == "1"

    def e2e_body():
        if h2d_copy:
            x_dev.copy_(pinned, non_blocking=True)
        tree2.truncate(0)
        tree2.extend(x_dev if h2d_copy else pinned)
        sg.prm_connect_knn(cc, tree2, k, part=part, want_rows=False,
                           order_out=order_e if d2h_copy else order_h.data_ptr(),
                           packed_out=code_dev if d2h_copy else (code_h.data_ptr(), False))
        if d2h_copy:
            code_h[first:first + count].copy_(code_dev[first:first + count], non_blocking=True)
            order_h[first:first + count].copy_(order_e[first:first + count], non_blocking=True)

    e2e_graph = sg.CapturedStep(e2e_body, device=local)
    for _ in range(max(3, args.warmup)):
        e2e_graph.replay()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_ms = timed_loop(e2e_graph.replay, args.steps)
    assert torch.equal(order_h[first:first + count], order[first:first + count].cpu()), \
        "e2e row permutation differs from the device-resident run"
    assert torch.equal(code_h[first:first + count], full[first:first + count].cpu()) and \
        bool((code_h[:first] == SENTINEL).all()) and bool((code_h[first + count:] == SENTINEL).all()), \
        "e2e result differs from the device-resident run"

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    # Dominant kernel: k_prm_knn<10, true> -- the row's exact k nearest from the cell grid AND the
    # visibility of its k candidate edges (free bounding boxes from the run planes, probes, tile walk),
    # output rows written from shared memory.  ALGORITHMIC bytes per launch:
    # visibility 33 B per edge (two fp64 endpoints in, one byte out) + 0.125 B (one map bit) per
    # interpolant of the full Bresenham line; kNN every input byte once (cell-sorted poses 16 B + ids
    # 4 B per node, cell starts 4 B per cell) + 9 B per row (bound and flag for the exact path).
    edges_rank = count * k
    vis_bytes = 33.0 * edges_rank + 0.125 * px_total / world
    knn_bytes = 20.0 * n + 4.0 * (n / 1.5) + 9.0 * count
    alg_bytes = vis_bytes + knn_bytes
    fused_ms = knn_kernel_ms                                    # (timers 2 and 3 bracket the same kernel)
    vis_s = fused_ms / 1e3
    step_ms = total_ms / args.steps
    prof = NCU_FUSED if world == 1 and SCALE == 1.0 else None
    roof = {"kernel": f"k_prm_knn<{k}, true>: exact k nearest of {count} rows among {n} nodes from the cell grid + "
                      f"visibility of their {edges_rank} candidate edges ({px_total / world:.4g} interpolants) on the "
                      f"128 MiB map, rows written from shared memory ({100 * fused_ms / step_ms:.0f}% of the step)",
            "bound": "hbm", "achieved": alg_bytes / vis_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_bytes / vis_s / 1e9 / hbm_peak, "traffic": prof["dram_bytes"] if prof else None,
            "peak_source": peak_src, "kernel_ms": fused_ms, "kernel_launches_timed": knn_n,
            "algorithmic_bytes_per_launch": alg_bytes,
            "algorithmic_bytes": {"visibility": vis_bytes, "knn": knn_bytes},
            "edges_per_s_kernel": edges_rank / vis_s, "interpolants_per_s_kernel": px_total / world / vis_s,
            "rows_per_s_kernel": count / vis_s, "effective_pairs_per_s": float(count) * n / vis_s,
            "ncu": prof,
            # what does bound it: warp instructions issued (ncu count of one launch) against 4 issue slots
            # per SM and cycle at the clock sampled during the timed loop
            "issue": ({"warp_instructions_per_s": prof["warp_instructions"] / vis_s,
                       "peak": 148 * 4 * float(clocks["sm_mhz"]) * 1e6 if clocks and clocks.get("sm_mhz") else None,
                       "frac": (prof["warp_instructions"] / vis_s) / (148 * 4 * float(clocks["sm_mhz"]) * 1e6)
                       if clocks and clocks.get("sm_mhz") else None} if prof else None),
            "note": "not an HBM stream: the kernel gathers from L1 / L2 resident arrays (cell-sorted poses 16 MB, "
                    "free-run planes 16 MiB, tile codes 8 MiB) and is bound by the instructions it issues (list "
                    "insertions in lock-step over the warp; ncu: issue slots, lanes per instruction); the HBM figure "
                    "says how far the algorithmic bytes are from mattering"}
    roof_build = {"kernel": "k_prm_grid_build: bounding box, histogram, scan, placement, per-cell order + cell-sorted "
                            "poses in one cooperative launch (the second kernel of the step)",
                  "ncu": NCU_BUILD if SCALE == 1.0 else None,
                  "note": "five grid-wide barriers and dependent-load latency; not timed inside bench.py (its time is "
                          "phases_ms.connect_call minus the kernels timed here minus the short launches)"}

    call_ms = ph["connect_call"]                               # grid build + fused kernel + the exact path's launches
    sub = {"knn_queries_per_s_per_gpu": count / (call_ms / 1e3),
           "connect_call_ms": call_ms,
           "fused_kernel_ms": fused_ms, "fused_kernel_ms_second_timer": vis_ms,
           "knn_effective_pairs_per_s_per_gpu": float(count) * n / (call_ms / 1e3),
           "visible_batch_edges_per_s_per_gpu": edges_rank / (call_ms / 1e3),
           "interpolant_checks_per_s_per_gpu": px_total / world / (call_ms / 1e3),
           "prm_build_ms": total_ms / args.steps,
           "roadmap_edges_visible": n_vis,
           "build_plus_queries_ms": total_ms / args.steps + q_ms}
    queries = {"ms": q_ms, "queries_per_s": nq / (q_ms / 1e3), "phases_ms": q_ph, "queries": nq,
               "k_connect": K_CONNECT, "sharding": f"round-robin over {world} rank(s); hop counts all-gathered",
               "graph": "undirected (both directions of every visible candidate, duplicates removed)",
               "arcs": road.n_edges, "components": n_comp,
               "connected_endpoints_fraction": float(((ends >= 0).float().mean()).item()),
               "reachable_fraction": float((hh >= 0).mean()),
               "mean_hops_reachable": float(hh[hh >= 0].mean()) if (hh >= 0).any() else None,
               "max_hops": int(hh.max()), "parity": query_check,
               "note": "a k = 10 roadmap on 112-px corridors breaks wherever a corridor stretch of ~90 px holds no "
                       "sample (about 40 places in this maze), so only a few percent of the random start/goal "
                       "pairs are connected; the component labels settle the others without a search"}

    # the sequential planners' inner loop on the same 1M nodes taken as a tree:
    # nearest + the one-launch extend step (radius 64: ~50 candidates per sample), host to host
    rrt = rrt_extend_line(sg, cc, tree, nodes_h) if world == 1 and os.environ.get("BENCH_RRT", "1") == "1" else None
    # cfg2 (BASELINE configs[1], the round-1 headline) as a second line, N = 1 only
    cfg2 = cfg2_line(sg, dev, local, flush) if world == 1 and os.environ.get("BENCH_CFG2", "1") == "1" else None
    # CPU baseline: the unmodified reference on this box's host cores, in a child process (fork
    # pool; this process holds a CUDA context), bounded sample
    cpu = None
    if world == 1 and os.environ.get("BENCH_CPU", "1") == "1":
        cpu = cpu_baseline_subprocess(budget_s=14.0)

    out = {
        "metric": METRIC, "value": edges_total * args.steps / (total_ms / 1e3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(mz, n, nq),
        "clocks": clocks,
        "e2e": {"value": edges_total * args.steps / (e2e_ms / 1e3), "unit": UNIT,
                "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(pinned.numel() * 8),
                "d2h_bytes_per_step": int(count * k * 4 + count * 4),
                "result": "packed adjacency int32 [rows of this rank, cell-sorted][k] = (neighbour + 1) * 2 + visible, "
                          "and their node indices int32 [rows of this rank]",
                "d2h": "copy from a device buffer" if d2h_copy else
                       "the call writes the pinned host buffers: per-slice DMA copies of the rows overlapping the next slices' kernels, row ids copied beside the kNN",
                "h2d": "copy into a device buffer" if h2d_copy else
                       "read from the pinned host buffer by the node-store append kernel (zero-copy)"},
        "gpu_launches": int(launches_per_step * args.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "launch": "the step (and the e2e step) is replayed from a CUDA graph captured by the library "
                  "(PRMConnectGraph / CapturedStep); phases_ms and roofline.kernel_ms come from the same step "
                  "launched eagerly",
        "multi_gpu": {"rows_per_rank": count, "exchange": exchange, "adjacency_check": adjacency_check},
        "roofline": roof,
        "roofline_build": roof_build,
        "measured_gather_peaks": {"l2_random_sector_gathers_per_s_64MiB": microbench(ctx, 1, 64 << 20, 512),
                                  "hbm_random_sector_gathers_per_s_8GiB": microbench(ctx, 1, 8 << 30, 512)},
        "cpu_baseline": cpu,
        "phases_ms": ph,
        "submetrics": sub,
        "query_stage": queries,
        "rrt_extend": rrt,
        "cfg2": cfg2,
    }
    print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def microbench(ctx, which, nbytes, iters):
    import ctypes as C

    from sbp_env_b200 import _lib

    ms, units = C.c_double(), C.c_double()
    _lib.check(_lib.load().sbp_microbench(ctx.handle, which, nbytes, iters, C.byref(ms), C.byref(units)))
    return units.value / (ms.value / 1e3)


def check_queries(road, ends, hops_mine, m_q, n_check=24):
    """scipy's BFS (csgraph, unweighted) on the same CSR for a few queries of rank 0."""
    try:
        import scipy.sparse as sp
        from scipy.sparse.csgraph import breadth_first_order
    except Exception as e:      # noqa: BLE001
        return f"scipy unavailable ({type(e).__name__})"
    rp, col = road.csr()
    rp, col = rp.cpu().numpy(), col.cpu().numpy()
    nn = len(rp) - 1
    A = sp.csr_matrix((np.ones(len(col), np.int8), col, rp), shape=(nn, nn))
    e = ends.cpu().numpy()
    s_idx, t_idx = e[:m_q], e[m_q:]
    order = np.argsort(-(hops_mine >= 0).astype(np.int64), kind="stable")[:n_check]   # reachable ones first
    ok, n_reach = True, 0
    for q in order:
        s, t = int(s_idx[q]), int(t_idx[q])
        if s < 0 or t < 0:
            ok &= hops_mine[q] == -1
            continue
        _, pred = breadth_first_order(A, s, directed=True, return_predecessors=True)
        want, v = 0, t
        while v != s:                       # depth of t in the BFS tree = fewest hops
            v = int(pred[v])
            if v < 0:
                want = -1
                break
            want += 1
        ok &= int(hops_mine[q]) == want
        n_reach += want >= 0
    assert ok, "hop counts differ from scipy's BFS on the same graph"
    return f"hop counts of {len(order)} queries ({n_reach} reachable) equal scipy.sparse.csgraph BFS on the same CSR"


def rrt_extend_line(sg, cc, tree, nodes_h, m=1500, radius=64.0):
    """Per-sample latency of the RRT* inner loop on a 1M-node tree: NodeStore.nearest, then
    DeviceTree.extend (choose parent + rewire decisions in one launch; nothing appended, so every
    sample sees the same tree), against the two-call host path of round 1 (near_visible + host
    arithmetic)."""
    from sbp_env_b200 import planner_ops as ops

    n = len(nodes_h)
    rng = np.random.default_rng(8)
    costs = np.linalg.norm(nodes_h - nodes_h[0], axis=1)
    dt = sg.DeviceTree.from_arrays(cc, tree, nodes_h, costs, np.zeros(n, np.int32))
    X = nodes_h[rng.integers(0, n, m)] + rng.standard_normal((m, 2)) * 3.0
    X = X[cc.feasible_batch(X)]
    for x in X[:20]:
        dt.extend(x, tree.nearest(x), radius, append=False)
    t0 = time.perf_counter()
    for x in X:
        tree.nearest(x)
    tn = time.perf_counter() - t0
    nns = [tree.nearest(x) for x in X]
    t1 = time.perf_counter()
    for x, nn in zip(X, nns):
        dt.extend(x, nn, radius, append=False)
    t2 = time.perf_counter()
    v0 = cc.stats.visible_cnt
    for x, nn in zip(X[:200], nns):
        ops.choose_least_cost_parent(cc, tree, dt.costs, x, nn, radius, poses=nodes_h)
    t3 = time.perf_counter()
    cands = (cc.stats.visible_cnt - v0) / 200
    return {"tree_nodes": n, "samples": int(len(X)), "radius": radius, "mean_candidates": float(cands),
            "nearest_us": 1e6 * tn / len(X), "extend_us": 1e6 * (t2 - t1) / len(X),
            "host_path_choose_parent_us": 1e6 * (t3 - t2) / 200,
            "ambiguous_samples_settled_on_host": int(dt.ambiguous)}


# ------------------------------------------------------------ cfg2 line --------
def cfg2_line(sg, dev, local, flush):
    """BASELINE.json configs[1] (the round-1 headline): PRM connection on intel_lab, 50 000 free
    samples, k = 10 -> 500 000 candidate edges, device-resident, CUDA-graph replay."""
    import torch

    from _util import load_map

    free = load_map("intel_lab")
    W, H = free.shape
    cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8), device=local)
    rng = np.random.default_rng(0)
    pts, have = [], 0
    while have < 50_000:
        c = rng.random((100_000, 2)) * [W, H]
        c = c[cc.feasible_batch(c)]
        pts.append(c)
        have += len(c)
    nodes = np.ascontiguousarray(np.concatenate(pts)[:50_000])
    tree = sg.NodeStore(2, capacity=50_000, device=local)
    tree.extend(nodes)
    g = sg.PRMConnectGraph(cc, tree, 10)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    launches0 = tree.ctx.launches
    sg.prm_connect_knn(cc, tree, 10)
    launches = int(tree.ctx.launches - launches0)
    micro = micro_edges(sg, cc, W, H, nodes, dev, flush)
    return {"micro": micro, "workload": "cfg2: PRM connection on intel_lab 579x581, 50000 free samples, k=10 -> 500000 candidate edges",
            "ms_per_step": ms, "edges_per_s": 500_000 / (ms / 1e3), "gpu_launches_per_step": launches,
            "roadmap_edges_visible": int(g.vis.sum().item())}


def micro_edges(sg, cc, W, H, nodes, dev, flush):
    """Uniform / length-bucketed random edge batches and point feasibility on intel_lab (map
    resident in shared memory), SURVEY.md 8(d) "micro"."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(4)
    scale = torch.tensor([W, H], device=dev, dtype=torch.float64)
    hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

    def time_visible(P, Q, iters=5):
        cc.visible_batch(P, Q)
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = cc.visible_batch(P, Q)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        px = (torch.maximum((P[:, 0].trunc() - Q[:, 0].trunc()).abs(),
                            (P[:, 1].trunc() - Q[:, 1].trunc()).abs()) + 1).sum().item()
        t = float(np.median(ts)) / 1e3
        nn = P.shape[0]
        ab = 33.0 * nn + 0.125 * px
        return {"edges": nn, "ms": t * 1e3, "interpolants_full_line": px, "interpolants_per_s": px / t,
                "edges_per_s": nn / t, "visible_fraction": float(r.float().mean().item()),
                "algorithmic_GBps": ab / t / 1e9, "hbm_frac": ab / t / 1e9 / hbm}

    out = {}
    n_u = 1 << 24
    P = torch.rand((n_u, 2), generator=g, device=dev, dtype=torch.float64) * scale
    Q = torch.rand((n_u, 2), generator=g, device=dev, dtype=torch.float64) * scale
    out["uniform_2^24_edges"] = time_visible(P, Q)
    ts = []
    for _ in range(5):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cc.feasible_batch(P)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    tf = float(np.median(ts)) / 1e3
    out["feasible_2^24_points"] = {"points": n_u, "ms": tf * 1e3, "points_per_s": n_u / tf,
                                   "algorithmic_GBps": 17.0 * n_u / tf / 1e9, "hbm_frac": 17.0 * n_u / tf / 1e9 / hbm}
    fr = torch.from_numpy(nodes).to(dev)
    for L in (8, 32, 128):
        n_b = 1 << 22
        base = fr[torch.randint(0, len(nodes), (n_b,), generator=g, device=dev)]
        ang = torch.rand(n_b, generator=g, device=dev, dtype=torch.float64) * (2 * np.pi)
        out[f"free_start_len{L}_2^22_edges"] = time_visible(base, base + torch.stack([ang.cos(), ang.sin()], 1) * L)
    return out


# ------------------------------------------------------------ CPU legs --------
_REF = {}


def _ref_worker_init(img, poses):
    """Runs in every pool worker (fork): the unmodified reference, its checker around the
    pre-binarised uint8 grid (the reference's own __init__ would build an 8 GiB float64 image,
    collisionChecker.py:101-108), and the node array."""
    import ref_shims

    ref_shims.import_reference()
    from sbp_env.collisionChecker import CollisionChecker, ImgCollisionChecker
    from sbp_env.planners import prmPlanner

    cc = ImgCollisionChecker.__new__(ImgCollisionChecker)
    CollisionChecker.__init__(cc)
    cc._img = img
    _REF.update(cc=cc, nn=prmPlanner.nearest_neighbours, poses=poses, nodes=range(len(poses)))


def _ref_rows(job):
    """The reference's own code on a few rows: prmPlanner.nearest_neighbours (prmPlanner.py:28-44)
    with the radius that makes it return the row's k nearest, `if m_g is v: continue`, and
    ImgCollisionChecker.visible(m_g.pos, v.pos) (prmPlanner.py:91-97)."""
    rows, radii = job
    cc, nn, poses, nodes = _REF["cc"], _REF["nn"], _REF["poses"], _REF["nodes"]
    out = []
    t0 = time.perf_counter()
    for v, r in zip(rows, radii):
        m_near = nn(nodes, poses, poses[v], r)
        cand = [m for m in m_near if m != v]
        out.append((int(v), cand, [bool(cc.visible(poses[m], poses[v])) for m in cand]))
    return out, time.perf_counter() - t0


def port_rows(om, nodes, rows, k):
    """The oracle port (C, OpenMP) of the step for the given rows: kNN (k+1, self dropped) + visible."""
    import c_oracle as orc

    idx, dist = orc.knn(nodes, nodes[rows], k + 1)
    is_self = idx == np.asarray(rows)[:, None]
    drop = np.where(is_self.any(1), is_self.argmax(1), k)
    keep = np.ones_like(idx, bool)
    keep[np.arange(len(rows)), drop] = False
    nbr = idx[keep].reshape(len(rows), k)
    vis = om.visible2d(nodes[nbr.reshape(-1)], np.repeat(nodes[rows], k, axis=0))
    return nbr, vis.reshape(len(rows), k), dist


def run_reference(args):
    """The reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import multiprocessing as mp

    import c_oracle as orc
    import ref_shims

    cores = os.cpu_count() or 1
    mz, nodes, S_h, T_h = problem()
    n, k = len(nodes), K_NN
    img = mz.raster_host()                                   # the reference's `_img` ([W][H] uint8)
    om = orc.Map.__new__(orc.Map)                            # dense grid shared with the port (no copy)
    om.img, om.W, om.H = img, mz.W, mz.H
    orc.set_threads(cores)
    have_ref = ref_shims.reference_available()
    steps, warmup = args.steps, args.warmup
    budget = args.budget if args.budget > 0 else 150.0
    per_step_s = max(0.5, budget / max(1, steps + warmup))
    rng = np.random.default_rng(11)

    def radii_for(rows):
        # (untimed) the radius that makes the reference's radius query return exactly the row's k
        # nearest: just above the (k+1)-th smallest distance (the row itself is the first)
        _, dist = orc.knn(nodes, nodes[rows], k + 1)
        return np.nextafter(dist[:, k], np.inf)

    result = {}
    if have_ref:
        pool = mp.get_context("fork").Pool(cores, initializer=_ref_worker_init, initargs=(img, nodes))
        probe_rows = rng.integers(0, n, cores)
        probe_jobs = [([r], [x]) for r, x in zip(probe_rows, radii_for(probe_rows))]
        pool.map(_ref_rows, probe_jobs)                       # (workers import the reference here)
        t0 = time.perf_counter()
        pool.map(_ref_rows, probe_jobs)
        row_s = time.perf_counter() - t0                      # one row per core, in parallel
        rpw = max(1, int(per_step_s / max(row_s, 1e-3)))      # rows per worker and step
        ts, row_secs, checked = [], [], 0
        for s in range(warmup + steps):
            rows = rng.integers(0, n, cores * rpw)
            radii = radii_for(rows)
            jobs = [(rows[w::cores], radii[w::cores]) for w in range(cores)]
            t0 = time.perf_counter()
            res = pool.map(_ref_rows, jobs)
            dt = time.perf_counter() - t0
            if s >= warmup:
                ts.append(dt)
                row_secs += [sec / max(1, len(j[0])) for (_, sec), j in zip(res, jobs)]
            # parity of the port against the reference on this sample
            flat = [r for part, _ in res for r in part]
            rr = np.array([r[0] for r in flat])
            pn, pv, _ = port_rows(om, nodes, rr, k)
            for (v, cand, vis), a, b in zip(flat, pn, pv):
                if len(cand) == k:                            # (distance ties at the radius add candidates)
                    assert sorted(cand) == sorted(a.tolist()) and \
                        dict(zip(cand, vis)) == dict(zip(a.tolist(), b.tolist())), "port differs from the reference"
                    checked += 1
        pool.close()
        edges_step = cores * rpw * k
        val = edges_step * len(ts) / sum(ts)
        result = {"value": val, "kind": "reference", "cores": cores,
                  "sample": f"{cores * rpw} random rows of {n} per step ({edges_step} candidate edges): the reference's "
                            f"prmPlanner.nearest_neighbours over all {n} nodes + ImgCollisionChecker.visible per candidate, "
                            f"unmodified (oracle/_ref), multiprocessing over {cores} cores; {checked} rows also equal to the C port",
                  "value_1core": k / float(np.mean(row_secs)),
                  "ms_per_step": 1e3 * sum(ts) / len(ts)}
    # the C / OpenMP port on the same kind of sample (second figure; the only one when the
    # reference copy is not staged)
    t0 = time.perf_counter()
    port_rows(om, nodes, rng.integers(0, n, 64), k)
    probe = time.perf_counter() - t0
    prow = int(min(4096, max(64, 64 * min(per_step_s, 3.0) / max(probe, 1e-4))))
    pts = []
    for s in range(3):
        rows = rng.integers(0, n, prow)
        t0 = time.perf_counter()
        port_rows(om, nodes, rows, k)
        pts.append(time.perf_counter() - t0)
    port = {"value": prow * k / float(np.median(pts)), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{prow} random rows per step, oracle/oracle.c (brute-force kNN over {n} nodes + Bresenham) with OpenMP on {cores} threads"}
    if not result:
        result = dict(port, ms_per_step=1e3 * float(np.median(pts)))
    cfg = config_dict(mz, n, len(S_h))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": result["value"], "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": steps, "warmup": warmup,
        "ms_per_step": result["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": result["value"], "unit": UNIT, "cores": cores, "kind": result["kind"],
                         "sample": result["sample"], "value_1core": result.get("value_1core"), "port": port},
        "e2e": {"value": result["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_subprocess(budget_s):
    """`bench.py --impl reference` in a child process with a small budget; its cpu_baseline object."""
    try:
        o = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2",
                            "--warmup", "1", "--budget", str(budget_s)], capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
        line = [l for l in o.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:      # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--budget", type=float, default=0.0,
                    help="reference arm: seconds of CPU sampling for the whole run (0 = 150)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
