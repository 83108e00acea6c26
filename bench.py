#!/usr/bin/env python
"""bench.py -- the hot path on BASELINE.json configs[1]: PRM roadmap connection on
maps/intel_lab (579x581), 50 000 free samples, k = 10 nearest neighbours, the
500 000 candidate edges checked through visible_batch.

One step = one pass of the hot path over one batch: exact fp64 kNN of 50 000 rows
against the 50 000-node roadmap (2.5e9 pairs; answered from a cell grid rebuilt inside
every step, the brute-force scan behind it is timed separately as roofline_bruteforce),
and integer-Bresenham visibility of the 500 000 candidate edges.  At N GPUs (weak
scaling) rank 0's batch is the roadmap's own nodes (= PRM build, self excluded) and
rank r > 0 connects a further batch of 50 000 free samples to the replicated
roadmap; the per-rank neighbour/visibility blocks are all-gathered (NCCL) inside the
step -- the only collective, used to assemble the adjacency.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the CPU restatement of the reference (oracle/, all host
threads) on the same workload, bounded sample per step.  Prints ONE JSON line.
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
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_NODES = 50_000
K_NN = 10
MAP = "intel_lab"
METRIC = "edge collision checks/sec (PRM candidate edges: exact kNN connect + visible_batch)"
UNIT = "edges/s"
# dram__bytes_read.sum + dram__bytes_write.sum of one k_knn_filter launch on this workload
# (ncu --set full, profiles/r1_knn_filter_full.txt)
FILTER_DRAM_TRAFFIC = 18273024
# the same for one k_knn_cells_t launch (profiles/r1_knn_cells_full.txt), and that capture's
# instruction / L1 sector counts (static facts about the kernel on this workload)
CELLS_DRAM_TRAFFIC = 1956864
CELLS_NCU = {"warp_instructions": 15318739, "l1_global_load_sectors": 2879287,
             "issue_active_pct": 39.4, "threads_per_instruction": 21.8, "warps_per_sm": 9.0,
             "source": "profiles/r1_knn_cells_full.txt"}
WORKLOAD = ("cfg2: PRM roadmap connection on intel_lab 579x581, 50000 free samples, k=10 -> 500000 "
            "candidate edges per step (exact kNN of 50000 rows among 50000 nodes = 2.5e9 pairs + Bresenham visible_batch)")


def load_free():
    from _util import load_map

    return load_map(MAP)            # bool [W][H], the reference's `_img == 1`


def draw_free_samples(free, n, seed, feasible_fn):
    """First n accepted of default_rng(seed).random((., 2)) * [W, H] filtered by
    feasible."""
    W, H = free.shape
    rng = np.random.default_rng(seed)
    out = []
    have = 0
    while have < n:
        c = rng.random((2 * n, 2)) * [W, H]
        c = c[feasible_fn(c)]
        out.append(c)
        have += len(c)
    return np.ascontiguousarray(np.concatenate(out)[:n])


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
    from sbp_env_b200 import _lib

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    free = load_free()
    W, H = free.shape
    cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8), device=local)
    ctx = cc.device_map.ctx
    nodes = draw_free_samples(free, N_NODES, 0, cc.feasible_batch)
    tree = sg.NodeStore(2, capacity=N_NODES, device=local)
    tree.extend(nodes)
    if rank == 0:
        batch_host, queries = nodes, None
    else:
        batch_host = draw_free_samples(free, N_NODES, 100 + rank, cc.feasible_batch)
        queries = torch.from_numpy(batch_host).to(dev)
    m, k = N_NODES, K_NN
    edges_per_rank = m * k
    gathered = {}

    names = ("knn", "edges", "visible", "gather")

    # N > 1: the packed adjacency is assembled on every rank by ONE kernel that packs the block
    # and stores it over NVLink into every rank's symmetric buffer, plus a signal-pad barrier
    # (PeerAdjacencyGather); NCCL all-gather of the packed block when peer memory is unavailable
    peer = None
    gather_kind = "none (N = 1)"
    if world > 1:
        try:
            peer = sg.distributed.PeerAdjacencyGather(m, k, device=local)
            gather_kind = f"one pack + {peer.kind} kernel, symmetric-memory barrier"
        except Exception as e:      # noqa: BLE001
            gather_kind = f"NCCL all-gather of the packed block (peer memory unavailable: {type(e).__name__})"

    # where the fabric has the multicast mapping, the visibility kernel itself sends the packed
    # adjacency through the switch (one captured graph per buffer half) and only the barrier follows
    fused_exchange = peer is not None and peer.fused
    if fused_exchange:
        gather_kind = "fused: the visibility kernel stores the packed adjacency through the NVSwitch " \
                      "multicast mapping as edges are decided; symmetric-memory barrier"

    def gather(nbr, vis):
        if world == 1:
            gathered["nbr"], gathered["vis"] = nbr, vis
        elif peer is not None:
            gathered["packed"] = peer.gather(nbr, vis)
        else:
            gathered["packed"] = sg.distributed.all_gather_adjacency(nbr, vis, unpack=False)

    def eager_step(ev=None):
        idx = [0]

        def mark(_name):
            if ev is not None:
                ev[idx[0]].record()
                idx[0] += 1

        if fused_exchange:
            h = peer_step[0] & 1
            peer_step[0] += 1
            if rendezvous_eager[0]:
                nbr, vis = sg.prm_connect_knn(cc, tree, k, queries=queries, mark=mark,
                                              packed_out=peer.target_with_rendezvous(h))
                gathered["packed"] = peer.view(h)
            else:
                nbr, vis = sg.prm_connect_knn(cc, tree, k, queries=queries, mark=mark, packed_out=peer.target(h))
                gathered["packed"] = peer.complete(h)
        else:
            nbr, vis = sg.prm_connect_knn(cc, tree, k, queries=queries, mark=mark)
            gather(nbr, vis)
        mark("gather")
        return nbr, vis

    peer_step = [0]
    rendezvous_eager = [fused_exchange and os.environ.get("BENCH_EXCHANGE", "barrier") == "rendezvous"]

    # The step is ~20 short kernels; issued one by one from Python the launch gaps cost ~15 %,
    # so the timed loop replays the step from a CUDA graph (the library's
    # PRMConnectGraph); the all-gather (N > 1) follows the replay on the same stream.
    barrier_in_graph = False
    rendezvous = False
    if fused_exchange:
        mode = os.environ.get("BENCH_EXCHANGE", "barrier")
        if mode == "rendezvous":
            # the kernel itself meets the other ranks: its last CTA publishes an epoch through the
            # switch and waits for every rank's (no barrier launch).  Measured at N = 2: 0.126 ms per
            # step against 0.120 ms with the captured symmetric-memory barrier, hence not the default.
            gsteps = [sg.PRMConnectGraph(cc, tree, k, queries=queries,
                                         packed_out=peer.target_with_rendezvous(h)) for h in (0, 1)]
            rendezvous = True
            gather_kind += " -- replaced by a rendezvous inside the kernel (last CTA: multicast epoch + wait)"
        else:
            # the barrier is captured into the step's graph when torch's symmetric-memory barrier
            # allows it (BENCH_EXCHANGE=eager_barrier keeps it as a separate launch after the replay)
            if mode != "eager_barrier":
                try:
                    gsteps = [sg.PRMConnectGraph(cc, tree, k, queries=queries, packed_out=peer.target(h),
                                                 after=(lambda h=h: peer.complete(h))) for h in (0, 1)]
                    barrier_in_graph = True
                except Exception as e:      # noqa: BLE001
                    print(f"[bench] barrier not capturable ({type(e).__name__}: {e}); eager barrier", file=sys.stderr)
            ok = torch.tensor([1 if barrier_in_graph else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            barrier_in_graph = bool(ok.item())
            if not barrier_in_graph:
                gsteps = [sg.PRMConnectGraph(cc, tree, k, queries=queries, packed_out=peer.target(h)) for h in (0, 1)]
            gather_kind += " (captured in the step's graph)" if barrier_in_graph else " (separate launch)"
        dist.barrier()
    else:
        gstep = sg.PRMConnectGraph(cc, tree, k, queries=queries)

    def step():
        if fused_exchange:
            h = peer_step[0] & 1
            peer_step[0] += 1
            nbr, vis = gsteps[h].replay()
            gathered["packed"] = peer.view(h) if (barrier_in_graph or rendezvous) else peer.complete(h)
        else:
            nbr, vis = gstep.replay()
            gather(nbr, vis)
        return nbr, vis

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

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
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    def align_ranks():
        # N > 1: the untimed L2 flush does not take the same time on every rank; without this the
        # timed step of an early rank would include waiting for a rank that is still flushing
        if world > 1:
            if peer is not None:
                peer.hdl.barrier(channel=2)
            else:
                dist.barrier()

    for s in range(args.steps):
        flush.zero_()                      # evict L2 between timed iterations (not timed)
        align_ranks()                      # (not timed) all ranks start the step together
        starts[s].record()
        step()
        ends[s].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall1 = time.time()
    step_ms = np.array([starts[s].elapsed_time(ends[s]) for s in range(args.steps)])
    total_ms = torch.tensor([float(step_ms.sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None

    # ---- phase breakdown and the dominant kernel's duration: the same step launched eagerly.
    # A spin kernel ahead of each step lets the host run ahead, so that the events measure
    # the GPU and not the Python launch path; sbp_tune("time_kernels") brackets k_knn_filter
    # with its own CUDA events on the launching stream.
    import ctypes as C

    lib = _lib.load()
    psteps = min(args.steps, 10)
    for _ in range(2):
        eager_step()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    lib.sbp_tune(b"time_kernels", 2)       # event pair around the cell-grid kNN kernel
    p_start = [torch.cuda.Event(enable_timing=True) for _ in range(psteps)]
    phase_ev = [[torch.cuda.Event(enable_timing=True) for _ in names] for _ in range(psteps)]
    for s in range(psteps):
        flush.zero_()
        torch.cuda._sleep(4_000_000)       # ~2 ms head start for the host
        align_ranks()
        p_start[s].record()
        eager_step(phase_ev[s])
    torch.cuda.synchronize()
    lib.sbp_tune(b"time_kernels", 0)
    kt_ms, kt_n = C.c_double(), C.c_int64()
    _lib.check(lib.sbp_ctx_kernel_time(ctx.handle, C.byref(kt_ms), C.byref(kt_n)))
    cells_ms, cells_n = kt_ms.value / max(1, kt_n.value), int(kt_n.value)
    launches_per_step = (ctx.launches - launches0) // psteps
    launches = launches_per_step * args.steps
    ph = {}
    for i, nme in enumerate(names):
        prev = [p_start[s] if i == 0 else phase_ev[s][i - 1] for s in range(psteps)]
        ph[nme] = float(np.mean([prev[s].elapsed_time(phase_ev[s][i]) for s in range(psteps)]))

    # parity spot check of the timed configuration (outside the timed region)
    nbr, vis = step()
    torch.cuda.synchronize()
    n_vis = int(vis.sum().item())
    nbr_e, vis_e = eager_step()
    assert torch.equal(nbr, nbr_e) and torch.equal(vis, vis_e), "graph replay differs from the eager step"
    if world > 1:      # the assembled adjacency holds this rank's block at its place
        torch.cuda.synchronize()
        gn, gv = sg.distributed.unpack_adjacency(gathered["packed"][rank * m:(rank + 1) * m])
        assert torch.equal(gn, nbr_e) and torch.equal(gv, vis_e), "assembled adjacency differs from the local block"

    # ---- the brute-force scan (the path of queries the cell grid cannot answer), same rows:
    # kNN alone with sbp_tune("knn_cells", 0), k_knn_filter bracketed by its own events
    Xq = queries if queries is not None else tree.rows_device(0, m)
    kq = k if queries is not None else k + 1
    lib.sbp_tune(b"knn_cells", 0)
    lib.sbp_tune(b"time_kernels", 1)
    try:
        idx_bf, _ = tree.knn(Xq, kq, return_dist=False)
        torch.cuda.synchronize()
        _lib.check(lib.sbp_ctx_kernel_time(ctx.handle, C.byref(kt_ms), C.byref(kt_n)))
        bf_a, bf_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bf_ms = []
        for _ in range(psteps):
            flush.zero_()
            torch.cuda._sleep(2_000_000)
            bf_a.record()
            tree.knn(Xq, kq, return_dist=False)
            bf_b.record()
            torch.cuda.synchronize()
            bf_ms.append(bf_a.elapsed_time(bf_b))
        _lib.check(lib.sbp_ctx_kernel_time(ctx.handle, C.byref(kt_ms), C.byref(kt_n)))
        filter_ms, filter_n = kt_ms.value / max(1, kt_n.value), int(kt_n.value)
        bf_call_ms = float(np.median(bf_ms))
    finally:
        lib.sbp_tune(b"time_kernels", 0)
        lib.sbp_tune(b"knn_cells", 2)
    idx_grid, _ = tree.knn(Xq, kq, return_dist=False)
    assert torch.equal(idx_bf, idx_grid), "cell-grid kNN differs from the brute-force scan"

    # ---- end to end through the public API with host buffers -------------------
    pinned = torch.from_numpy(batch_host).pin_memory()
    # the step's result travels to the host as the packed adjacency, 4 bytes per candidate edge:
    # (neighbour + 1) * 2 + visible (distributed.pack_adjacency / unpack_adjacency)
    code_h = torch.empty((m, k), dtype=torch.int32).pin_memory()
    tree2 = sg.NodeStore(2, capacity=N_NODES, device=local) if rank == 0 else None

    x_dev = torch.empty((m, 2), dtype=torch.float64, device=dev)
    code_dev = torch.empty((m, k), dtype=torch.int32, device=dev)
    e2e_copy = os.environ.get("BENCH_E2E_COPY", "0") == "1"

    def e2e_body():
        # H2D of this step's samples: rank 0 rebuilds its roadmap from them, and the append kernel
        # reads the pinned host rows itself (zero-copy: transfer and AoS -> SoA conversion in one
        # launch; BENCH_E2E_COPY=1: explicit copy first); ranks > 0 query with them, so they are copied
        if rank != 0 or e2e_copy:
            x_dev.copy_(pinned, non_blocking=True)
        # D2H of the adjacency block: the visibility kernel packs it and (default) stores it straight
        # into the pinned host buffer -- mapped into the device's address space (UVA) -- as the edges
        # are decided, so the transfer overlaps the kernel; BENCH_E2E_COPY=1: device buffer + copy
        dst = code_dev if e2e_copy else (code_h.data_ptr(), False)
        if rank == 0:
            tree2.truncate(0)
            tree2.extend(x_dev if e2e_copy else pinned)    # roadmap rebuilt from the samples
            sg.prm_connect_knn(cc, tree2, k, packed_out=dst)
        else:
            sg.prm_connect_knn(cc, tree, k, queries=x_dev, packed_out=dst)
        if e2e_copy:
            code_h.copy_(code_dev, non_blocking=True)

    # the user-facing call: the whole host-to-host step captured once (CapturedStep) and
    # replayed; H2D, store rebuild, connection and D2H all run inside every replay
    e2e_graph = sg.CapturedStep(e2e_body, device=local)

    def e2e_step():
        e2e_graph.replay()

    for _ in range(max(3, args.warmup)):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_s = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e_e = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    for s in range(args.steps):
        flush.zero_()
        align_ranks()
        e_s[s].record()
        e2e_step()
        e_e[s].record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([sum(e_s[s].elapsed_time(e_e[s]) for s in range(args.steps))], device=dev,
                          dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    e2e_ms = float(e2e_ms.item())
    nb_e2e, vs_e2e = sg.distributed.unpack_adjacency(code_h)
    assert int(vs_e2e.sum()) == n_vis and torch.equal(nb_e2e, nbr.cpu()), \
        "e2e result differs from the device-resident run"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        hbm_peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    evals = float(m) * float(len(tree))
    knn_s = ph["knn"] / 1e3
    filt_s = filter_ms / 1e3
    cells_s = cells_ms / 1e3
    n_nodes = float(len(tree))
    # Dominant kernel: k_knn_cells_t (exact kNN from the cell grid, one thread per query).
    # ALGORITHMIC bytes per launch = every byte it has to touch once: node coordinates (16 B) and
    # cell-sorted ids (4 B) per node, the cell starts (one int per cell, 1.5 nodes per cell), the
    # query rows (16 B) and per query the index row ((k + 1) x 8 B), its seed (8 B) and flag (1 B).
    alg_cells = n_nodes * (16 + 4 + 4.0 / 1.5) + m * (16 + kq * 8 + 9)
    roof = {"kernel": f"k_knn_cells_t<2,{kq}> (exact kNN from the cell grid, one thread per query; "
                      f"{100 * cells_ms / sum(ph.values()):.0f}% of the step)",
            "bound": "hbm", "achieved": alg_cells / cells_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": alg_cells / cells_s / 1e9 / hbm_peak, "traffic": CELLS_DRAM_TRAFFIC,
            "peak_source": peak_src, "kernel_ms": cells_ms, "kernel_launches_timed": cells_n,
            "algorithmic_bytes_per_launch": alg_cells,
            "note": "a gather over L1 / L2-resident arrays (1.6 MB), not an HBM stream: the kernel is "
                    "bound by its warps' serial chains (dependent loads cell start -> id -> node, "
                    "lock-step sorted insertions) at ~9 warps per SM; see issue and l2_gather",
            "issue": dict(CELLS_NCU, warp_instructions_per_query=CELLS_NCU["warp_instructions"] / float(m),
                          issue_slot_frac=CELLS_NCU["warp_instructions"]
                          / (148 * 4 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6) / cells_s),
            "l2_gather": {"l1_sector_GBps": CELLS_NCU["l1_global_load_sectors"] * 32 / cells_s / 1e9},
            "queries_per_s_kernel": m / cells_s,
            "effective_pair_rate_per_s": evals / cells_s}
    # The brute-force scan k_knn_filter (all m x n pairs), which the step used before the cell grid
    # and which still serves the queries the grid cannot answer.  ALGORITHMIC bytes per launch
    # (8 d / Q_tile per distance evaluation when Q_tile queries share a streamed
    # node; the scan streams the fp32 shadow, 4 (d + 1) bytes per node, to CTAs of Q_tile = 256
    # queries) + the query rows + the candidate lists.
    q_tile = 256
    alg_bytes = evals * 4 * 3 / q_tile + m * 16 + m * 8 + m * 24 * 4
    brute = {"kernel": "k_knn_filter<2> (brute-force scan: all 2.5e9 pairs), kNN call with "
                       "sbp_tune('knn_cells', 0)",
             "bound": "hbm", "achieved": alg_bytes / filt_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
             "frac": alg_bytes / filt_s / 1e9 / hbm_peak, "traffic": FILTER_DRAM_TRAFFIC,
             "kernel_ms": filter_ms, "knn_call_ms": bf_call_ms, "kernel_launches_timed": filter_n,
             "note": "the scan shares every streamed node among 256 queries, so it is bound by the "
                     "FMA pipe, not by HBM: see fma_pipe"}

    def mb(which, nbytes, iters):
        ms, units = C.c_double(), C.c_double()
        _lib.check(lib.sbp_microbench(ctx.handle, which, nbytes, iters, C.byref(ms), C.byref(units)))
        return units.value / (ms.value / 1e3)

    fp64_peak = mb(2, 0, 4096)
    # FMA-pipe roofline of the scan: d fp32 FMAs per (query, node) pair (norm expansion, issued as
    # packed FFMA2) against the MEASURED packed-FMA rate of this GPU (sbp_microbench 3: independent
    # fma.rn.f32x2 chains; 127 FMA/clk/SM = the scalar FFMA rate as well).  The scan also needs
    # ~1.1 min/compare lane-operations per pair (FMNMX3 runs at half the FMNMX rate, microbench 5/6)
    # which share issue and, partly, datapath with the FMAs (microbench 7: 8 FFMA2 + 8 FMNMX take
    # 20.7 cycles, not 16): `mixed_bound_frac` relates the kernel to that instruction mix.
    fma_peak = mb(3, 0, 4096)
    mix_rate = mb(7, 0, 4096)            # lane-instructions/s of the 1:1 FFMA2 : FMNMX mix
    sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
    # per warp and group of 8 nodes x 4 queries: 32 FFMA2 + 12 FMNMX3 (= 24 FMNMX slots) + 4 FMNMX
    # + 7 FSETP ~ 32 packed + 35 ALU operations ~ 67 mixed instructions at the measured mix rate
    mix_bound_s = evals / 32.0 * 67.0 / mix_rate
    brute["fma_pipe"] = {"model": "d = 2 fp32 FMAs per (query, node) pair, packed FFMA2",
                        "achieved_fma_per_s": evals * 2 / filt_s,
                        "peak_fma_per_s_measured": fma_peak,
                        "frac": evals * 2 / filt_s / fma_peak,
                        "mixed_bound_ms": mix_bound_s * 1e3,
                        "mixed_bound_frac": mix_bound_s / filt_s,
                        "dist_evals_per_s_kernel": evals / filt_s,
                        "fp64_equivalent": {"fp64_ops_per_eval": 6, "ops_per_s": evals * 6 / (bf_call_ms / 1e3),
                                            "measured_peak_unfused_fp64_ops_per_s": fp64_peak,
                                            "ratio": evals * 6 / (bf_call_ms / 1e3) / fp64_peak}}

    # ---- micro: uniform random edge batches on the same map ------
    micro = None
    if world == 1:        # N = 1 only: keeps the N > 1 runs short
        micro = {}
        smem_gather = mb(0, int(cc.device_map.info()["packed_bytes"]), 2048)
        l2_gather = mb(1, 64 << 20, 512)
        hbm_gather = mb(1, 8 << 30, 512)
        micro["measured_gather_peaks"] = {"smem_random_word_gathers_per_s": smem_gather,
                                          "l2_random_sector_gathers_per_s_64MiB": l2_gather,
                                          "hbm_random_sector_gathers_per_s_8GiB": hbm_gather}
        g = torch.Generator(device=dev)
        g.manual_seed(4)
        scale = torch.tensor([W, H], device=dev, dtype=torch.float64)

        def time_visible(P, Q, iters=5):
            cc.visible_batch(P, Q)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for _ in range(iters):
                flush.zero_()
                a.record()
                r = cc.visible_batch(P, Q)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            px = (torch.maximum((P[:, 0].trunc() - Q[:, 0].trunc()).abs(),
                                (P[:, 1].trunc() - Q[:, 1].trunc()).abs()) + 1).sum().item()
            cnt = torch.zeros(1, dtype=torch.int64, device=dev)
            cc._visible_device(P, Q, px_tested=cnt)
            t = float(np.median(ts)) / 1e3
            n = P.shape[0]
            ab = 33.0 * n + 0.125 * px
            return {"edges": n, "ms": t * 1e3, "interpolants_full_line": px,
                    "interpolants_per_s": px / t, "pixel_tests_executed": int(cnt.item()),
                    "pixel_tests_per_s": int(cnt.item()) / t, "edges_per_s": n / t,
                    "visible_fraction": float(r.float().mean().item()),
                    "algorithmic_GBps": ab / t / 1e9, "hbm_frac": ab / t / 1e9 / hbm_peak}

        n_u = 1 << 24
        P = torch.rand((n_u, 2), generator=g, device=dev, dtype=torch.float64) * scale
        Q = torch.rand((n_u, 2), generator=g, device=dev, dtype=torch.float64) * scale
        micro["uniform_2^24_edges"] = time_visible(P, Q)
        # point feasibility: a pure stream of 16 B in + 1 B out per point
        cc.feasible_batch(P)
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
        micro["feasible_2^24_points"] = {"points": n_u, "ms": tf * 1e3, "points_per_s": n_u / tf,
                                         "algorithmic_GBps": 17.0 * n_u / tf / 1e9,
                                         "hbm_frac": 17.0 * n_u / tf / 1e9 / hbm_peak}
        fr = torch.from_numpy(nodes).to(dev)
        for L in (8, 32, 128):
            n_b = 1 << 22
            base = fr[torch.randint(0, len(nodes), (n_b,), generator=g, device=dev)]
            ang = torch.rand(n_b, generator=g, device=dev, dtype=torch.float64) * (2 * np.pi)
            Qb = base + torch.stack([ang.cos(), ang.sin()], 1) * L
            micro[f"free_start_len{L}_2^22_edges"] = time_visible(base, Qb)
        del P, Q

        # single-query nearest over a 1M-node tree (RRT-style, latency bound)
        big = sg.NodeStore(2, capacity=1 << 20, device=local)
        big.extend(torch.rand((1 << 20, 2), generator=g, device=dev, dtype=torch.float64) * scale)
        q1 = torch.rand((1, 2), generator=g, device=dev, dtype=torch.float64) * scale
        big.knn(q1, 1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            big.knn(q1, 1)
        b.record()
        torch.cuda.synchronize()
        t1 = a.elapsed_time(b) / 50 / 1e3
        micro["nearest_1q_over_2^20_nodes"] = {"ms": t1 * 1e3, "dist_evals_per_s": (1 << 20) / t1,
                                               "node_stream_GBps": (1 << 20) * 16 / t1 / 1e9}

    # ---- CPU baseline: the oracle port on this box's host cores, bounded sample ----
    cpu = cpu_prm_sample(free, nodes, target_s=12.0) if world == 1 else None

    out = {
        "metric": METRIC, "value": world * edges_per_rank * args.steps / (total_ms / 1e3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "map": "intel_lab 579x581 (occupancy fixture tests/golden/maps.npz)",
                   "nodes": N_NODES, "k": K_NN, "edges_per_step_per_gpu": edges_per_rank,
                   "dist_evals_per_step_per_gpu": evals, "l2": "flushed (256 MiB write) between timed steps; N > 1: ranks aligned by an untimed barrier after the flush",
                   "launch": "the step (and the e2e step) is replayed from a CUDA graph captured by the "
                             "library (PRMConnectGraph / CapturedStep); phases_ms and roofline.kernel_ms "
                             "come from the same step launched eagerly",
                   "multi_gpu": "rank r connects batch r of 50000 samples to the replicated roadmap; "
                                "the packed neighbour/visibility blocks are assembled on every rank inside the step"},
        "clocks": clocks,
        "e2e": {"value": world * edges_per_rank * args.steps / (e2e_ms / 1e3), "unit": UNIT,
                "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(pinned.numel() * 8),
                "d2h_bytes_per_step": int(code_h.numel() * 4),
                "result": "packed adjacency int32 [m][k] = (neighbour + 1) * 2 + visible",
                "d2h": "copy from a device buffer" if e2e_copy else
                       "stored by the visibility kernel straight into the pinned host buffer (zero-copy)",
                "h2d": "copy into a device buffer" if (e2e_copy or rank != 0) else
                       "read from the pinned host buffer by the node-store append kernel (zero-copy)"},
        "gpu_launches": int(launches),
        "gpu_launches_per_step": int(launches_per_step),
        "adjacency_gather": gather_kind,
        "roofline": roof,
        "roofline_bruteforce": brute,
        "cpu_baseline": cpu,
        "phases_ms": ph,
        "submetrics": {"knn_queries_per_s_per_gpu": m / knn_s,
                       "knn_effective_pairs_per_s_per_gpu": evals / knn_s,
                       "knn_bruteforce_dist_evals_per_s_per_gpu": evals / (bf_call_ms / 1e3),
                       "visible_batch_edges_per_s_per_gpu": edges_per_rank / (ph["visible"] / 1e3),
                       "prm_build_ms": total_ms / args.steps,
                       "roadmap_edges_visible": n_vis},
        "micro": micro,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------ CPU legs --------
def cpu_prm_rows(om, nodes, first, count, k):
    """The oracle port of the step for `count` rows: kNN (k+1, self dropped) + visible."""
    import c_oracle as orc

    idx, _ = orc.knn(nodes, nodes[first:first + count], k + 1)
    rows = np.arange(first, first + count)[:, None]
    is_self = idx == rows
    drop = np.where(is_self.any(1), is_self.argmax(1), k)
    keep = np.ones_like(idx, bool)
    keep[np.arange(count), drop] = False
    nbr = idx[keep].reshape(count, k)
    vis = om.visible2d(nodes[nbr.reshape(-1)], np.repeat(nodes[first:first + count], k, axis=0))
    return nbr, vis.reshape(count, k)


def cpu_prm_sample(free, nodes, target_s):
    import c_oracle as orc

    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    om = orc.Map(free)
    t0 = time.perf_counter()
    cpu_prm_rows(om, nodes, 0, 256, K_NN)
    probe = time.perf_counter() - t0
    rows = int(min(len(nodes), max(256, 256 * target_s / max(probe, 1e-4))))
    t0 = time.perf_counter()
    cpu_prm_rows(om, nodes, 0, rows, K_NN)
    dt = time.perf_counter() - t0
    # the reference itself is single-threaded: one core on a smaller sample too
    orc.set_threads(1)
    rows1 = max(256, rows // (2 * cores))
    t0 = time.perf_counter()
    cpu_prm_rows(om, nodes, 0, rows1, K_NN)
    dt1 = time.perf_counter() - t0
    orc.set_threads(cores)
    return {"value": rows * K_NN / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {rows} of {len(nodes)} rows of the same workload (kNN k+1 against all "
                      f"{len(nodes)} nodes + visible of their {rows * K_NN} edges), oracle/oracle.c with "
                      f"OpenMP on {cores} threads, {dt:.2f} s",
            "value_1core": rows1 * K_NN / dt1,
            "sample_1core": f"first {rows1} rows on one thread, {dt1:.2f} s"}


def run_reference(args):
    """The reference's algorithm on the host cores (oracle C port, all threads)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import c_oracle as orc

    cores = os.cpu_count() or 1
    orc.set_threads(cores)
    free = load_free()
    om = orc.Map(free)
    nodes = draw_free_samples(free, N_NODES, 0, om.feasible2d)
    t0 = time.perf_counter()
    cpu_prm_rows(om, nodes, 0, 256, K_NN)
    probe = time.perf_counter() - t0
    budget = 150.0 / max(1, args.steps + args.warmup)          # whole run within a few minutes
    rows = int(min(len(nodes), max(256, 256 * min(budget, 6.0) / max(probe, 1e-4))))
    for _ in range(args.warmup):
        cpu_prm_rows(om, nodes, 0, rows, K_NN)
    ts = []
    for s in range(args.steps):
        first = (s * rows) % max(1, len(nodes) - rows)
        t0 = time.perf_counter()
        cpu_prm_rows(om, nodes, first, rows, K_NN)
        ts.append(time.perf_counter() - t0)
    val = rows * K_NN * len(ts) / sum(ts)
    sample = (f"{rows} of {len(nodes)} rows per step (kNN k+1 against all nodes + visible of "
              f"{rows * K_NN} edges), oracle/oracle.c OpenMP x{cores}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(ts) / len(ts), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "nodes": N_NODES, "k": K_NN},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
