"""cfg5 PRM connection (1M nodes on the 32k^2 maze, k = 10): time of the captured step for a few
settings of the knobs: row slices whose walk overlaps the next slice's kNN (default sweep), or --
with any argument -- first ring radius x nodes per cell.  Run under gpurun.

    python tools/prm_sweep.py [walk | slices | cells]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import sbp_env_b200 as sg  # noqa: E402
import workloads  # noqa: E402

dev = torch.device("cuda", 0)
mz, cc, tree, nodes, S, T = workloads.cfg5_on_device(dev)
n, k = len(tree), 10
ctx = tree.ctx
packed = torch.empty((n, k), dtype=torch.int32, device=dev)
order = torch.empty((n,), dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ref = None
import itertools

mode = sys.argv[1] if len(sys.argv) > 1 else "walk"
if mode == "slices":
    grid = [(150, 0, sl, 1) for sl in (1, 2, 3, 4, 6, 8)]
elif mode == "cells":
    grid = [(d, r, 1, 1) for d, r in itertools.product((100, 150, 200, 300), (0, 1, 2))]
elif mode == "build":
    grid = [(150, 0, 1, 1)] * 4
else:
    grid = [(150, 0, 1, ws) for ws in (0, 1, 0, 1)]
for it, (dens, rinit, sl, ws) in enumerate(grid):
    if True:
        if mode == "build":
            ctx.tune("prm_build_ctas", (3, 2, 1, 3)[it])
            print("build CTAs per SM", (3, 2, 1, 3)[it], end=": ")
        ctx.tune("knn_cell_density", dens)
        ctx.tune("knn_cells_rinit", rinit)
        ctx.tune("prm_slices", sl)
        ctx.tune("prm_walk_sorted", ws)
        g = sg.PRMConnectGraph(cc, tree, k, packed_out=packed, want_rows=False, order_out=order)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            g.replay()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        node = torch.empty_like(packed)
        node[order.long()] = packed
        if ref is None:
            ref = node.clone()
        same = bool(torch.equal(ref, node))
        ts.sort()
        print(f"density {dens / 100:.2f} rinit {rinit} slices {sl} walk_sorted {ws}: {ts[len(ts) // 2]:.4f} ms (min {ts[0]:.4f}) same={same}", flush=True)
        del g
