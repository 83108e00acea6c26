"""Round-2 probe (run under gpurun): cfg5 phase times with the r1 kernels, and the hop-count /
search-time profile of the 4096 queries on the UNDIRECTED roadmap (both directions of every
visible candidate, what the reference's radius graph contains, prmPlanner.py:91-101).

    python tools/r2_probe.py [scale]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import sbp_env_b200 as sg  # noqa: E402
import workloads  # noqa: E402

dev = torch.device("cuda", 0)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
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
    return float(np.median(ts))


t0 = time.perf_counter()
mz, cc, tree, nodes, S, T = workloads.cfg5_on_device(dev, scale)
torch.cuda.synchronize()
out = {"setup_s": time.perf_counter() - t0, "nodes": len(tree), "queries": len(S)}
k = 10
ev = {}


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    ev[name] = e


sg.prm_connect_knn(cc, tree, k)
torch.cuda.synchronize()
flush.zero_()
s0 = torch.cuda.Event(enable_timing=True)
s0.record()
nbr, vis = sg.prm_connect_knn(cc, tree, k, mark=mark)
torch.cuda.synchronize()
out["knn_ms"] = s0.elapsed_time(ev["knn"])
out["visible_ms"] = ev["edges"].elapsed_time(ev["visible"])
out["visible_fraction"] = float(vis.float().mean().item())
n = len(tree)
rows = torch.arange(n, device=dev).repeat_interleave(k)
src, dst, keep = nbr.reshape(-1), rows, vis.reshape(-1)
# directed (r1) and undirected graphs
g_dir = sg.Roadmap(n)
out["csr_directed_ms"] = timeit(lambda: g_dir.build_from_knn(nbr, vis))
src2 = torch.cat([src[keep], dst[keep]])
dst2 = torch.cat([dst[keep], src[keep]])
g_und = sg.Roadmap(n)
out["csr_undirected_ms"] = timeit(lambda: g_und.build_from_edges(src2, dst2))
out["edges_directed"], out["edges_undirected_with_duplicates"] = g_dir.n_edges, g_und.n_edges
# query endpoints -> nearest visible roadmap node (k = 1 connect, like r1's tools/bench_configs.py)
SG = torch.from_numpy(np.concatenate([S, T])).to(dev)
nb_q, vis_q = sg.prm_connect_knn(cc, tree, 1, queries=SG)
out["connect_ms"] = timeit(lambda: sg.prm_connect_knn(cc, tree, 1, queries=SG))
ends = torch.where(vis_q[:, 0], nb_q[:, 0], torch.full_like(nb_q[:, 0], -1))
q = len(S)
s_idx, t_idx = ends[:q].contiguous(), ends[q:].contiguous()
for name, g in (("directed", g_dir), ("undirected", g_und)):
    nq = min(q, 256 if name == "undirected" else q)          # bound the probe's time
    t0 = time.perf_counter()
    hops = g.shortest_paths(s_idx[:nq], t_idx[:nq])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    hh = hops.cpu().numpy()
    out[f"bfs_{name}"] = {"queries": nq, "ms": dt * 1e3, "ms_per_query_slot": dt * 1e3 / max(1, nq),
                          "reachable_fraction": float((hh >= 0).mean()),
                          "mean_hops": float(hh[hh >= 0].mean()) if (hh >= 0).any() else None,
                          "max_hops": int(hh.max())}
print(json.dumps(out, indent=1))
