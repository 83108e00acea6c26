"""Tuning sweeps (run under gpurun): visible kernel variants x workloads; kNN chunking."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sbp_env_b200 as sg  # noqa: E402
from sbp_env_b200 import _lib  # noqa: E402
from _util import load_map  # noqa: E402
from sbp_env_b200.runtime import tune  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
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


free = load_map("intel_lab")
W, H = free.shape
cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8))
scale = torch.tensor([W, H], device=dev, dtype=torch.float64)
fp = torch.from_numpy(np.argwhere(free).astype(np.float64) + 0.5).to(dev)
n = 1 << 22
sets = {}
sets["uniform"] = (torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * scale,
                   torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * scale)
for L in (4, 16, 64, 256):
    base = fp[torch.randint(0, len(fp), (n,), generator=g, device=dev)]
    ang = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * (2 * np.pi)
    ln = (torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 0.5 + 0.75) * L
    sets[f"free_len{L}"] = (base, base + torch.stack([ang.cos(), ang.sin()], 1) * ln[:, None])
variants = [("adaptive", 0, c) for c in (1, 4, 8, 12, 16, 24, 33)] + [(f"group{G}", G, 8) for G in (4, 8, 32)]
print("workload            " + " ".join(f"{v[0]}{'' if v[1] else ':' + str(v[2]):>3s}".rjust(12) for v in variants))
for name, (P, Q) in sets.items():
    row = []
    for _, G, c in variants:
        tune("visible_group", G)
        tune("coop_below", c)
        row.append(timeit(lambda: cc.visible_batch(P, Q)))
    print(f"{name:20s}" + " ".join(f"{t * 1e3:10.1f}us" for t in row), " vis=%.3f" % cc.visible_batch(P, Q).float().mean().item())
tune("visible_group", 0)
tune("coop_below", 12)

tree = sg.NodeStore(2, capacity=50000)
tree.extend(torch.rand((50000, 2), generator=g, device=dev, dtype=torch.float64) * scale)
X = tree.rows_device(0, 50000)
for S in (0, 1, 2, 3, 4, 6):
    tune("knn_chunks", S)
    print("knn 50k x 50k k=11 chunks", S, "ms", timeit(lambda: tree.knn(X, 11, return_dist=False)))
tune("knn_chunks", 0)
