"""Batched kNN in d = 4 (arm-like cloud) and d = 3 (isotropic cube), filter on / off -- run under gpurun."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sbp_env_b200 as sg  # noqa: E402
from sbp_env_b200 import _lib  # noqa: E402
from sbp_env_b200.runtime import tune  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


n = m = 50_000
for name, d, scale, off in (("arm d=4", 4, [1024, 1024, 2 * np.pi, 2 * np.pi], [0, 0, -np.pi, -np.pi]),
                            ("cube d=3", 3, [300, 300, 300], [0, 0, 0])):
    P = torch.rand((n, d), generator=g, device=dev, dtype=torch.float64) * torch.tensor(scale, device=dev) + torch.tensor(off, device=dev)
    tree = sg.NodeStore(d, capacity=n)
    tree.extend(P)
    X = tree.rows_device(0, m)
    for filt in (0, 1):
        tune("knn_filter", filt)
        ms = timeit(lambda: tree.knn(X, 11, return_dist=False))
        print(f"{name}: filter={filt} {ms:.3f} ms  {n * m / ms * 1e3:.3e} evals/s", flush=True)
tune("knn_filter", 1)
