"""Sweep of the kNN tuning knobs on cfg2 (50k x 50k, k=11) -- run under gpurun."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sbp_env_b200 as sg  # noqa: E402
from sbp_env_b200 import _lib  # noqa: E402
from sbp_env_b200.runtime import tune  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
lib = _lib.load()


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n, m in ((50_000, 50_000), (1_000_000, 100_000)):
    scale = torch.tensor([579.0, 581.0], device=dev, dtype=torch.float64)
    tree = sg.NodeStore(2, capacity=n)
    tree.extend(torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * scale)
    X = tree.rows_device(0, m)
    for filt, var, chunks in [(0, 0, 0)] + [(1, v, c) for v in (0, 2) for c in (0, 4, 6, 8, 10, 12, 16)]:
        tune("knn_filter", filt)
        tune("knn_filter_variant", var)
        tune("knn_filter_chunks", chunks)
        ms = timeit(lambda: tree.knn(X, 11, return_dist=False))
        print(f"n={n} m={m} filter={filt} variant={var} chunks={chunks}: {ms:.4f} ms  {n * m / ms * 1e3:.3e} evals/s", flush=True)
