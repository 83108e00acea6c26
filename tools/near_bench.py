"""cfg4-like radius query: 10 000 queries x 1M nodes on a 16384^2 extent, r = 64 -- run under gpurun."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sbp_env_b200 as sg  # noqa: E402
from sbp_env_b200 import _lib  # noqa: E402
from sbp_env_b200.runtime import tune  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1)
n, m = 1_000_000, 10_000
tree = sg.NodeStore(2, capacity=n)
tree.extend(torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * 16384)
X = torch.rand((m, 2), generator=g, device=dev, dtype=torch.float64) * 16384
for filt in (1, 0):
    tune("near_filter", filt)
    for _ in range(2):
        rp, ix = tree.near(X, 64.0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        rp, ix = tree.near(X, 64.0)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"near_filter={filt}: {ms:.3f} ms per call ({n * m / ms * 1e3:.3e} evals/s), {ix.numel() / m:.1f} neighbours per query", flush=True)
