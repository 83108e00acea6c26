"""Where the host-to-host step of bench.py spends its time (cfg5): append from pinned memory,
connection into a device buffer, connection into pinned host memory, copies.  Run under gpurun."""
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
pinned = nodes.cpu().pin_memory()
code_h = torch.empty((n, k), dtype=torch.int32).pin_memory()
code_d = torch.empty((n, k), dtype=torch.int32, device=dev)
order = torch.empty((n,), dtype=torch.int32, device=dev)
order_h = torch.empty((n,), dtype=torch.int32).pin_memory()
x_dev = torch.empty((n, 2), dtype=torch.float64, device=dev)
tree2 = sg.NodeStore(2, capacity=n)


def timeit(name, fn, reps=5):
    g = sg.CapturedStep(fn, device=0)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{name:60s} {ts[len(ts) // 2]:.3f} ms", flush=True)


def append_zero_copy():
    tree2.truncate(0)
    tree2.extend(pinned)


def append_copy():
    x_dev.copy_(pinned, non_blocking=True)
    tree2.truncate(0)
    tree2.extend(x_dev)


timeit("append, kernel reads pinned host rows", append_zero_copy)
timeit("append, H2D copy + kernel", append_copy)
timeit("connect -> device buffer", lambda: sg.prm_connect_knn(cc, tree2, k, packed_out=code_d, order_out=order, want_rows=False))
timeit("connect -> pinned host buffer (zero-copy stores)", lambda: sg.prm_connect_knn(cc, tree2, k, packed_out=(code_h.data_ptr(), False), order_out=order, want_rows=False))
timeit("D2H copy of the packed rows (40 MB)", lambda: code_h.copy_(code_d, non_blocking=True))
timeit("D2H copy of the permutation (4 MB)", lambda: order_h.copy_(order, non_blocking=True))
tree2.ctx.tune("prm_slices", 1)
