"""Small single-kernel drivers for `ncu --set full` captures (run under gpurun).

    python tools/prof.py knn        # cfg2 kNN: 50k rows x 50k nodes, k=11
    python tools/prof.py visible    # 2^22 uniform edges on intel_lab (smem-resident map)
    python tools/prof.py visible_l2 # 2^22 edges on a 16384^2 blocky-noise grid (L2 path)
    python tools/prof.py arm        # 2^18 arm edges on the 1024^2 map
    python tools/prof.py nearest    # 1 query over 2^20 nodes
    python tools/prof.py step       # the bench step (cfg2 PRM connect: kNN + edges + visible), eager launches
    python tools/prof.py cfg5       # cfg5 PRM connect: 1M nodes on the 32k^2 maze, k = 10 (kNN + fused visible)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sbp_env_b200 as sg  # noqa: E402
from _util import load_map  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "knn"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)


def timeit(fn):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if what == "step":
    from _util import draw_free_samples

    free = load_map("intel_lab")
    cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8))
    nodes = draw_free_samples(free, 50_000, 0, cc.feasible_batch)
    tree = sg.NodeStore(2, capacity=50_000)
    tree.extend(nodes)
    ms = timeit(lambda: sg.prm_connect_knn(cc, tree, 10))
    print(what, "ms (eager, host-bound)", ms)
elif what == "cfg5":
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import workloads

    mz, cc, tree, nodes, S, T = workloads.cfg5_on_device(dev)
    # the bench's step: packed adjacency in cell-sorted row order (coalesced stores) + the permutation
    packed = torch.empty((len(tree), 10), dtype=torch.int32, device=dev)
    order = torch.empty((len(tree),), dtype=torch.int32, device=dev)
    ms = timeit(lambda: sg.prm_connect_knn(cc, tree, 10, packed_out=packed, order_out=order, want_rows=False))
    print(what, "ms (eager)", ms)
elif what in ("knn", "nearest"):
    free = load_map("intel_lab")
    W, H = free.shape
    scale = torch.tensor([W, H], device=dev, dtype=torch.float64)
    n = 50_000 if what == "knn" else 1 << 20
    tree = sg.NodeStore(2, capacity=n)
    tree.extend(torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * scale)
    X = tree.rows_device(0, n) if what == "knn" else torch.rand((1, 2), generator=g, device=dev, dtype=torch.float64) * scale
    k = 11 if what == "knn" else 1
    ms = timeit(lambda: tree.knn(X, k, return_dist=False))
    print(what, "ms", ms, "evals/s", X.shape[0] * n / ms * 1e3)
elif what in ("visible", "visible_l2"):
    if what == "visible":
        free = load_map("intel_lab")
        cc = sg.ImgCollisionChecker.from_free_mask(free.astype(np.uint8))
        W, H = free.shape
    else:
        W = H = 16384
        coarse = (torch.rand((256, 256), generator=g, device=dev) < 0.75).to(torch.uint8)
        free = coarse.repeat_interleave(64, 0).repeat_interleave(64, 1).contiguous()
        cc = sg.ImgCollisionChecker.from_free_mask(free)
    scale = torch.tensor([W, H], device=dev, dtype=torch.float64)
    n = 1 << 22
    P = torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * scale
    Q = torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) * scale
    if what == "visible_l2":
        Q = P + (torch.rand((n, 2), generator=g, device=dev, dtype=torch.float64) - 0.5) * 160
    ms = timeit(lambda: cc.visible_batch(P, Q))
    print(what, "ms", ms, "edges/s", n / ms * 1e3)
elif what == "arm":
    free = load_map("4d")
    cc = sg.RobotArm4dCollisionChecker.__new__(sg.RobotArm4dCollisionChecker)
    sg.CollisionChecker.__init__(cc)
    cc.image_fname = None
    cc.stick_robot_length_config = [30.0, 30.0]
    cc._attach(free.astype(np.uint8), 0)
    n = 1 << 18
    lo = torch.tensor([0, 0, -np.pi, -np.pi], device=dev, dtype=torch.float64)
    hi = torch.tensor([1024, 1024, np.pi, np.pi], device=dev, dtype=torch.float64)
    C1 = torch.rand((n, 4), generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo
    C2 = C1 + torch.randn((n, 4), generator=g, device=dev, dtype=torch.float64) * torch.tensor([6, 6, .3, .3], device=dev)
    ms = timeit(lambda: cc._visible_device(C1, C2, resolve=False))
    print(what, "ms", ms, "edges/s", n / ms * 1e3)
