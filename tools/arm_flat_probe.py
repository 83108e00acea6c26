"""Arm edge visibility (1024^2 map, L = 30, 30): RRT-like edges and uniform 4-D pairs, lane-group kernel
(arm_flat = 0) against the flattened kernel (arm_flat = 1) and the one with the free-box test of the links (2).  Run under gpurun."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import sbp_env_b200 as sg  # noqa: E402
from _util import load_map  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(0)
free = load_map("4d")
cc = sg.RobotArm4dCollisionChecker.__new__(sg.RobotArm4dCollisionChecker)
sg.CollisionChecker.__init__(cc)
cc.image_fname = None
cc.stick_robot_length_config = [30.0, 30.0]
cc._attach(free.astype(np.uint8), 0)
lo = torch.tensor([0, 0, -np.pi, -np.pi], device=dev, dtype=torch.float64)
hi = torch.tensor([1024, 1024, np.pi, np.pi], device=dev, dtype=torch.float64)
n = 1 << 20
C1 = torch.rand((n, 4), generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo
C1 = C1[cc._feasible_device(C1, resolve=False) == 1][: 1 << 19].contiguous() if hasattr(cc, "_feasible_device") else C1
sets = {"rrt-like (sigma 6 px, 0.3 rad)": C1 + torch.randn(C1.shape, generator=g, device=dev, dtype=torch.float64) * torch.tensor([6, 6, .3, .3], device=dev),
        "uniform pairs": torch.rand(C1.shape, generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo}
for name, C2 in sets.items():
    C2 = C2.contiguous()
    if name.startswith("uniform"):
        a, b = C1[: 1 << 15].contiguous(), C2[: 1 << 15].contiguous()
    else:
        a, b = C1, C2
    res = {}
    for flat in (0, 1, 2):
        sg.runtime.tune("arm_flat", flat)
        r = cc._visible_device(a, b, resolve=False)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = cc._visible_device(a, b, resolve=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[flat] = r.clone()
        npx = (torch.maximum((a[:, 0].trunc() - b[:, 0].trunc()).abs(), (a[:, 1].trunc() - b[:, 1].trunc()).abs()) + 1).clamp(min=2).sum().item()
        t = sorted(ts)[2]
        print(f"{name:32s} arm_flat {flat}: {t:.3f} ms  {len(a) / t * 1e3:.3e} edges/s  {npx / t * 1e3:.3e} configs/s  visible {float((r == 1).float().mean()):.3f}", flush=True)
    assert torch.equal(res[0], res[1]) and torch.equal(res[0], res[2]), name
sg.runtime.tune("arm_flat", 2)

# feasible_batch: 2^20 uniform configs (62 % infeasible) and the feasible ones among them
Cu = torch.rand((1 << 20, 4), generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo
for name, X in (("feasible, uniform configs", Cu), ("feasible, free configs", C1)):
    res = {}
    for flat in (1, 2):
        sg.runtime.tune("arm_flat", flat)
        r = cc._feasible_device(X, resolve=False)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = cc._feasible_device(X, resolve=False)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[flat] = r.clone()
        t = sorted(ts)[2]
        print(f"{name:32s} arm_flat {flat}: {t:.3f} ms  {len(X) / t * 1e3:.3e} configs/s  free {float((r == 1).float().mean()):.3f}", flush=True)
    assert torch.equal(res[1], res[2]), name
sg.runtime.tune("arm_flat", 2)
