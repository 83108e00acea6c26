"""Arm-visibility kernel on two regimes (mostly blocked / mostly free edges) -- run under gpurun."""
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


free = load_map("4d")
cc = sg.RobotArm4dCollisionChecker.__new__(sg.RobotArm4dCollisionChecker)
sg.CollisionChecker.__init__(cc)
cc.image_fname = None
cc.stick_robot_length_config = [30.0, 30.0]
cc._attach(free.astype(np.uint8), 0)
n = 1 << 18
lo = torch.tensor([0, 0, -np.pi, -np.pi], device=dev, dtype=torch.float64)
hi = torch.tensor([1024, 1024, np.pi, np.pi], device=dev, dtype=torch.float64)
U = torch.rand((4 * n, 4), generator=g, device=dev, dtype=torch.float64) * (hi - lo) + lo
ok = cc.feasible_batch(U)
ms = timeit(lambda: cc._feasible_device(U, resolve=False))
print(f"feasible: {4 * n} configs {ms:.4f} ms  {4 * n / ms * 1e3:.3e} configs/s", flush=True)
sets = {"random": U[:n], "feasible_start": U[ok][:n]}
for name, C1 in sets.items():
    m = C1.shape[0]
    C2 = C1 + torch.randn((m, 4), generator=g, device=dev, dtype=torch.float64) * torch.tensor([6, 6, .3, .3], device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    vis = cc._visible_device(C1, C2, cfg_tested=cnt, resolve=False)
    print(name, "edges", m, "visible frac", float((vis == 1).double().mean()), "configs tested", int(cnt.item()), flush=True)
    for group in (8, 4, 16):
        tune("arm_group", group)
        ms = timeit(lambda: cc._visible_device(C1, C2, resolve=False))
        print(f"  group={group}: {ms:.4f} ms  {m / ms * 1e3:.3e} edges/s", flush=True)
