"""sbp_env_b200 -- the collision-checking / nearest-neighbour hot path of
soraxas/sbp-env on B200 (sm_100a): hand-written CUDA kernels behind a C ABI
(``include/sbp_gpu.h``), reached from Python classes that mirror the reference's
``CollisionChecker`` / ``Engine`` interfaces.  No CPU implementation exists in this
package: without ``libsbpgpu.so`` and a B200 every call raises ``SbpGpuError``.
"""
from ._lib import SbpGpuError  # noqa: F401
from .collision_checker import (CollisionChecker, ImgCollisionChecker,  # noqa: F401
                                RobotArm4dCollisionChecker, bresenham_pixels)
from .engine import Engine, ImageEngine, RobotArmEngine, euclidean_dist  # noqa: F401
from .device_tree import DeviceTree  # noqa: F401
from .node_store import NodeStore  # noqa: F401
from .planning import (CapturedStep, PRMConnectGraph, connect_to_roadmap, feasible_batch,  # noqa: F401
                       knn, near, prm_connect_knn, prm_connect_radius, roadmap_edges_to_host,
                       visible_batch)
from . import distributed, planner_hooks, planner_ops, runtime, samplers  # noqa: F401
from .roadmap import Roadmap, prm_get_solution  # noqa: F401
from .stats import Stats  # noqa: F401

__version__ = "0.1.0"
