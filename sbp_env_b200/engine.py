"""Engines that own a GPU-backed checker: the construction boundary of the path.

Mirrors ``Engine`` / ``ImageEngine`` / ``RobotArmEngine`` of
/root/reference/sbp_env/engine.py:13-120 (``cc``, ``dist``, ``transform``,
``lower`` / ``upper``, ``get_dimension``).  Pass one to the reference's
``generate_args(engine=...)`` (__entry_point.py:120-158) and the unmodified planners
and samplers run on the GPU checker.  ``KlamptEngine`` / ``BlackBoxEngine`` are out
of scope.  The reference's constructors also select the
pygame visualiser (engine.py:75, :105); display is out of scope here.
"""
from __future__ import annotations

import math

import numpy as np

from . import collision_checker


def euclidean_dist(p1: np.ndarray, p2: np.ndarray):
    """Distance over the FIRST TWO coordinates only, like engine.py:13-24."""
    p = p1 - p2
    return math.sqrt(p[0] ** 2 + p[1] ** 2)


class Engine:
    def __init__(self, cc: collision_checker.CollisionChecker):
        self.cc = cc

    def dist(self, *args):
        return euclidean_dist(*args)

    def transform(self, qs: np.ndarray) -> np.ndarray:
        """Project samples in [0, 1) onto the bounds (engine.py:50-57)."""
        return qs * (self.upper - self.lower) + self.lower

    def get_dimension(self) -> int:
        raise NotImplementedError("Must derive from this class")


class ImageEngine(Engine):
    def __init__(self, args, image, device: int = 0):
        super().__init__(collision_checker.ImgCollisionChecker(image, device=device))

    @property
    def lower(self) -> np.ndarray:
        return np.array([0, 0])

    @property
    def upper(self) -> np.ndarray:
        return np.array(self.cc._img.shape)

    def get_dimension(self) -> int:
        return 2


class RobotArmEngine(Engine):
    def __init__(self, args, image, device: int = 0):
        super().__init__(collision_checker.RobotArm4dCollisionChecker(
            image, stick_robot_length_config=args.rover_arm_robot_lengths, device=device))

    @property
    def lower(self) -> np.ndarray:
        return np.array([0, 0, -np.pi, -np.pi])

    @property
    def upper(self) -> np.ndarray:
        return np.array([self.cc._img.shape[0], self.cc._img.shape[1], np.pi, np.pi])

    def get_dimension(self) -> int:
        return 4
