"""The planning statistics singleton the checkers report into.

Mirrors ``Stats`` of /root/reference/sbp_env/utils/common.py:195-298.  When the
reference package itself is importable its own ``Stats`` class is used, so that the
reference's Env / tqdm / CSV logger (env.py:191-212) and the GPU checkers count into
the same object; otherwise a local mirror with the same attributes is used.
"""
from __future__ import annotations

import sys


class _LocalStats:
    _instance = None

    @classmethod
    def has_instance(cls) -> bool:
        return cls._instance is not None

    @classmethod
    def build_instance(cls, **kwargs):
        if cls._instance is not None:
            raise ValueError(f"There are already an existing instance of Stats! {cls._instance}")
        cls._instance = cls(**kwargs)
        return cls._instance

    @classmethod
    def get_instance(cls):
        if cls._instance is None:
            raise ValueError("Stats instance had not been built yet!")
        return cls._instance

    @classmethod
    def clear_instance(cls) -> None:
        cls._instance = None

    def __init__(self, showSampledPoint=True):
        self.invalid_samples_connections = 0
        self.invalid_samples_obstacles = 0
        self.valid_sample = 0
        self.sampledNodes = []
        self.showSampledPoint = showSampledPoint
        self.sampler_success = 0
        self.sampler_success_all = 0
        self.sampler_fail = 0
        self.visible_cnt = 0
        self.feasible_cnt = 0
        self.lsampler_restart_counter = 0
        self.lsampler_randomwalk_counter = 0

    def add_invalid(self, obs):
        if obs:
            self.invalid_samples_obstacles += 1
        else:
            self.invalid_samples_connections += 1

    def add_free(self):
        self.valid_sample += 1

    def add_sampled_node(self, pos):
        if not self.showSampledPoint:
            return
        self.sampledNodes.append(pos)

    def __repr__(self):
        return "Stats<{}>".format("|".join(
            f"{a}={getattr(self, a)}" for a in dir(self)
            if not a.startswith("_") and not callable(getattr(self, a))))


def _resolve():
    mod = sys.modules.get("sbp_env.utils.common")
    if mod is None:
        try:
            import sbp_env.utils.common as mod  # noqa: F811 - the installed reference
        except Exception:
            mod = None
    return mod.Stats if mod is not None and hasattr(mod, "Stats") else _LocalStats


class Stats:
    """Facade over whichever ``Stats`` class is live (reference's or the mirror)."""

    @staticmethod
    def has_instance() -> bool:
        return _resolve().has_instance()

    @staticmethod
    def build_instance(**kwargs):
        return _resolve().build_instance(**kwargs)

    @staticmethod
    def get_instance():
        return _resolve().get_instance()

    @staticmethod
    def clear_instance() -> None:
        _resolve().clear_instance()
