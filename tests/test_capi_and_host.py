"""CPU-only: the C-ABI library loads and exports every symbol include/sbp_gpu.h
declares; host-side logic (Bresenham enumeration, Stats mirror, sharding helpers);
and the product path fails loudly -- never falls back -- without a CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

from _util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sbp_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from sbp_env_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in sbp_gpu.h but not exported"
    # and the Python binding table covers exactly the declared ABI
    assert sorted(_lib.ABI_SYMBOLS) == declared
    assert _lib.load().sbp_abi_version() == 2


def test_no_cuda_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import sbp_env_b200 as sg

    with pytest.raises(sg.SbpGpuError, match="no CPU path"):
        sg.NodeStore(2)
    with pytest.raises(sg.SbpGpuError):
        sg.ImgCollisionChecker.from_free_mask(np.ones((4, 4), np.uint8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sbp_env_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "c_oracle" not in src and "import oracle" not in src and "liboracle" not in src, fn
    for fn in os.listdir(os.path.join(pkg, "csrc")):
        if fn.endswith((".cu", ".cuh")):
            assert "oracle" not in open(os.path.join(pkg, "csrc", fn)).read().lower(), fn


def test_bresenham_enumeration_matches_golden():
    from sbp_env_b200 import ImgCollisionChecker, RobotArm4dCollisionChecker

    g = golden("img_cc")
    off = g["line_off"]
    for i, (s, e) in enumerate(zip(g["line_S"], g["line_E"])):
        exp = g["line_xy"][off[i]:off[i + 1]]
        assert ImgCollisionChecker.get_line(s, e) == [tuple(p) for p in exp.tolist()]
        assert RobotArm4dCollisionChecker._get_line(s, e) == exp.tolist()
    assert ImgCollisionChecker.get_line((0, 0), (3, 4)) == [(0, 0), (1, 1), (1, 2), (2, 3), (3, 4)]


def test_create_ranges_bit_exact():
    from sbp_env_b200 import RobotArm4dCollisionChecker

    g = golden("arm_cc")
    off = g["cr_off"]
    for i, (s, e, N) in enumerate(zip(g["cr_start"], g["cr_stop"], g["cr_N"])):
        got = RobotArm4dCollisionChecker.create_ranges(s, e, int(N)).reshape(-1)
        assert np.array_equal(got, g["cr_val"][off[i]:off[i + 1]])
    assert RobotArm4dCollisionChecker.get_pt_from_angle_and_length((1.0, 2.0), 0.0, 3.0) == (4.0, 2.0)


def test_stats_mirror():
    from sbp_env_b200.stats import Stats, _LocalStats

    Stats.clear_instance()
    assert not Stats.has_instance()
    s = Stats.build_instance()
    assert Stats.get_instance() is s
    with pytest.raises(ValueError):
        Stats.build_instance()
    s.add_free(); s.add_invalid(True); s.add_invalid(False)
    assert (s.valid_sample, s.invalid_samples_obstacles, s.invalid_samples_connections) == (1, 1, 1)
    assert "visible_cnt=0" in repr(s)
    Stats.clear_instance()
    with pytest.raises(ValueError):
        Stats.get_instance()
    assert isinstance(s, _LocalStats) or type(s).__name__ == "Stats"


def test_engine_mirror_host_side():
    from sbp_env_b200.engine import Engine, euclidean_dist

    # first two dims only (engine.py:13-24)
    assert euclidean_dist(np.array([0.0, 0, 9, 9]), np.array([3.0, 4, 0, 0])) == 5.0

    class E(Engine):
        lower = np.array([0, 0]); upper = np.array([10, 20])

    e = E(cc=None)
    assert np.array_equal(e.transform(np.array([0.5, 0.5])), [5.0, 10.0])


def test_shard_helpers():
    from sbp_env_b200.distributed import shard_by_cost, shard_range

    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    costs = np.random.default_rng(0).integers(1, 500, 10000)
    b = shard_by_cost(costs, 4)
    assert b[0] == 0 and b[-1] == len(costs) and (np.diff(b) > 0).all()
    loads = [costs[b[i]:b[i + 1]].sum() for i in range(4)]
    assert max(loads) - min(loads) <= 2 * costs.max()
