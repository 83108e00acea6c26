"""-m gpu: 4-DoF arm checker on the GPU against golden fixtures and the oracle."""
import numpy as np
import pytest

import c_oracle as orc
from _util import golden, kat_png, load_map, map_png, unpack
from sbp_env_b200.runtime import tune  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


def pt(*a):
    return np.array(a, float)


def test_reference_kat(sg):
    # tests/test_4d_arm_collision_checker.py:23-75 of the reference
    cc = sg.RobotArm4dCollisionChecker(kat_png(), stick_robot_length_config=[0.1, 0.1])
    s, e = np.array([1, 5, 3, 19]), np.array([-11, 54, 32, -219])
    r = cc.create_ranges(s, e, 10)
    assert r.shape == (4, 10) and np.isclose(r[:, 0], s).all() and np.isclose(r[:, -1], e).all()
    a, b = np.random.rand(4) * 10, np.random.rand(4) * 10
    out = cc._interpolate_configs(a, b)
    assert out.shape[0] >= 2 and out.shape[1] == 4
    assert np.isclose(out[0, :2].astype(np.uint8), a[:2].astype(np.uint8)).all()
    assert np.isclose(out[0, 2:], a[2:]).all() and np.isclose(out[-1, 2:], b[2:]).all()
    assert cc.visible(pt(0.5, 2.5, 1, 2.5), pt(1.5, 3.5, 0, -1.5))
    assert cc.visible(pt(0.5, 2.5, -1, 1), pt(0.5, 3.5, 2, 0.4))
    assert not cc.visible(pt(0.5, 2.5, 1, 2.5), pt(1.5, 5.5, 0, -1.5))
    assert not cc.visible(pt(0.5, 0.5, -1, 1), pt(3.5, 0.5, 2, 0.4))
    for c in [pt(0.5, 2.5, 1, 2.5), pt(1.5, 3.5, 0, -1.5), pt(0.5, 3.5, 2, 0.4)]:
        assert cc.feasible(c)
    for c in [(3.5, 0.5, 1, 1), (-1.5, 3.5, 1, 1), (1.78, 4.5, 1, 1)]:
        assert not cc.feasible(c)
    assert cc._pt_feasible((1.5, 0.5)) and not cc._pt_feasible((3.5, 0.5))
    with pytest.raises(ValueError):
        cc.feasible((float("nan"), 0.5, 0, 0))
    with pytest.raises(ValueError):
        cc.visible(pt(0.5, 2.5, float("inf"), 0), pt(1.5, 3.5, 0, 0))   # math domain error


def test_golden_arm(sg):
    g = golden("arm_cc")
    for ci in range(int(g["n_cases"])):
        t = f"c{ci}"
        name = str(g[t + "_map"])
        L = g[t + "_L"].tolist()
        cc = sg.RobotArm4dCollisionChecker(kat_png() if name == "kat4x6" else map_png(name),
                                           stick_robot_length_config=L)
        C = g[t + "_C"]
        assert np.array_equal(cc.feasible_batch(C), unpack(g[t + "_feasible"], len(C))), t
        E1, E2 = g[t + "_E1"], g[t + "_E2"]
        assert np.array_equal(cc.visible_batch(E1, E2), unpack(g[t + "_visible"], len(E1))), t


def test_interpolate_configs_golden(sg):
    g = golden("arm_cc")
    cc = sg.RobotArm4dCollisionChecker(kat_png(), stick_robot_length_config=[0.1, 0.1])
    off = g["ic_off"]
    for i, (a, b) in enumerate(zip(g["ic_a"], g["ic_b"])):
        assert np.array_equal(cc._interpolate_configs(a, b), g["ic_val"][off[i]:off[i + 1]])
    off = g["cr_off"]
    for i, (s, e, N) in enumerate(zip(g["cr_start"], g["cr_stop"], g["cr_N"])):
        assert np.array_equal(cc.create_ranges(s, e, int(N)).reshape(-1), g["cr_val"][off[i]:off[i + 1]])


@pytest.mark.parametrize("name,L", [("4d", (30.0, 30.0)), ("maze1", (12.0, 3.0)), ("room1", (45.5, 20.25))])
def test_random_vs_oracle(sg, name, L):
    free = load_map(name)
    W, H = free.shape
    cc = sg.RobotArm4dCollisionChecker(map_png(name), stick_robot_length_config=list(L))
    om = orc.Map(free)
    rng = np.random.default_rng(7)
    lo, hi = np.array([0, 0, -np.pi, -np.pi]), np.array([W, H, np.pi, np.pi])
    n = 60_000
    C1 = rng.random((n, 4)) * (hi - lo) + lo
    C1[:2000, :2] = (rng.random((2000, 2)) * 1.6 - 0.3) * [W, H]       # partly outside
    C2 = C1 + rng.standard_normal((n, 4)) * [7, 7, 0.4, 0.4]
    C2[-3000:] = rng.random((3000, 4)) * (hi - lo) + lo                 # long edges
    assert np.array_equal(cc.feasible_batch(C1), om.arm_feasible(C1, *L))
    from sbp_env_b200 import _lib

    lib = _lib.load()
    want_fwd, want_rev = om.arm_visible(C1, C2, *L), om.arm_visible(C2, C1, *L)
    try:
        # (flat, group): the flattened (edge, config) kernel (default), then the lane-group kernel with
        # 8 (its default), 4 and 32 lanes per edge
        # (flat = 2: the flattened kernel with the free-box test of the links, the default)
        for flat, group in ((2, 0), (1, 0), (0, 0), (0, 4), (0, 32)):
            tune("arm_flat", flat)
            tune("arm_group", group)
            assert np.array_equal(cc.feasible_batch(C1), om.arm_feasible(C1, *L)), (flat, group)
            assert np.array_equal(cc.feasible_batch(C2[-3000:]), om.arm_feasible(C2[-3000:], *L)), (flat, group)
            assert np.array_equal(cc.visible_batch(C1, C2), want_fwd), group
            # H8: argument order matters for the arm; check the reverse direction too
            assert np.array_equal(cc.visible_batch(C2, C1), want_rev), group
            # tiny batches (fewer edges than one warp) and a single edge
            assert np.array_equal(cc.visible_batch(C1[:37], C2[:37]), want_fwd[:37]), group
            assert np.array_equal(cc.visible_batch(C1[-1:], C2[-1:]), want_fwd[-1:]), group
    finally:
        tune("arm_group", 0)
        tune("arm_flat", 2)


def test_guard_band_is_resolved_exactly(sg):
    """Axis-aligned angles on integer bases put every FK point ON a pixel boundary: the
    device must decline (SBP_AMBIGUOUS) and the host libm resolution must match."""
    import torch

    free = load_map("4d")
    W, H = free.shape
    cc = sg.RobotArm4dCollisionChecker(map_png("4d"), stick_robot_length_config=[30.0, 30.0])
    om = orc.Map(free)
    rng = np.random.default_rng(3)
    n = 5000
    C1 = np.zeros((n, 4))
    C1[:, :2] = np.floor(rng.random((n, 2)) * [W, H])
    C1[:, 2] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi], n)
    C1[:, 3] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi], n)
    C2 = C1.copy()
    C2[:, :2] += np.floor(rng.standard_normal((n, 2)) * 5)
    assert np.array_equal(cc.feasible_batch(C1), om.arm_feasible(C1, 30.0, 30.0))
    assert cc.last_ambiguous > 0
    assert np.array_equal(cc.visible_batch(C1, C2), om.arm_visible(C1, C2, 30.0, 30.0))
    assert cc.last_ambiguous > 0
    got = cc.visible_batch(torch.from_numpy(C1).cuda(), torch.from_numpy(C2).cuda()).cpu().numpy()
    assert np.array_equal(got, om.arm_visible(C1, C2, 30.0, 30.0))
    got = cc.feasible_batch(torch.from_numpy(C1).cuda()).cpu().numpy()       # device-tensor path
    assert np.array_equal(got, om.arm_feasible(C1, 30.0, 30.0))
    # uniform random configs almost never land in the band
    U = rng.random((200000, 4)) * [W, H, 6, 6] - [0, 0, 3, 3]
    cc.feasible_batch(U)
    assert cc.last_ambiguous <= 5


def test_arm_prescreen_bucket(sg):
    """cc.prescreen for the 4-D arm checker: one launch decides a whole bucket of configurations,
    feasible(c) of a row is answered from the remembered booleans (guard-band rows included)."""
    free = load_map("4d")
    W, H = free.shape
    cc = sg.RobotArm4dCollisionChecker(map_png("4d"), stick_robot_length_config=[30.0, 30.0])
    om = orc.Map(free)
    rng = np.random.default_rng(17)
    bucket = rng.random((10000, 4)) * [W, H, 2 * np.pi, 2 * np.pi] - [0, 0, np.pi, np.pi]
    bucket[:50, :2] = np.floor(bucket[:50, :2])
    bucket[:50, 2:] = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi], (50, 2))     # on pixel boundaries
    want = om.arm_feasible(bucket, 30.0, 30.0)
    launches = []
    inner = cc._feasible_host
    cc._feasible_host = lambda C: (launches.append(len(C)), inner(C))[1]
    try:
        assert np.array_equal(cc.prescreen(bucket), want) and launches == [10000]
        for i in (0, 7, 49, 5000, 9999):
            assert bool(cc.feasible(bucket[i])) == bool(want[i])
        assert launches == [10000]
        assert bool(cc.feasible([100.5, 200.25, 0.3, -1.2])) == bool(om.arm_feasible(np.array([[100.5, 200.25, 0.3, -1.2]]), 30.0, 30.0)[0])
        assert launches == [10000, 1]
    finally:
        del cc._feasible_host
