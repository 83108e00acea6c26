"""Pins the oracle (oracle/oracle.c) to the reference.

(i) the reference's own known-answer tests, restated here with their file:line,
(ii) tests/golden/*.npz, produced by oracle/gen_golden.py from the imported,
unmodified reference.  CPU only.
"""
import numpy as np
import pytest

import c_oracle as orc
from _util import golden, load_map, unpack

EPS = 1e-5


def _pair(a, b):
    # reference helper tests/test_image_space_collision_checker.py:17-22
    return np.array(a, float) + EPS, np.array(b, float) - EPS


@pytest.fixture(scope="module")
def kat():
    return orc.Map(load_map("kat4x6"))


def test_kat_binarisation(kat):
    # tests/test_image_space_collision_checker.py:33-34 -- only grey == 255 is free
    expect = (np.array([[255, 255, 0, 0, 100, 200]] * 4) == 255).T
    assert np.array_equal(kat.img.astype(bool), expect)


def test_kat_img_visible_feasible(kat):
    # tests/test_image_space_collision_checker.py:44-60
    P, Q = zip(_pair([0, 2], [2, 4]), _pair([0, 2], [0, 4]), _pair([0, 2], [2, 6]), _pair([0, 0], [4, 0]))
    assert kat.visible2d(np.array(P), np.array(Q)).tolist() == [True, True, False, False]
    pts = [(1.5, 0.5), (1.5, 3.5), (1.78, 2.5), (3.5, 0.5), (-1.5, 3.5), (1.78, 4.5)]
    assert kat.feasible2d(np.array(pts)).tolist() == [True, True, True, False, False, False]


def test_kat_coor_before_collision(kat):
    # tests/test_image_space_collision_checker.py:36-42
    P, Q = zip(_pair([0, 2], [2, 6]), _pair([0, 2], [2, 4]))
    assert kat.coor_before_collision(np.array(P), np.array(Q)).tolist() == [[1, 4], [1, 3]]


def test_kat_get_line_docstring():
    # collisionChecker.py:163-169
    assert orc.get_line((0, 0), (3, 4)).tolist() == [[0, 0], [1, 1], [1, 2], [2, 3], [3, 4]]
    assert orc.get_line((3, 4), (0, 0)).tolist() == [[3, 4], [2, 3], [1, 2], [1, 1], [0, 0]]


def test_kat_arm(kat):
    # tests/test_4d_arm_collision_checker.py:59-75, link lengths 0.1
    pt = lambda *a: np.array(a, float)
    vis = [(pt(0.5, 2.5, 1, 2.5), pt(1.5, 3.5, 0, -1.5)), (pt(0.5, 2.5, -1, 1), pt(0.5, 3.5, 2, 0.4)),
           (pt(0.5, 2.5, 1, 2.5), pt(1.5, 5.5, 0, -1.5)), (pt(0.5, 0.5, -1, 1), pt(3.5, 0.5, 2, 0.4))]
    got = kat.arm_visible(np.array([a for a, _ in vis]), np.array([b for _, b in vis]), 0.1, 0.1)
    assert got.tolist() == [True, True, False, False]
    cfg = [pt(0.5, 2.5, 1, 2.5), pt(1.5, 3.5, 0, -1.5), pt(0.5, 3.5, 2, 0.4),
           pt(3.5, 0.5, 1, 1), pt(-1.5, 3.5, 1, 1), pt(1.78, 4.5, 1, 1)]
    assert kat.arm_feasible(np.array(cfg), 0.1, 0.1).tolist() == [True, True, True, False, False, False]


def test_kat_create_ranges():
    # tests/test_4d_arm_collision_checker.py:23-34
    s, e = np.array([1, 5, 3, 19.0]), np.array([-11, 54, 32, -219.0])
    r = orc.create_ranges(s, e, 10)
    assert r.shape == (4, 10)
    assert np.isclose(r[:, 0], s).all() and np.isclose(r[:, -1], e).all()


def test_kat_nearest():
    # tests/test_rrtPlanner.py:136-148
    poses = np.array([[0, 0], [1, 1], [2, 2], [3, 3.0]])
    idx, _ = orc.nearest(poses, poses + 1e-2)
    assert idx.tolist() == [0, 1, 2, 3]


@pytest.mark.parametrize("name", ["room1", "intel_lab", "maze1", "noise", "kat4x6"])
def test_golden_img(name):
    g = golden("img_cc")
    m = orc.Map(load_map(name))
    P, Q = g[name + "_P"], g[name + "_Q"]
    vis, st = m.visible2d(P, Q, return_status=True)
    assert np.array_equal(vis, unpack(g[name + "_visible"], len(P)))
    nan_rows = ~(np.isfinite(P).all(1) & np.isfinite(Q).all(1))
    assert np.array_equal(st == orc.STATUS_NAN, nan_rows)
    pts = g[name + "_pts"]
    feas, st = m.feasible2d(pts, return_status=True)
    assert np.array_equal(st, g[name + "_feasible_status"])
    ok = st == 0
    assert np.array_equal(feas[ok], unpack(g[name + "_feasible"], len(pts))[ok])
    sel = g[name + "_cbc_sel"]
    assert np.array_equal(m.coor_before_collision(P[sel], Q[sel]), g[name + "_cbc"])


def test_golden_get_line():
    g = golden("img_cc")
    off = g["line_off"]
    for i, (s, e) in enumerate(zip(g["line_S"], g["line_E"])):
        assert np.array_equal(orc.get_line(s, e), g["line_xy"][off[i]:off[i + 1]]), (s, e)


def test_golden_arm():
    g = golden("arm_cc")
    for ci in range(int(g["n_cases"])):
        t = f"c{ci}"
        m = orc.Map(load_map(str(g[t + "_map"])))
        L0, L1 = g[t + "_L"]
        C = g[t + "_C"]
        assert np.array_equal(m.arm_feasible(C, L0, L1), unpack(g[t + "_feasible"], len(C))), t
        E1, E2 = g[t + "_E1"], g[t + "_E2"]
        assert np.array_equal(m.arm_visible(E1, E2, L0, L1), unpack(g[t + "_visible"], len(E1))), t


def test_golden_create_ranges_bit_exact():
    g = golden("arm_cc")
    off = g["cr_off"]
    for i, (s, e, N) in enumerate(zip(g["cr_start"], g["cr_stop"], g["cr_N"])):
        got = orc.create_ranges(s, e, int(N)).reshape(-1)
        assert np.array_equal(got, g["cr_val"][off[i]:off[i + 1]])   # bit-exact fp64


@pytest.mark.parametrize("d", [2, 3, 4])
def test_golden_nn(d):
    g = golden("nn")
    poses, X = g[f"d{d}_poses"], g[f"d{d}_X"]
    idx, dist = orc.nearest(poses, X)
    assert np.array_equal(idx, g[f"d{d}_nearest"])
    kidx, kdist = orc.knn(poses, X, 20)
    assert np.array_equal(kidx, g[f"d{d}_knn20"])
    assert np.array_equal(kdist, g[f"d{d}_knn20_dist"])              # bit-exact fp64
    assert np.array_equal(orc.dist_matrix(poses, X[:8]), g[f"d{d}_dist8"])
    for ri in (0, 1):
        rp, ix = orc.near(poses, X, float(g[f"d{d}_prm_r{ri}"]), "full", closed=False)
        assert np.array_equal(rp, g[f"d{d}_prm_r{ri}_off"])
        assert np.array_equal(ix, g[f"d{d}_prm_r{ri}_idx"])
    mm = len(g[f"d{d}_xy_off"]) - 1
    rp, ix = orc.near(poses, X[:mm], float(g[f"d{d}_xy_r"]), "xy", closed=True)
    assert np.array_equal(rp, g[f"d{d}_xy_off"])
    assert np.array_equal(ix, g[f"d{d}_xy_idx"])
    # engine.dist goes through libm pow (engine.py:24): same libm here, so bit-exact
    assert np.array_equal(orc.dist_matrix(poses, X[:8], "xy"), g[f"d{d}_xy_dist8"])
