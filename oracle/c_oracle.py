"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of ``oracle/oracle.c``.

The functions mirror the reference's call sites (file:line under
/root/reference/sbp_env/) but take whole numpy batches.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

STATUS_OK, STATUS_NAN, STATUS_INF = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile oracle.c (gcc, -ffp-contract=off) when missing or stale."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i64, dbl, vp, i32 = C.c_int64, C.c_double, C.c_void_p, C.c_int
        L.orc_set_threads.argtypes = [i32]
        L.orc_max_threads.restype = i32
        L.orc_feasible2d.argtypes = [vp, i64, i64, vp, i64, vp, vp]
        L.orc_line_len.argtypes = [i64] * 4
        L.orc_line_len.restype = i64
        L.orc_get_line.argtypes = [i64] * 4 + [vp]
        L.orc_visible2d.argtypes = [vp, i64, i64, vp, vp, i64, vp, vp, vp, vp]
        L.orc_coor_before_collision.argtypes = [vp, i64, i64, vp, vp, i64, vp, vp]
        L.orc_arm_feasible.argtypes = [vp, i64, i64, dbl, dbl, vp, i64, vp, vp]
        L.orc_arm_visible.argtypes = [vp, i64, i64, dbl, dbl, vp, vp, i64, vp, vp, vp, vp]
        L.orc_create_ranges.argtypes = [vp, vp, i64, i64, vp]
        L.orc_nearest.argtypes = [vp, i64, i32, vp, i64, vp, vp]
        L.orc_knn.argtypes = [vp, i64, i32, vp, i64, i32, vp, vp]
        L.orc_near.argtypes = [vp, i64, i32, vp, i64, dbl, i32, i32, vp, vp]
        L.orc_dist_matrix.argtypes = [vp, i64, i32, vp, i64, i32, vp]
        _lib = L
    return _lib


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _img(free_xy):
    """free_xy: the reference's ``_img`` ([W][H], truthy == 1.0 means free)."""
    a = np.ascontiguousarray((np.asarray(free_xy) == 1).astype(np.uint8))
    assert a.ndim == 2
    return a


class Map:
    """Dense uint8 copy of a reference checker's ``_img`` ([W][H], 1 = free)."""

    def __init__(self, free_xy):
        self.img = _img(free_xy)
        self.W, self.H = self.img.shape

    # ImgCollisionChecker.feasible, collisionChecker.py:147-153
    def feasible2d(self, P, return_status=False):
        P = _f64(P, 2)
        n = len(P)
        out = np.zeros(n, np.uint8)
        st = np.zeros(n, np.uint8)
        lib().orc_feasible2d(_p(self.img), self.W, self.H, _p(P), n, _p(out), _p(st))
        return (out.astype(bool), st) if return_status else out.astype(bool)

    # ImgCollisionChecker.visible, collisionChecker.py:134-145
    def visible2d(self, P, Q, return_status=False, return_counts=False):
        P, Q = _f64(P, 2), _f64(Q, 2)
        n = len(P)
        assert len(Q) == n
        out = np.zeros(n, np.uint8)
        st = np.zeros(n, np.uint8)
        full = np.zeros(n, np.int64) if return_counts else None
        tested = np.zeros(n, np.int64) if return_counts else None
        lib().orc_visible2d(_p(self.img), self.W, self.H, _p(P), _p(Q), n, _p(out),
                            _p(st), _p(full), _p(tested))
        res = [out.astype(bool)]
        if return_status:
            res.append(st)
        if return_counts:
            res += [full, tested]
        return res[0] if len(res) == 1 else tuple(res)

    # ImgCollisionChecker.get_coor_before_collision, collisionChecker.py:118-132
    def coor_before_collision(self, P, Q):
        P, Q = _f64(P, 2), _f64(Q, 2)
        n = len(P)
        out = np.zeros((n, 2), np.int64)
        st = np.zeros(n, np.uint8)
        lib().orc_coor_before_collision(_p(self.img), self.W, self.H, _p(P), _p(Q), n,
                                        _p(out), _p(st))
        return out

    # RobotArm4dCollisionChecker.feasible, collisionChecker.py:404-426
    def arm_feasible(self, C4, L0, L1, return_status=False):
        C4 = _f64(C4, 4)
        n = len(C4)
        out = np.zeros(n, np.uint8)
        st = np.zeros(n, np.uint8)
        lib().orc_arm_feasible(_p(self.img), self.W, self.H, float(L0), float(L1), _p(C4),
                               n, _p(out), _p(st))
        return (out.astype(bool), st) if return_status else out.astype(bool)

    # RobotArm4dCollisionChecker.visible, collisionChecker.py:385-391
    def arm_visible(self, C1, C2, L0, L1, return_status=False, return_counts=False):
        C1, C2 = _f64(C1, 4), _f64(C2, 4)
        n = len(C1)
        assert len(C2) == n
        out = np.zeros(n, np.uint8)
        st = np.zeros(n, np.uint8)
        ncfg = np.zeros(n, np.int64) if return_counts else None
        tested = np.zeros(n, np.int64) if return_counts else None
        lib().orc_arm_visible(_p(self.img), self.W, self.H, float(L0), float(L1), _p(C1),
                              _p(C2), n, _p(out), _p(st), _p(ncfg), _p(tested))
        res = [out.astype(bool)]
        if return_status:
            res.append(st)
        if return_counts:
            res += [ncfg, tested]
        return res[0] if len(res) == 1 else tuple(res)


def py_trunc(v):
    """CPython ``int(float)`` for finite values (truncation toward zero)."""
    return int(v)


# get_line, collisionChecker.py:155-215 (endpoints are truncated like `map(int, .)`)
def get_line(start, end):
    x1, y1 = (int(v) for v in start)
    x2, y2 = (int(v) for v in end)
    n = lib().orc_line_len(x1, y1, x2, y2)
    out = np.zeros((n, 2), np.int64)
    lib().orc_get_line(x1, y1, x2, y2, _p(out))
    return out


# create_ranges, collisionChecker.py:346-364
def create_ranges(start, stop, N):
    start, stop = _f64(start).ravel(), _f64(stop).ravel()
    out = np.zeros((len(start), N), np.float64)
    lib().orc_create_ranges(_p(start), _p(stop), len(start), int(N), _p(out))
    return out


# find_nearest_neighbour_idx, rrtPlanner.py:268-279
def nearest(poses, X):
    poses = _f64(poses)
    d = poses.shape[1]
    X = _f64(X, d)
    m = len(X)
    idx = np.zeros(m, np.int64)
    dist = np.zeros(m, np.float64)
    lib().orc_nearest(_p(poses), len(poses), d, _p(X), m, _p(idx), _p(dist))
    return idx, dist


def knn(poses, X, k):
    poses = _f64(poses)
    d = poses.shape[1]
    X = _f64(X, d)
    m = len(X)
    idx = np.zeros((m, k), np.int64)
    dist = np.zeros((m, k), np.float64)
    lib().orc_knn(_p(poses), len(poses), d, _p(X), m, int(k), _p(idx), _p(dist))
    return idx, dist


# nearest_neighbours, prmPlanner.py:28-44 (metric="full", closed=False);
# the radius gate of choose_least_cost_parent, rrtPlanner.py:177-181
# (metric="xy", closed=True)
def near(poses, X, r, metric="full", closed=False):
    poses = _f64(poses)
    d = poses.shape[1]
    X = _f64(X, d)
    m = len(X)
    met = {"full": 0, "xy": 1}[metric]
    row_ptr = np.zeros(m + 1, np.int64)
    lib().orc_near(_p(poses), len(poses), d, _p(X), m, float(r), met, int(bool(closed)),
                   _p(row_ptr), None)
    idx = np.zeros(max(int(row_ptr[-1]), 1), np.int64)
    lib().orc_near(_p(poses), len(poses), d, _p(X), m, float(r), met, int(bool(closed)),
                   _p(row_ptr), _p(idx))
    return row_ptr, idx[: int(row_ptr[-1])]


def dist_matrix(poses, X, metric="full"):
    poses = _f64(poses)
    d = poses.shape[1]
    X = _f64(X, d)
    out = np.zeros((len(X), len(poses)), np.float64)
    lib().orc_dist_matrix(_p(poses), len(poses), d, _p(X), len(X),
                          {"full": 0, "xy": 1}[metric], _p(out))
    return out
