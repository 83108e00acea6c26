"""Device-resident RRT* tree: the node store plus ``cost`` / ``parent``
arrays on the device, and the inner loop of one accepted sample -- ``choose_least_cost_parent``
(linear path, rrtPlanner.py:173-196), ``add_newnode`` (:106-113) and the linear ``rewire``
(:250-266) -- as ONE kernel launch whose outcome comes back in a pinned struct
(``sbp_tree_extend_host``).

Exactness contract.  The reference's costs are sums of ``engine.dist`` values (libm ``pow``), which
may differ from the device's correctly rounded distance in the last bit.  The device therefore only
takes comparisons whose sides are more than 64 ulp apart; otherwise it reports the sample as
ambiguous without writing anything and the sample is decided by the host path
(:mod:`planner_ops`), its result pushed to the device arrays.  The host keeps the reference's exact
costs (one ``engine.dist`` call per new node and per rewired node) and hands values that differ
from the device's to the next call as fixes, so both sides hold the reference's costs bit for bit.
For the 2-D image checker; the arm checker stays on the host path (guard-band resolution).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib, planner_ops
from ._lib import check
from .engine import euclidean_dist


class DeviceTree:
    def __init__(self, cc, store, dist=euclidean_dist):
        if store.dim != 2 or hasattr(cc, "stick_robot_length_config"):
            raise ValueError("DeviceTree: 2-D image checker only")
        self.cc, self.store, self.dist = cc, store, dist
        h = C.c_void_p()
        check(_lib.load().sbp_tree_create(store._h, C.byref(h)), "sbp_tree_create")
        self._h = h
        self.costs, self.parents, self.poses = [], [], []
        self._fixes = {}
        self.ambiguous = 0                       # samples settled by the host path so far

    def __len__(self):
        return len(self.costs)

    def add_root(self, pos, cost: float = 0.0):
        """Append a node without a parent (the start node, rrtPlanner.py init)."""
        pos = np.asarray(pos, dtype=np.float64)
        idx = self.store.append(pos)
        assert idx == len(self.costs)
        self.costs.append(float(cost))
        self.parents.append(-1)
        self.poses.append(pos.copy())
        check(_lib.load().sbp_tree_set(self._h, idx, float(cost), -1), "sbp_tree_set")
        return idx

    @classmethod
    def from_arrays(cls, cc, store, poses, costs, parents, dist=euclidean_dist):
        """An existing tree: ``store`` already holds ``poses`` [n][2]; costs / parents are uploaded."""
        t = cls(cc, store, dist=dist)
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        assert len(store) == len(poses) == len(costs) == len(parents)
        c = np.ascontiguousarray(costs, dtype=np.float64)
        p = np.ascontiguousarray(parents, dtype=np.int32)
        check(_lib.load().sbp_tree_write(t._h, 0, len(c), c.ctypes.data, p.ctypes.data), "sbp_tree_write")
        t.costs, t.parents, t.poses = c.tolist(), p.tolist(), list(poses)
        return t

    def _push(self, idx, cost, parent):
        check(_lib.load().sbp_tree_set(self._h, int(idx), float(cost), int(parent)), "sbp_tree_set")

    def extend(self, newpos, nn, radius, rewire: bool = False, append: bool = True):
        """One accepted sample.  Returns ``(parent, cost, rewired)``: the chosen parent, the new
        node's cost (the reference's value) and the indices re-parented to the new node, ascending.
        ``nn``: index of the current nearest node or ``None``."""
        if rewire and not append:
            raise ValueError("rewire needs the new node in the tree (append=True)")
        newpos = np.ascontiguousarray(newpos, dtype=np.float64)
        lib = _lib.load()
        self.store.ctx.use_torch_stream()
        while len(self._fixes) > _lib.EXTEND_MAX_FIXES:          # (more fixes than one call takes)
            i, c = self._fixes.popitem()
            self._push(i, c, -2)
        fi = (C.c_int64 * max(1, len(self._fixes)))(*self._fixes.keys())
        fc = (C.c_double * max(1, len(self._fixes)))(*self._fixes.values())
        res = _lib.ExtendResult()
        check(lib.sbp_tree_extend_host(self._h, self.cc.device_map.handle, newpos.ctypes.data,
                                       -1 if nn is None else int(nn), float(radius), int(append), int(rewire),
                                       fi, fc, len(self._fixes), C.byref(res)), "sbp_tree_extend_host")
        self._fixes.clear()
        if res.status == _lib.EXTEND_NO_PARENT:
            raise LookupError("ERROR: Provided nn=None, and cannot find any valid nn by this function. "
                              "This newnode is not close to the root tree...?")
        if res.status == _lib.EXTEND_AMBIGUOUS:
            return self._extend_on_host(newpos, nn, radius, rewire, append)
        # Stats: one visible() per node inside the radius (rrtPlanner.py:179-181)
        self.cc.stats.visible_cnt += int(res.n_candidates)
        parent = int(res.parent)
        cost = self.costs[parent] + self.dist(self.poses[parent], newpos)        # rrtPlanner.py:192-194
        rewired = []
        if append:
            new = int(res.new_idx)
            assert new == len(self.costs)
            self.costs.append(cost)
            self.parents.append(parent)
            self.poses.append(newpos.copy())
            if cost != res.cost:
                self._fixes[new] = cost
            if rewire:
                order = sorted(range(res.n_rewired), key=lambda j: res.rewired[j])
                # the reference evaluates visible(n, new) for every node inside the radius that is
                # not at the new node's or its parent's position (rrtPlanner.py:256-260)
                self.cc.stats.visible_cnt += int(res.n_rewire_checked)
                for j in order:
                    n = int(res.rewired[j])
                    c = self.dist(newpos, self.poses[n])
                    exact = cost + c
                    self.costs[n] = exact
                    self.parents[n] = new
                    if exact != res.cost + res.rewired_c[j]:
                        self._fixes[n] = exact
                    rewired.append(n)
        return parent, cost, rewired

    def _extend_on_host(self, newpos, nn, radius, rewire, append):
        self.ambiguous += 1
        P = np.asarray(self.poses)
        parent, cost, cand = planner_ops.choose_least_cost_parent(self.cc, self.store, self.costs, newpos, nn, radius,
                                                                  dist=self.dist, poses=P)
        rewired = []
        if append:
            new = self.store.append(newpos)
            self.costs.append(cost)
            self.parents.append(parent)
            self.poses.append(newpos.copy())
            self._push(new, cost, parent)
            if rewire:
                rewired = planner_ops.rewire(self.cc, self.store, self.costs, self.parents, new, radius,
                                             dist=self.dist, poses=np.asarray(self.poses), candidates=cand)
                for n in rewired:
                    self._push(n, self.costs[n], new)
        return parent, cost, rewired

    def read_device(self):
        """``(cost fp64 [n], parent int32 [n])`` as the device holds them (pending fixes applied first)."""
        for i, c in self._fixes.items():
            self._push(i, c, -2)
        self._fixes.clear()
        n = len(self.costs)
        cost, par = np.empty(n, np.float64), np.empty(n, np.int32)
        check(_lib.load().sbp_tree_read(self._h, 0, n, cost.ctypes.data, par.ctypes.data), "sbp_tree_read")
        return cost, par

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.load().sbp_tree_destroy(self._h)
                self._h = None
        except Exception:
            pass
