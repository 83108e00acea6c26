"""GPU-backed drop-ins for the reference's collision checkers.

Mirrors ``/root/reference/sbp_env/collisionChecker.py`` for the two in-scope
classes -- same class names, constructor arguments, methods, attributes and error
behaviour -- so an ``Engine`` can own one unchanged (``engine.cc``) and planners /
samplers keep calling ``cc.feasible(p)`` / ``cc.visible(p1, p2)``:

  CollisionChecker             collisionChecker.py:10-35
  ImgCollisionChecker          collisionChecker.py:83-215
  RobotArm4dCollisionChecker   collisionChecker.py:299-503

and adds the batched entry points ``feasible_batch`` / ``visible_batch`` /
``first_blocked_batch``.  Every collision answer -- scalar or batched -- is computed
by the sm_100a kernels of ``libsbpgpu.so``; there is no host implementation to fall
back to (``SbpGpuError`` is raised without the library or a B200).

Out of scope: ``BlackBoxCollisionChecker`` (arbitrary host
functor) and ``KlamptCollisionChecker`` (third-party 3-D simulator).
"""
from __future__ import annotations

import copy
import ctypes as C
import math
import typing

import numpy as np

from . import _lib
from ._lib import check
from .runtime import (DeviceMap, _is_torch_tensor, as_device_f64, as_host_f64, ptr,
                      raise_for_flags)
from .stats import Stats


class CollisionChecker:
    """Abstract collision checker (collisionChecker.py:10-35)."""

    def __init__(self):
        # bind (or create) the process-global Stats at construction (:13-16)
        if not Stats.has_instance():
            Stats.build_instance()
        self.stats = Stats.get_instance()

    def visible(self, pos1, pos2):
        raise NotImplementedError("Must derive from this class")

    def feasible(self, p):
        raise NotImplementedError("Must derive from this class")


def _load_binarised(img) -> np.ndarray:
    """PIL open -> "L" -> /255 -> everything but exact white becomes 0 (:101-105)."""
    from PIL import Image

    image = np.array(Image.open(img).convert("L"))
    image = image / 255
    image[image != 1.0] = 0
    return image


class _GridChecker(CollisionChecker):
    """State shared by both checkers: ``_img`` on the host (API parity: engines read
    ``cc._img.shape``), its bit-packed twin on the device, deepcopy / pickle rules."""

    _device = 0

    def _attach(self, img_xy, device):
        self._img = img_xy
        self._device = int(device)
        self._map = DeviceMap(img_xy, self._device)

    @property
    def image(self) -> np.ndarray:
        """The input image (collisionChecker.py:110-116, :342-344)."""
        return self._img.T

    @property
    def device_map(self) -> DeviceMap:
        if self.__dict__.get("_map") is None:       # after unpickling
            self._map = DeviceMap(self._img, self._device)
        return self._map

    # tests/test_rrtPlanner.py:17 deep-copies the whole PlanningOptions, engine and
    # checker included.  The packed map is immutable: share the device handle, copy
    # the rest like the default protocol would (``stats`` included).
    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_map", "_img"):
                new.__dict__[k] = v
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_map"] = None                            # raw device handles never travel
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)

    def _mocked(self, name: str) -> bool:
        # the reference's planner tests assign MagicMocks on the INSTANCE
        # (tests/test_rrtPlanner.py:31-33); batched calls must honour them
        return name in self.__dict__

    # ---- sampler pre-screening ------------------------------
    _dim = 2

    def prescreen(self, P) -> np.ndarray:
        """Decide a whole bucket of candidate states with ONE ``feasible_batch`` launch and
        remember the answers: a later ``feasible(p)`` whose ``p`` is bit-identical to a row
        of ``P`` returns the remembered value without touching the device.

        The reference's samplers draw 10 000 states at a time (``RandomnessManager``,
        randomness.py:150-190) and then test them one by one
        (``Sampler.get_valid_next_pos``, samplers/baseSampler.py:58-66): pre-screening the
        bucket when it is drawn (:func:`sbp_env_b200.samplers.attach_prescreen`) turns
        ~2 launches per accepted node into one per 10 000 draws, with the same draws in the
        same order and the same ``Stats`` counts.  Returns the bool array; a bucket with
        non-finite rows is not remembered (``feasible`` must raise for those rows)."""
        P = as_host_f64(P, self._dim)
        try:
            ok = self._feasible_host(P)
        except (ValueError, OverflowError):
            self.__dict__.pop("_prescreened", None)
            raise
        self._prescreened = {row.tobytes(): bool(v) for row, v in zip(P, ok)}
        return ok

    def _remembered(self, pt: np.ndarray):
        cache = self.__dict__.get("_prescreened")
        return None if cache is None else cache.get(pt.tobytes())


class ImgCollisionChecker(_GridChecker):
    """2-D image-space checker (collisionChecker.py:83-215) on the GPU."""

    def __init__(self, img, device: int = 0):
        """
        :param img: a file-like object (e.g. a filename) for the image as the
            environment that the planning operates in
        :param device: CUDA device ordinal
        """
        super().__init__()
        self.image_fname = img
        image = _load_binarised(img)
        # transposed like the reference (:108): _img[x, y]
        self._attach(image.T, device)

    @classmethod
    def from_free_mask(cls, free_xy, device: int = 0, image_fname=None):
        """Build from an occupancy grid ``free_xy[x, y]`` (truthy = free) without going
        through PIL: for procedurally generated 8k-32k maps.  ``free_xy`` may be a
        numpy array or a CUDA uint8 tensor (then ``_img`` stays on the device)."""
        self = cls.__new__(cls)
        CollisionChecker.__init__(self)
        self.image_fname = image_fname
        if _is_torch_tensor(free_xy):
            self._attach(free_xy, device)
        else:
            self._attach(np.ascontiguousarray(np.asarray(free_xy) != 0).astype(np.uint8), device)
        return self

    # ---- reference interface -------------------------------------------------
    def feasible(self, p, save_stats=True):
        """collisionChecker.py:147-153.  ``IndexError`` -> False, wrap-around of negative
        indices like numpy, ``ValueError`` for NaN, ``OverflowError`` for inf."""
        if save_stats:
            self.stats.feasible_cnt += 1
        pt = np.array([float(p[0]), float(p[1])], dtype=np.float64).reshape(1, 2)
        hit = self._remembered(pt)
        if hit is not None:
            return np.bool_(hit)
        return np.bool_(self._feasible_host(pt)[0])

    def visible(self, pos1, pos2):
        """collisionChecker.py:134-145: all Bresenham pixels between the int()-truncated
        endpoints free; a NaN endpoint gives False (the ValueError is swallowed)."""
        self.stats.visible_cnt += 1
        P = np.array([float(pos1[0]), float(pos1[1])], dtype=np.float64).reshape(1, 2)
        Q = np.array([float(pos2[0]), float(pos2[1])], dtype=np.float64).reshape(1, 2)
        return bool(self._visible_host(P, Q)[0])

    def get_coor_before_collision(self, pos1, pos2):
        """collisionChecker.py:118-132: the first blocked pixel walking pos1 -> pos2, or
        the last pixel of the line when it is free."""
        P = np.array([float(pos1[0]), float(pos1[1])], dtype=np.float64).reshape(1, 2)
        Q = np.array([float(pos2[0]), float(pos2[1])], dtype=np.float64).reshape(1, 2)
        # the reference calls feasible() (with stats) once per examined pixel (:130)
        xy, n_examined = self._first_blocked_host(P, Q, with_counts=True)
        self.stats.feasible_cnt += int(n_examined[0])
        return (int(xy[0, 0]), int(xy[0, 1]))

    @staticmethod
    def get_line(start, end):
        """Bresenham pixel list start -> end (collisionChecker.py:155-215).

        A host-side enumeration utility (the likelihood sampler iterates over it,
        samplers/likelihoodPolicySampler.py:122); collision answers never use it."""
        return [tuple(p) for p in bresenham_pixels(start, end).tolist()]

    # ---- batched entry points --------------------------------------------------
    def feasible_batch(self, P, save_stats=True):
        """``[feasible(p) for p in P]`` in one launch.  numpy in -> numpy bool out;
        CUDA tensor in -> CUDA bool tensor out (no host round trip)."""
        if self._mocked("feasible"):
            return np.array([bool(self.feasible(p)) for p in np.asarray(P)], dtype=bool)
        n = len(P)
        if save_stats:
            self.stats.feasible_cnt += n
        if _is_torch_tensor(P):
            return self._feasible_device(P)
        return self._feasible_host(as_host_f64(P, 2))

    def visible_batch(self, P, Q):
        """``[visible(p, q) for p, q in zip(P, Q)]`` in one launch."""
        if self._mocked("visible"):
            return np.array([bool(self.visible(p, q)) for p, q in zip(np.asarray(P), np.asarray(Q))],
                            dtype=bool)
        n = len(P)
        self.stats.visible_cnt += n
        if _is_torch_tensor(P):
            return self._visible_device(P, Q)
        return self._visible_host(as_host_f64(P, 2), as_host_f64(Q, 2))

    def first_blocked_batch(self, P, Q):
        """Batched :meth:`get_coor_before_collision`: int32 [n][2]."""
        return self._first_blocked_host(as_host_f64(P, 2), as_host_f64(Q, 2))

    # ---- launches ----------------------------------------------------------------
    def _feasible_host(self, P):
        n = len(P)
        out = np.empty(n, np.uint8)
        flags = C.c_int32(0)
        check(_lib.load().sbp_feasible2d_host(self.device_map.handle, ptr(P), n, ptr(out),
                                              C.byref(flags)), "sbp_feasible2d_host")
        raise_for_flags(flags.value, "feasible")
        return out.astype(bool)

    def _visible_host(self, P, Q):
        n = len(P)
        if len(Q) != n:
            raise ValueError("P and Q must have the same length")
        out = np.empty(n, np.uint8)
        flags = C.c_int32(0)
        check(_lib.load().sbp_visible2d_host(self.device_map.handle, ptr(P), ptr(Q), n, ptr(out),
                                             C.byref(flags)), "sbp_visible2d_host")
        raise_for_flags(flags.value, "visible", nan_is_error=False)   # ValueError is caught (:143)
        return out.astype(bool)

    def _first_blocked_host(self, P, Q, with_counts=False):
        n = len(P)
        out = np.empty((n, 2), np.int32)
        flags = C.c_int32(0)
        check(_lib.load().sbp_first_blocked2d_host(self.device_map.handle, ptr(P), ptr(Q), n,
                                                   ptr(out), C.byref(flags)),
              "sbp_first_blocked2d_host")
        raise_for_flags(flags.value, "get_coor_before_collision")
        if (out == np.iinfo(np.int32).min).any():
            raise ValueError("get_coor_before_collision: |coordinate| >= 2**30 is not supported")
        if not with_counts:
            return out
        # pixels examined = index of the reported pixel along the line + 1
        x1 = np.trunc(P[:, 0]).astype(np.int64)
        y1 = np.trunc(P[:, 1]).astype(np.int64)
        examined = np.maximum(np.abs(out[:, 0] - x1), np.abs(out[:, 1] - y1)) + 1
        return out, examined

    def _feasible_device(self, P):
        import torch

        m = self.device_map
        m.ctx.use_torch_stream()
        P = as_device_f64(P, 2, device=self._device)
        n = P.shape[0]
        out = torch.empty(n, dtype=torch.uint8, device=P.device)
        flags = torch.zeros(1, dtype=torch.int32, device=P.device)
        check(_lib.load().sbp_feasible2d(m.handle, ptr(P), n, ptr(out), ptr(flags)), "sbp_feasible2d")
        self._last_flags = flags     # inspected lazily: no synchronisation on this path
        return out.view(torch.bool)          # the kernels write 0 / 1: reinterpret, no copy

    def _visible_device(self, P, Q, px_tested=None):
        import torch

        m = self.device_map
        m.ctx.use_torch_stream()
        P, Q = as_device_f64(P, 2, device=self._device), as_device_f64(Q, 2, device=self._device)
        n = P.shape[0]
        if Q.shape[0] != n:
            raise ValueError("P and Q must have the same length")
        out = torch.empty(n, dtype=torch.uint8, device=P.device)
        flags = torch.zeros(1, dtype=torch.int32, device=P.device)
        check(_lib.load().sbp_visible2d(m.handle, ptr(P), ptr(Q), n, ptr(out), ptr(flags),
                                        ptr(px_tested)), "sbp_visible2d")
        self._last_flags = flags
        return out.view(torch.bool)          # the kernels write 0 / 1: reinterpret, no copy

    def knn_edges_visible(self, tree, nbr_in, first_row: int = 0, queries=None, packed_out=None):
        """``tree.knn_edges`` + :meth:`visible_batch` in ONE launch (prmPlanner.py:91-101):
        the candidate edges of the kNN answer ``nbr_in`` are resolved against the node store
        inside the visibility kernel.  Returns ``(nbr, vis)`` as CUDA tensors [m][k_out];
        ``vis`` already excludes empty neighbour slots.  Counts m * k_out visible calls."""
        import torch

        m = self.device_map
        m.ctx.use_torch_stream()
        nbr_in = nbr_in.contiguous()
        rows, k_in = nbr_in.shape
        k_out = k_in if queries is not None else k_in - 1
        dev = nbr_in.device
        nbr = torch.empty((rows, k_out), dtype=torch.int64, device=dev)
        vis = torch.empty((rows, k_out), dtype=torch.uint8, device=dev)
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        X = as_device_f64(queries, 2, device=self._device) if queries is not None else None
        self.stats.visible_cnt += rows * k_out
        # packed adjacency written by the kernel itself: a local int32 tensor, or (address, True)
        # = the multicast address of a symmetric buffer (PeerAdjacencyGather.target)
        p_ptr, p_mc, p_sync = None, 0, None
        if packed_out is not None:
            if isinstance(packed_out, tuple):
                p_ptr, p_mc = C.c_void_p(int(packed_out[0])), int(bool(packed_out[1]))
                if len(packed_out) > 2 and packed_out[2] is not None:   # in-kernel rendezvous block
                    p_sync = (C.c_int64 * 6)(*[int(v) for v in packed_out[2]])
            else:
                if packed_out.dtype != torch.int32 or packed_out.numel() != rows * k_out \
                        or not packed_out.is_contiguous() or packed_out.device != dev:
                    raise ValueError("packed_out must be a contiguous CUDA int32 tensor [m][k]")
                p_ptr = ptr(packed_out)
        check(_lib.load().sbp_knn_edges_visible2d(m.handle, tree._h, ptr(nbr_in), rows, k_in,
                                                  int(first_row), ptr(X), ptr(nbr), ptr(vis),
                                                  p_ptr, p_mc, p_sync, ptr(flags)), "sbp_knn_edges_visible2d")
        self._last_flags = flags
        return nbr, vis.view(torch.bool)

    def prm_connect(self, tree, k: int, part=(0, 1), packed_out=None, want_rows: bool = True, order_out=None):
        """The whole roadmap connection step for the rows of ``tree`` in ONE library call
        (``sbp_prm_connect2d``; prmPlanner.py:91-101 with the k nearest neighbours as candidates):
        exact kNN of every stored node among all nodes, own entry dropped, and
        ``visible(m_g.pos, v.pos)`` of the k candidates, rows walked in the cell-sorted order of
        the kNN's grid.

        ``part=(i, parts)``: only run i of the ``parts`` contiguous runs of that order (the shard
        of rank i; the other rows of the outputs are left untouched).  ``packed_out``: int32 CUDA
        tensor [n][k], or ``(address, multicast)`` -- pinned host memory, or the NVSwitch
        multicast address of a symmetric buffer (``distributed.PeerAdjacencyBuffer.target``).
        ``order_out``: int32 CUDA tensor [n]; when given the rows of every output are in
        CELL-SORTED order -- row p belongs to node ``order_out[p]`` (written by the call) -- and the
        rows of a part are one contiguous, coalesced block (``sbp_prm_connect2d_sorted``): the form
        for pinned host memory and the multicast mapping.  With several parts only the part's own
        slice ``order_out[p0 : p0 + count]`` is written (a rank builds the grid for its band of cell
        rows only); with a multicast ``packed_out`` pass the multicast ADDRESS of the order buffer
        (``PeerAdjacencyBuffer.order_target``).
        Returns ``(nbr int64 [n][k], vis bool [n][k])`` CUDA tensors, or ``None`` with
        ``want_rows=False`` (only the packed adjacency is written)."""
        import torch

        m = self.device_map
        m.ctx.use_torch_stream()
        n = len(tree)
        dev = torch.device("cuda", m.ctx.device)
        nbr = torch.empty((n, k), dtype=torch.int64, device=dev) if want_rows else None
        vis = torch.empty((n, k), dtype=torch.uint8, device=dev) if want_rows else None
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        p_ptr, p_mc = None, 0
        if packed_out is not None:
            if isinstance(packed_out, tuple):
                p_ptr, p_mc = C.c_void_p(int(packed_out[0])), int(bool(packed_out[1]))
            else:
                if packed_out.dtype != torch.int32 or packed_out.numel() != n * k \
                        or not packed_out.is_contiguous() or packed_out.device != dev:
                    raise ValueError("packed_out must be a contiguous CUDA int32 tensor [n][k]")
                p_ptr = ptr(packed_out)
        elif not want_rows:
            raise ValueError("nothing to compute: want_rows=False and no packed_out")
        i, parts = int(part[0]), int(part[1])
        base, rem = divmod(n, parts)
        self.stats.visible_cnt += (base + (1 if i < rem else 0)) * k
        if order_out is not None:
            if isinstance(order_out, int):               # raw address (the multicast mapping, with packed_out)
                o_ptr = C.c_void_p(order_out)
            else:
                if order_out.dtype != torch.int32 or order_out.numel() != n or not order_out.is_contiguous() \
                        or order_out.device != dev:
                    raise ValueError("order_out must be a contiguous CUDA int32 tensor [n]")
                if p_mc:
                    raise ValueError("with a multicast packed_out, order_out must be a multicast address too")
                o_ptr = ptr(order_out)
            check(_lib.load().sbp_prm_connect2d_sorted(m.handle, tree._h, int(k), i, parts, ptr(nbr), ptr(vis),
                                                       p_ptr, p_mc, o_ptr, ptr(flags)),
                  "sbp_prm_connect2d_sorted")
        else:
            check(_lib.load().sbp_prm_connect2d(m.handle, tree._h, int(k), i, parts, ptr(nbr), ptr(vis),
                                                p_ptr, p_mc, ptr(flags)), "sbp_prm_connect2d")
        self._last_flags = flags
        return (nbr, vis.view(torch.bool)) if want_rows else None

    def check_device_flags(self, what="batch"):
        """Synchronise and raise what CPython would have raised for the last device-side
        batch (NaN in ``feasible`` -> ValueError, inf -> OverflowError)."""
        f = self.__dict__.get("_last_flags")
        if f is not None:
            raise_for_flags(int(f.item()), what)


def bresenham_pixels(start, end) -> np.ndarray:
    """All pixels of the reference's Bresenham line, start -> end order, int64 [n][2].

    Closed form of collisionChecker.py:172-215 (derivation in csrc/collide2d.cu):
    after the steep / order swaps pixel i is ``(a1 + i, b1 + ystep * ((i*|db| + c) // da))``
    with ``c = da - 1 - da // 2``."""
    x1, y1 = (int(v) for v in start)
    x2, y2 = (int(v) for v in end)
    dx, dy = x2 - x1, y2 - y1
    steep = abs(dy) > abs(dx)
    a1, b1, a2, b2 = (y1, x1, y2, x2) if steep else (x1, y1, x2, y2)
    swapped = a1 > a2
    if swapped:
        a1, a2, b1, b2 = a2, a1, b2, b1
    da, adb = a2 - a1, abs(b2 - b1)
    ystep = 1 if b1 < b2 else -1
    i = np.arange(da + 1, dtype=np.int64)
    k = (i * adb + (da - 1 - da // 2)) // da if da else np.zeros(1, np.int64)
    a, b = a1 + i, b1 + ystep * k
    pts = np.stack([b, a] if steep else [a, b], axis=1)
    return pts[::-1].copy() if swapped else pts


class RobotArm4dCollisionChecker(_GridChecker):
    """4-D (x, y, r0, r1) two-link stick arm in image space (collisionChecker.py:299-503)
    on the GPU."""

    _dim = 4

    def __init__(self, img, map_mat: typing.Optional[np.ndarray] = None,
                 stick_robot_length_config: typing.Optional[typing.Sequence[float]] = None,
                 device: int = 0):
        """
        :param img: a file-like object (e.g. a filename) for the image of the environment
        :param map_mat: if given, ignores ``img`` and uses this matrix as the map
        :param stick_robot_length_config: the two link lengths of the stick arm
        """
        super().__init__()
        self.image_fname = img
        if map_mat is None:
            image = _load_binarised(img)
        else:
            image = (map_mat / map_mat.max()).astype(int)          # :333-334
        if stick_robot_length_config is not None:
            self.stick_robot_length_config = stick_robot_length_config
        self._attach(image.T, device)                              # :340

    def _lengths(self):
        L = self.stick_robot_length_config
        return float(L[0]), float(L[1])

    # ---- reference interface -------------------------------------------------
    @staticmethod
    def create_ranges(start, stop, N: int, endpoint: bool = True):
        """Batch of linspaces, ``steps * arange(N) + start`` (collisionChecker.py:346-364).
        Host helper kept for API parity; the kernels evaluate the same three fp64
        operations per angle on the device."""
        start, stop = np.asarray(start), np.asarray(stop)
        divisor = N - 1 if endpoint == 1 else N
        steps = (1.0 / divisor) * (stop - start)
        return steps[:, None] * np.arange(N) + start[:, None]

    def _interpolate_configs(self, c1, c2):
        """[N][4] interpolated configs (collisionChecker.py:366-383); host helper."""
        loc = self._get_line(c1[:2], c2[:2])
        if len(loc) == 1:
            loc.append(loc[0])
        rot = self.create_ranges(np.asarray(c1[2:]), np.asarray(c2[2:]), N=len(loc))
        return np.concatenate([np.array(loc).T, rot]).T

    def visible(self, pos1, pos2):
        """collisionChecker.py:385-391."""
        self.stats.visible_cnt += 1
        a = np.array([float(v) for v in pos1[:4]], dtype=np.float64).reshape(1, 4)
        b = np.array([float(v) for v in pos2[:4]], dtype=np.float64).reshape(1, 4)
        return bool(self._visible_host(a, b)[0])

    def _pt_feasible(self, p):
        """collisionChecker.py:393-402: one pixel test (no stats)."""
        pt = np.array([float(p[0]), float(p[1])], dtype=np.float64).reshape(1, 2)
        out = np.empty(1, np.uint8)
        flags = C.c_int32(0)
        check(_lib.load().sbp_feasible2d_host(self.device_map.handle, ptr(pt), 1, ptr(out),
                                              C.byref(flags)), "sbp_feasible2d_host")
        raise_for_flags(flags.value, "_pt_feasible")
        return np.bool_(out[0])

    def feasible(self, p, save_stats=True):
        """collisionChecker.py:404-426."""
        if save_stats:
            self.stats.feasible_cnt += 1
        c = np.array([float(v) for v in p[:4]], dtype=np.float64).reshape(1, 4)
        hit = self._remembered(c)
        if hit is not None:
            return hit
        return bool(self._feasible_host(c)[0])

    @staticmethod
    def get_pt_from_angle_and_length(pt1, angle, line_length):
        """collisionChecker.py:428-441 (used by the visualiser to draw the arm)."""
        return (pt1[0] + line_length * math.cos(angle), pt1[1] + line_length * math.sin(angle))

    @staticmethod
    def _get_line(start, end):
        """List-of-lists variant of the Bresenham line (collisionChecker.py:443-503)."""
        return bresenham_pixels(start, end).tolist()

    # ---- batched entry points --------------------------------------------------
    def feasible_batch(self, Cfg, save_stats=True):
        if self._mocked("feasible"):
            return np.array([bool(self.feasible(c)) for c in np.asarray(Cfg)], dtype=bool)
        if save_stats:
            self.stats.feasible_cnt += len(Cfg)
        if _is_torch_tensor(Cfg):
            return self._feasible_device(Cfg)
        return self._feasible_host(as_host_f64(Cfg, 4))

    def visible_batch(self, C1, C2):
        if self._mocked("visible"):
            return np.array([bool(self.visible(a, b)) for a, b in zip(np.asarray(C1), np.asarray(C2))],
                            dtype=bool)
        self.stats.visible_cnt += len(C1)
        if _is_torch_tensor(C1):
            return self._visible_device(C1, C2)
        return self._visible_host(as_host_f64(C1, 4), as_host_f64(C2, 4))

    # ---- launches ----------------------------------------------------------------
    #: number of guard-band entries settled by the host libm in the last call
    last_ambiguous = 0

    def _feasible_host(self, Cfg):
        n = len(Cfg)
        L0, L1 = self._lengths()
        out = np.empty(n, np.uint8)
        flags, amb = C.c_int32(0), C.c_int64(0)
        check(_lib.load().sbp_arm_feasible_host(self.device_map.handle, L0, L1, ptr(Cfg), n, ptr(out),
                                                C.byref(flags), C.byref(amb)), "sbp_arm_feasible_host")
        raise_for_flags(flags.value, "feasible")
        self.last_ambiguous = int(amb.value)
        return out.astype(bool)

    def _visible_host(self, C1, C2):
        n = len(C1)
        if len(C2) != n:
            raise ValueError("C1 and C2 must have the same length")
        L0, L1 = self._lengths()
        out = np.empty(n, np.uint8)
        flags, amb = C.c_int32(0), C.c_int64(0)
        check(_lib.load().sbp_arm_visible_host(self.device_map.handle, L0, L1, ptr(C1), ptr(C2), n,
                                               ptr(out), C.byref(flags), C.byref(amb)),
              "sbp_arm_visible_host")
        raise_for_flags(flags.value, "visible")          # the arm checker has no guard (:385-391)
        self.last_ambiguous = int(amb.value)
        return out.astype(bool)

    def _feasible_device(self, Cfg, resolve=True):
        """Device tensor in, CUDA bool tensor out (guard-band entries settled like
        :meth:`_visible_device`)."""
        import torch

        m = self.device_map
        m.ctx.use_torch_stream()
        Cfg = as_device_f64(Cfg, 4, device=self._device)
        n = Cfg.shape[0]
        L0, L1 = self._lengths()
        out = torch.empty(n, dtype=torch.uint8, device=Cfg.device)
        flags = torch.zeros(1, dtype=torch.int32, device=Cfg.device)
        check(_lib.load().sbp_arm_feasible(m.handle, L0, L1, ptr(Cfg), n, ptr(out), ptr(flags)),
              "sbp_arm_feasible")
        self._last_flags = flags
        if not resolve:
            return out
        amb = torch.nonzero(out == _lib.AMBIGUOUS).flatten()
        self.last_ambiguous = int(amb.numel())
        if amb.numel():
            a = Cfg[amb].cpu().numpy()
            o = np.full(len(a), _lib.AMBIGUOUS, np.uint8)
            check(_lib.load().sbp_arm_resolve_host(m.handle, L0, L1, ptr(a), None, len(a), ptr(o), None),
                  "sbp_arm_resolve_host")
            out[amb] = torch.from_numpy(o).to(out.device)
        return out.view(torch.bool)          # the kernels write 0 / 1: reinterpret, no copy

    def _visible_device(self, C1, C2, cfg_tested=None, resolve=True):
        """Device tensors in, CUDA bool tensor out.  Guard-band entries (rare) are
        settled through one small D2H / host libm / H2D round trip unless
        ``resolve=False`` (then the raw 0/1/2 bytes are returned)."""
        import torch

        m = self.device_map
        m.ctx.use_torch_stream()
        C1, C2 = as_device_f64(C1, 4, device=self._device), as_device_f64(C2, 4, device=self._device)
        n = C1.shape[0]
        L0, L1 = self._lengths()
        out = torch.empty(n, dtype=torch.uint8, device=C1.device)
        flags = torch.zeros(1, dtype=torch.int32, device=C1.device)
        check(_lib.load().sbp_arm_visible(m.handle, L0, L1, ptr(C1), ptr(C2), n, ptr(out), ptr(flags),
                                          ptr(cfg_tested)), "sbp_arm_visible")
        self._last_flags = flags
        if not resolve:
            return out
        amb = torch.nonzero(out == _lib.AMBIGUOUS).flatten()
        self.last_ambiguous = int(amb.numel())
        if amb.numel():
            a = C1[amb].cpu().numpy()
            b = C2[amb].cpu().numpy()
            o = np.full(len(a), _lib.AMBIGUOUS, np.uint8)
            check(_lib.load().sbp_arm_resolve_host(m.handle, L0, L1, ptr(a), ptr(b), len(a), ptr(o),
                                                   None), "sbp_arm_resolve_host")
            out[amb] = torch.from_numpy(o).to(out.device)
        return out.view(torch.bool)          # the kernels write 0 / 1: reinterpret, no copy
