"""Device-resident, append-only node array with brute-force nearest / kNN / radius
queries: the GPU twin of the planners' ``poses`` arrays.

Reference call sites replaced (file:line under /root/reference/sbp_env/planners/):

  poses = np.empty((2*max_nodes+50, d)); poses[len(nodes)] = pos    rrtPlanner.py:24-26, :106-113
  find_nearest_neighbour_idx(pos, poses)                            rrtPlanner.py:268-279
  np.argsort(np.linalg.norm(poses - pos, axis=1))[:20]              rrtPlanner.py:151-152, :226-227
  nearest_neighbours(nodes, poses, pos, radius)                     prmPlanner.py:28-44
  engine.dist(new, p) <= radius gate                                rrtPlanner.py:177-181
  cross-tree nearest of Bi-RRT                                      birrtPlanner.py:82-85
  per-tree arrays and bulk extend of RRdT*                          rrdtPlanner.py:858-887

Contract: fp64, d = sqrt_rn of the left-to-right unfused sum
of squares over ALL dims (``metric="full"``) or the first two dims only
(``metric="xy"``, the ``engine.dist`` form); results ordered by (d, index); radius
rows are index-ascending; ``closed`` selects ``<=`` instead of ``<``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .runtime import _is_torch_tensor, as_device_f64, as_host_f64, get_context, ptr


class NodeStore:
    def __init__(self, dim: int, capacity: int = 1024, device: int = 0):
        self.dim = int(dim)
        self.ctx = get_context(device)
        h = C.c_void_p()
        check(_lib.load().sbp_nodes_create(self.ctx.handle, self.dim, int(capacity), C.byref(h)),
              "sbp_nodes_create")
        self._h = h
        # host mirror of the poses (the planners keep `poses` on the host anyway); it lets
        # the per-node host arithmetic (costs, step_from_to) avoid device reads
        self._mirror = np.empty((max(int(capacity), 16), self.dim), np.float64)
        self._mirror_n = 0          # rows valid in the mirror; < len(self) after device extends

    # ---- growth ------------------------------------------------------------------
    def append(self, x) -> int:
        """``poses[len(nodes)] = pos`` (rrtPlanner.py:112); returns the new node's index."""
        n = len(self)
        self.extend(np.asarray(x, dtype=np.float64).reshape(1, self.dim))
        return n

    def extend(self, X) -> None:
        """Bulk append of rows (RRdT* ``extend_tree``, rrdtPlanner.py:878-887)."""
        if _is_torch_tensor(X) and not X.is_cuda:
            import torch

            zero_copy = X.is_pinned() and X.dtype == torch.float64 and X.is_contiguous() and \
                (X.dim() == 2 and X.shape[1] == self.dim or X.dim() == 1 and X.numel() == self.dim)
            if not zero_copy:
                # pageable, or a pinned tensor that would need a dtype / layout conversion (which
                # yields a pageable temporary): the staged host path below
                X = X.detach().to(torch.float64).contiguous().numpy()
        if _is_torch_tensor(X):
            # CUDA tensor, or PINNED host tensor: pinned memory is mapped into the device's address
            # space (UVA), so the append kernel reads it over PCIe itself -- the host-to-device
            # transfer and the AoS -> SoA conversion are one asynchronous, graph-capturable launch
            # (the caller must keep the tensor unchanged until the stream has passed the launch)
            self.ctx.use_torch_stream()
            X = as_device_f64(X, self.dim, device=self.ctx.device, allow_pinned=True)
            check(_lib.load().sbp_nodes_append(self._h, ptr(X), X.shape[0], 1), "sbp_nodes_append")
            # the launch is asynchronous: keep the source alive (a pinned block handed back to
            # torch's allocator could be recycled under the kernel) until the next append
            self._keepalive = X
        else:
            X = as_host_f64(X, self.dim)
            n0 = len(self)
            check(_lib.load().sbp_nodes_append(self._h, ptr(X), len(X), 0), "sbp_nodes_append")
            if self._mirror_n == n0:                      # mirror was complete: keep it so
                need = n0 + len(X)
                if need > len(self._mirror):
                    grown = np.empty((max(need, 2 * len(self._mirror)), self.dim), np.float64)
                    grown[:n0] = self._mirror[:n0]
                    self._mirror = grown
                self._mirror[n0:need] = X
                self._mirror_n = need

    def truncate(self, n: int) -> None:
        check(_lib.load().sbp_nodes_truncate(self._h, int(n)), "sbp_nodes_truncate")
        self._mirror_n = min(self._mirror_n, int(n))

    def poses_rows(self, idx) -> np.ndarray:
        """Host copy of the given rows ([len(idx)][d]); served from the host mirror when
        the store was filled from host arrays, else read back from the device."""
        idx = np.asarray(idx, dtype=np.int64).reshape(-1)
        if len(idx) == 0:
            return np.empty((0, self.dim), np.float64)
        if self._mirror_n == len(self):
            return self._mirror[idx]
        out = np.empty((len(idx), self.dim), np.float64)
        row = np.empty(self.dim, np.float64)
        for j, i in enumerate(idx):
            check(_lib.load().sbp_nodes_read(self._h, int(i), 1, ptr(row)), "sbp_nodes_read")
            out[j] = row
        return out

    def __len__(self) -> int:
        n = C.c_int64()
        check(_lib.load().sbp_nodes_size(self._h, C.byref(n)), "sbp_nodes_size")
        return int(n.value)

    @property
    def poses(self) -> np.ndarray:
        """Host copy of ``poses[:n]`` (parity dumps)."""
        n = len(self)
        out = np.empty((n, self.dim), np.float64)
        if n:
            check(_lib.load().sbp_nodes_read(self._h, 0, n, ptr(out)), "sbp_nodes_read")
        return out

    def rows_device(self, first: int = 0, count: int = None):
        """CUDA tensor [count][d] holding stored rows first..first+count (no host trip)."""
        import torch

        self.ctx.use_torch_stream()
        count = len(self) - first if count is None else int(count)
        out = torch.empty((count, self.dim), dtype=torch.float64,
                          device=torch.device("cuda", self.ctx.device))
        check(_lib.load().sbp_nodes_rows_device(self._h, int(first), count, ptr(out)),
              "sbp_nodes_rows_device")
        return out

    # ---- queries -------------------------------------------------------------------
    def nearest(self, x) -> int:
        """``np.argmin(np.linalg.norm(poses - pos, axis=1))`` (rrtPlanner.py:278-279): the first
        minimum, or -- like np.argmin -- the first node at a NaN distance when there is one."""
        if _is_torch_tensor(x):
            return int(self.nearest_batch(x.reshape(1, self.dim))[0][0])
        # one host query: one launch, the answer comes back through pinned memory
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(self.dim)
        self.ctx.use_torch_stream()
        idx = C.c_int64()
        check(_lib.load().sbp_nodes_nearest1_host(self._h, x.ctypes.data, C.byref(idx), None),
              "sbp_nodes_nearest1_host")
        return int(idx.value)

    def nearest_batch(self, X, return_dist: bool = True):
        """:meth:`nearest` for every row of X: ``(idx int64 [m], dist fp64 [m])``; numpy in ->
        numpy out, CUDA tensor in -> CUDA tensors out."""
        if _is_torch_tensor(X):
            import torch

            self.ctx.use_torch_stream()
            X = as_device_f64(X, self.dim, device=self.ctx.device)
            m = X.shape[0]
            idx = torch.empty(m, dtype=torch.int64, device=X.device)
            dist = torch.empty(m, dtype=torch.float64, device=X.device) if return_dist else None
            check(_lib.load().sbp_nodes_nearest(self._h, ptr(X), m, ptr(idx), ptr(dist)), "sbp_nodes_nearest")
            return idx, dist
        X = as_host_f64(X, self.dim)
        m = len(X)
        idx = np.empty(m, np.int64)
        dist = np.empty(m, np.float64) if return_dist else None
        check(_lib.load().sbp_nodes_nearest_host(self._h, ptr(X), m, ptr(idx), ptr(dist)),
              "sbp_nodes_nearest_host")
        return idx, dist

    def knn(self, X, k: int, return_dist: bool = True, precision: int = 64):
        """k nearest nodes of every row of X, ordered by (distance, index).

        numpy in -> numpy out; CUDA tensor in -> CUDA tensors out.
        Returns ``(idx int64 [m][k], dist [m][k])``; rows are padded with ``-1`` / ``inf`` when
        fewer than k nodes are stored.  ``precision=64``: fp64 distances, bit-exact with
        ``np.linalg.norm``; ``precision=32``: the same rows (the selection stays exact) with
        float32 distances, ``|d32 - d64| <= 2**-24 * d64``."""
        k = int(k)
        if precision not in (32, 64):
            raise ValueError("precision is 64 or 32")
        if _is_torch_tensor(X):
            import torch

            self.ctx.use_torch_stream()
            X = as_device_f64(X, self.dim, device=self.ctx.device)
            m = X.shape[0]
            idx = torch.empty((m, k), dtype=torch.int64, device=X.device)
            dt = torch.float64 if precision == 64 else torch.float32
            dist = torch.empty((m, k), dtype=dt, device=X.device) if return_dist else None
            check(_lib.load().sbp_nodes_knn(self._h, ptr(X), m, k, precision, ptr(idx), ptr(dist)),
                  "sbp_nodes_knn")
            return idx, dist
        X = as_host_f64(X, self.dim)
        m = len(X)
        idx = np.empty((m, k), np.int64)
        dist = np.empty((m, k), np.float64 if precision == 64 else np.float32) if return_dist else None
        check(_lib.load().sbp_nodes_knn_host(self._h, ptr(X), m, k, precision, ptr(idx), ptr(dist)),
              "sbp_nodes_knn_host")
        return idx, dist

    def near(self, X, r: float, metric: str = "full", closed: bool = False):
        """All nodes within radius r of every row of X, as CSR ``(row_ptr, idx)`` with
        node indices ascending in each row (the order of the reference's loops,
        prmPlanner.py:40-44, rrtPlanner.py:176-187).

        ``metric="full", closed=False`` is ``prmPlanner.nearest_neighbours``;
        ``metric="xy", closed=True`` is the ``engine.dist(...) <= radius`` gate."""
        met = _lib.METRIC[metric]
        lib = _lib.load()
        if _is_torch_tensor(X):
            import torch

            self.ctx.use_torch_stream()
            X = as_device_f64(X, self.dim, device=self.ctx.device)
            m = X.shape[0]
            row_ptr = torch.empty(m + 1, dtype=torch.int64, device=X.device)
            needed = torch.zeros(1, dtype=torch.int64, device=X.device)
            # one call (count -> scan -> ordered fill) into a generously sized buffer; the
            # size of the answer has to reach the host anyway, and only an overflow costs a
            # second call
            cap = max(1 << 20, 64 * m)
            idx = torch.empty(cap, dtype=torch.int64, device=X.device)
            check(lib.sbp_nodes_near(self._h, ptr(X), m, float(r), met, int(bool(closed)),
                                     ptr(row_ptr), ptr(idx), cap, ptr(needed)), "sbp_nodes_near")
            total = int(needed.item())
            if total > cap:
                idx = torch.empty(total, dtype=torch.int64, device=X.device)
                check(lib.sbp_nodes_near(self._h, ptr(X), m, float(r), met, int(bool(closed)),
                                         ptr(row_ptr), ptr(idx), total, ptr(needed)), "sbp_nodes_near")
            return row_ptr, idx[:total]
        X = as_host_f64(X, self.dim)
        m = len(X)
        row_ptr = np.empty(m + 1, np.int64)
        needed = C.c_int64(0)
        check(lib.sbp_nodes_near_host(self._h, ptr(X), m, float(r), met, int(bool(closed)),
                                      ptr(row_ptr), None, 0, C.byref(needed)), "sbp_nodes_near_host")
        total = int(needed.value)
        idx = np.empty(max(total, 1), np.int64)
        if total:
            check(lib.sbp_nodes_near_host(self._h, ptr(X), m, float(r), met, int(bool(closed)),
                                          ptr(row_ptr), ptr(idx), total, C.byref(needed)),
                  "sbp_nodes_near_host")
        return row_ptr, idx[:total]

    # ---- batched planner primitive ---------------------------------------------------
    def gather_edges(self, nbr, first_row: int = 0):
        """Device-side endpoints of the candidate edges ``row -> nbr[row][j]``:
        ``(P, Q, valid)`` with P = neighbour pose, Q = row pose, the argument order of
        ``cc.visible(m_g.pos, v.pos)`` in PRMPlanner.build_graph (prmPlanner.py:97)."""
        import torch

        self.ctx.use_torch_stream()
        nbr = nbr.contiguous()
        m, k = nbr.shape
        P = torch.empty((m * k, self.dim), dtype=torch.float64, device=nbr.device)
        Q = torch.empty((m * k, self.dim), dtype=torch.float64, device=nbr.device)
        valid = torch.empty(m * k, dtype=torch.uint8, device=nbr.device)
        check(_lib.load().sbp_nodes_gather_edges(self._h, ptr(nbr), m, k, int(first_row), ptr(P),
                                                 ptr(Q), ptr(valid)), "sbp_nodes_gather_edges")
        return P, Q, valid

    def near_visible(self, cc, x, r: float, metric: str = "xy", closed: bool = True,
                     direction: int = 0, cap: int = 256):
        """Nodes within radius r of the single configuration x together with the
        visibility of the connecting edge, in ONE kernel launch: ``(idx, vis)`` sorted by
        node index (the order of the reference's loops).

        ``direction=0`` checks ``cc.visible(x, node)`` (rrtPlanner.py:179-181,
        prmPlanner.py:116); ``direction=1`` checks ``cc.visible(node, x)`` (rrtPlanner.py:239,
        :260).  ``cc`` is the GPU checker whose map lives on the same device; the arm
        checker is used with its link lengths.  ``cc.stats.visible_cnt`` is advanced by the
        number of edges checked, as the reference's loop would."""
        arm = hasattr(cc, "stick_robot_length_config")
        L0, L1 = (float(v) for v in cc.stick_robot_length_config[:2]) if arm else (0.0, 0.0)
        xh = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(self.dim))
        lib = _lib.load()
        while True:
            idx = np.empty(cap, np.int64)
            vis = np.empty(cap, np.uint8)
            hits, flags = C.c_int32(0), C.c_int32(0)
            rc = lib.sbp_nodes_near_visible_host(self._h, cc.device_map.handle, int(arm), L0, L1,
                                                 ptr(xh), float(r), _lib.METRIC[metric],
                                                 int(bool(closed)), int(direction), ptr(idx), ptr(vis),
                                                 cap, C.byref(hits), C.byref(flags))
            if rc == -4 and hits.value > cap:            # SBP_ERR_CAPACITY: grow and retry
                cap = int(hits.value) * 2
                continue
            check(rc, "sbp_nodes_near_visible_host")
            break
        from .runtime import raise_for_flags

        # the image checker swallows NaN (ValueError caught, collisionChecker.py:143-144)
        raise_for_flags(flags.value, "visible", nan_is_error=arm)
        n = int(hits.value)
        cc.stats.visible_cnt += n
        return idx[:n].copy(), vis[:n].astype(bool)

    def knn_edges(self, nbr_in, first_row: int = 0, queries=None):
        """kNN rows -> candidate-edge batch in one launch: ``(nbr, P, Q, valid)``.

        Without ``queries`` row i is stored node ``first_row + i`` and its own entry is
        dropped (``m_g is v``, prmPlanner.py:94-95): k_out = k_in - 1.  With ``queries``
        ([m][d] CUDA tensor of external points) all k_in columns are kept."""
        import torch

        self.ctx.use_torch_stream()
        nbr_in = nbr_in.contiguous()
        m, k_in = nbr_in.shape
        k_out = k_in if queries is not None else k_in - 1
        dev = nbr_in.device
        nbr = torch.empty((m, k_out), dtype=torch.int64, device=dev)
        P = torch.empty((m * k_out, self.dim), dtype=torch.float64, device=dev)
        Q = torch.empty((m * k_out, self.dim), dtype=torch.float64, device=dev)
        valid = torch.empty(m * k_out, dtype=torch.uint8, device=dev)
        X = as_device_f64(queries, self.dim, device=self.ctx.device) if queries is not None else None
        check(_lib.load().sbp_nodes_knn_edges(self._h, ptr(nbr_in), m, k_in, int(first_row), ptr(X),
                                              ptr(nbr), ptr(P), ptr(Q), ptr(valid)),
              "sbp_nodes_knn_edges")
        return nbr, P, Q, valid

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None:
                _lib.load().sbp_nodes_destroy(self._h)
                self._h = None
        except Exception:
            pass
