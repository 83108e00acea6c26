"""The PRM roadmap on the device.

Replaces the networkx ``DiGraph`` of ``PRMPlanner`` (``/root/reference/sbp_env/planners/
prmPlanner.py``): ``build_graph`` (:80-101) adds the directed edge ``m_g -> v`` for every
visible candidate pair, ``get_solution`` (:128-154) asks ``nx.has_path`` /
``nx.shortest_path(graph, start, goal)`` -- unweighted, i.e. the path with the fewest edges
-- and accumulates ``engine.dist`` along it.  Here the candidate blocks of
:func:`sbp_env_b200.planning.prm_connect_knn` (or any edge list) become a CSR graph in three
launches, and whole batches of (start, goal) queries are answered by one launch of a
level-synchronous BFS (one CTA per query).

Parity: reachability and hop counts are exactly networkx's.  Among several fewest-hop paths
networkx returns an adjacency-order dependent one; :meth:`Roadmap.shortest_paths` returns the
one in which every node has the lowest-index predecessor on the previous BFS level
(deterministic).  Costs are accumulated with the reference's own ``engine.dist`` on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check
from .engine import euclidean_dist
from .runtime import _is_torch_tensor, get_context, ptr


class Roadmap:
    def __init__(self, n_nodes: int, device: int = 0):
        self.ctx = get_context(device)
        self.n_nodes = int(n_nodes)
        h = C.c_void_p()
        check(_lib.load().sbp_graph_create(self.ctx.handle, self.n_nodes, C.byref(h)), "sbp_graph_create")
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None and _lib is not None:
            try:
                _lib.load().sbp_graph_destroy(h)
            except Exception:
                pass

    # ---- construction --------------------------------------------------------------
    @classmethod
    def from_knn(cls, nbr, vis, n_nodes: int = None, first_row: int = 0, device: int = None,
                 undirected: bool = False):
        """Roadmap of the candidate blocks ``nbr`` / ``vis`` ([m][k], CUDA or numpy): arc
        ``nbr[i][j] -> first_row + i`` wherever ``vis[i][j]`` (prmPlanner.py:91-101).

        ``undirected=True`` adds the mirrored arcs as well: the reference connects by radius,
        which is symmetric, so its DiGraph holds both ``m_g -> v`` and ``v -> m_g``; a k-NN
        candidate set is not symmetric by itself."""
        if device is None:
            device = nbr.device.index if _is_torch_tensor(nbr) else 0
        n = int(n_nodes) if n_nodes is not None else int(first_row + nbr.shape[0])
        g = cls(n, device)
        g.build_from_knn(nbr, vis, first_row, undirected=undirected)
        return g

    @classmethod
    def from_packed(cls, packed, n_nodes: int = None, first_row: int = 0, device: int = None,
                    undirected: bool = False):
        """The same from the packed adjacency ``(nbr + 1) * 2 + vis`` (CUDA int32 [m][k]): the
        form the connection step writes and the multi-GPU exchange assembles on every rank."""
        if device is None:
            device = packed.device.index if _is_torch_tensor(packed) else 0
        n = int(n_nodes) if n_nodes is not None else int(first_row + packed.shape[0])
        g = cls(n, device)
        g.build_from_packed(packed, first_row, undirected=undirected)
        return g

    def _dev(self, a, dtype):
        import torch

        if not _is_torch_tensor(a):
            a = torch.from_numpy(np.ascontiguousarray(a))
        return a.to(device=torch.device("cuda", self.ctx.device), dtype=dtype).contiguous()

    def build_from_knn(self, nbr, vis, first_row: int = 0, undirected: bool = False):
        import torch

        self.ctx.use_torch_stream()
        nbr = self._dev(nbr, torch.int64)
        vis = self._dev(vis, torch.uint8)
        m, k = nbr.shape
        if tuple(vis.shape) != (m, k):
            raise ValueError("nbr and vis must have the same [m][k] shape")
        check(_lib.load().sbp_graph_build_knn(self._h, ptr(nbr), ptr(vis), m, k, int(first_row),
                                              int(bool(undirected))), "sbp_graph_build_knn")
        return self

    def build_from_packed(self, packed, first_row: int = 0, undirected: bool = False, row_ids=None):
        """``row_ids`` (int32 [m], CUDA): row r of ``packed`` belongs to node ``row_ids[r]`` -- the
        cell-sorted rows and ``order_out`` of ``prm_connect_knn(..., order_out=...)``."""
        import torch

        self.ctx.use_torch_stream()
        packed = self._dev(packed, torch.int32)
        if packed.dim() != 2:
            raise ValueError("packed adjacency must be [m][k]")
        m, k = packed.shape
        if row_ids is not None:
            row_ids = self._dev(row_ids, torch.int32)
            if row_ids.numel() != m:
                raise ValueError("row_ids must hold one node index per row of packed")
            check(_lib.load().sbp_graph_build_packed_rows(self._h, ptr(packed), ptr(row_ids), m, k,
                                                          int(bool(undirected))), "sbp_graph_build_packed_rows")
            return self
        check(_lib.load().sbp_graph_build_packed(self._h, ptr(packed), m, k, int(first_row),
                                                 int(bool(undirected))), "sbp_graph_build_packed")
        return self

    def build_from_edges(self, src, dst, keep=None, undirected: bool = False):
        """Arcs ``src[e] -> dst[e]`` (e.g. the output of
        :func:`sbp_env_b200.planning.prm_connect_radius`), ``keep`` an optional mask."""
        import torch

        self.ctx.use_torch_stream()
        src, dst = self._dev(src, torch.int64), self._dev(dst, torch.int64)
        if src.shape != dst.shape:
            raise ValueError("src and dst must have the same length")
        keep = None if keep is None else self._dev(keep, torch.uint8)
        check(_lib.load().sbp_graph_build_edges(self._h, ptr(src), ptr(dst), ptr(keep), src.numel(),
                                                int(bool(undirected))), "sbp_graph_build_edges")
        return self

    def components(self):
        """``(labels int32 [n] CUDA, n_components)``: weakly connected components, label =
        smallest node index of the component (``nx.has_path`` needs equal labels)."""
        import torch

        self.ctx.use_torch_stream()
        labels = torch.empty(self.n_nodes, dtype=torch.int32, device=torch.device("cuda", self.ctx.device))
        nc = C.c_int64()
        check(_lib.load().sbp_graph_components(self._h, ptr(labels), C.byref(nc)), "sbp_graph_components")
        return labels, int(nc.value)

    # ---- inspection ----------------------------------------------------------------
    @property
    def n_edges(self) -> int:
        n, e = C.c_int64(), C.c_int64()
        check(_lib.load().sbp_graph_info(self._h, C.byref(n), C.byref(e)), "sbp_graph_info")
        return int(e.value)

    def csr(self):
        """``(row_ptr int64 [n + 1], col int32 [E])`` CUDA tensors; row u lists the heads of
        the edges u -> v in unspecified order."""
        import torch

        self.ctx.use_torch_stream()
        e = self.n_edges
        dev = torch.device("cuda", self.ctx.device)
        row_ptr = torch.empty(self.n_nodes + 1, dtype=torch.int64, device=dev)
        col = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        check(_lib.load().sbp_graph_csr(self._h, ptr(row_ptr), ptr(col), col.numel()), "sbp_graph_csr")
        return row_ptr, col[:e]

    def edges(self):
        """Sorted numpy array [E][2] of the directed edges (src, dst)."""
        row_ptr, col = self.csr()
        rp, c = row_ptr.cpu().numpy(), col.cpu().numpy().astype(np.int64)
        src = np.repeat(np.arange(self.n_nodes, dtype=np.int64), np.diff(rp))
        e = np.stack([src, c], 1)
        return e[np.lexsort((e[:, 1], e[:, 0]))]

    # ---- queries -------------------------------------------------------------------
    def shortest_paths(self, start, goal, max_len: int = 0):
        """Fewest-hop paths for a batch of (start, goal) node indices in ONE launch.

        :return: ``hops`` int32 [nq] (-1 = no path, ``nx.has_path`` False) and, when
            ``max_len > 0``, ``paths`` int64 [nq][max_len] (start..goal, -1 padded; all -1
            when the path has more than ``max_len`` nodes).  CUDA tensors when the inputs
            are CUDA tensors, numpy otherwise."""
        import torch

        self.ctx.use_torch_stream()
        on_dev = _is_torch_tensor(start)
        s, t = self._dev(np.atleast_1d(start) if not on_dev else start, torch.int64), \
            self._dev(np.atleast_1d(goal) if not _is_torch_tensor(goal) else goal, torch.int64)
        if s.shape != t.shape:
            raise ValueError("start and goal must have the same length")
        nq = s.numel()
        hops = torch.empty(nq, dtype=torch.int32, device=s.device)
        paths = torch.empty((nq, max_len), dtype=torch.int64, device=s.device) if max_len > 0 else None
        check(_lib.load().sbp_graph_bfs(self._h, ptr(s), ptr(t), nq, ptr(hops), ptr(paths), int(max_len)),
              "sbp_graph_bfs")
        if not on_dev:
            hops = hops.cpu().numpy()
            paths = None if paths is None else paths.cpu().numpy()
        return (hops, paths) if max_len > 0 else hops

    def has_path(self, start, goal):
        h = self.shortest_paths(start, goal)
        return h >= 0


def prm_get_solution(cc, tree, roadmap, start_pos, goal_pos, epsilon, dist=euclidean_dist,
                     max_len: int = 4096, poses=None, exclude=()):
    """``PRMPlanner.get_solution`` (prmPlanner.py:128-154) on the device roadmap.

    start / goal are connected to the roadmap node that ``get_nearest_free`` picks among the
    nodes with ``norm(pose - pos) < epsilon`` (:103-126, :131-138); the fewest-hop path
    between the two is searched on the device and its cost accumulated on the host with
    ``engine.dist`` exactly like :145-150 (``solution_path[0].cost = dist(path[0], start) = 0``).

    :return: ``(c_max, path)`` -- ``path`` the node indices from the start's roadmap node to
            the goal's; ``(inf, None)`` when either end cannot be connected or no path exists.

    ``exclude``: node indices ``get_nearest_free`` skips by identity (:113: ``n is start_pt or n is
    goal_pt``) -- in a reference run node 0 *is* ``env.start_pt`` (rrtPlanner.py init), so a
    drop-in caller passes ``exclude=(0,)``."""
    from .planner_ops import prm_nearest_free

    s = prm_nearest_free(cc, tree, start_pos, epsilon, exclude=exclude, dist=dist, poses=poses)
    g = prm_nearest_free(cc, tree, goal_pos, epsilon, exclude=exclude, dist=dist, poses=poses)
    if s is None or g is None:
        return float("inf"), None
    hops, paths = roadmap.shortest_paths(np.array([s]), np.array([g]), max_len=max_len)
    if hops[0] < 0:
        return float("inf"), None
    if paths[0, 0] < 0:
        raise ValueError(f"path of {int(hops[0]) + 1} nodes does not fit max_len={max_len}")
    path = paths[0, : int(hops[0]) + 1]
    P = tree.poses_rows(path) if poses is None else np.asarray(poses)[path]
    cost = dist(P[0], P[0])
    for i in range(1, len(P)):
        cost = cost + dist(P[i], P[i - 1])
    return cost, path
