"""Synthetic workloads of BASELINE.json configs[3] / configs[4], shared
by bench.py, tools/ and the -m gpu tests.  Harness code, not product.

cfg5: a seeded (default_rng(32768)) recursive-backtracker maze on a 256 x 256 cell grid, cell
pitch 128 px, walls 16 px thick (the style of maps/maze1.png) = a 32 768^2 occupancy grid;
1 M uniform free nodes (default_rng(2)); 4 096 start / goal pairs (default_rng(3)).

The maze is defined by its passage arrays (`east`, `south`); pixel (x, y) is free iff it lies
in the open interior of its cell or in a wall strip that a passage opens.  That predicate is
evaluated three ways, all from the same arrays: analytically for sample points (numpy, no
raster needed), as a device raster for the GPU checker (torch, 1 GiB uint8 -> bit-packed by
libsbpgpu), and as a host raster for the reference's NumPy checker (`_img`, uint8).
"""
from __future__ import annotations

import numpy as np


class Maze:
    def __init__(self, cells: int = 256, pitch: int = 128, wall: int = 16, seed: int = 32768):
        self.cells, self.pitch, self.wall, self.seed = cells, pitch, wall, seed
        self.W = self.H = cells * pitch
        rng = np.random.default_rng(seed)
        east = np.zeros((cells, cells), bool)      # passage between (cx, cy) and (cx + 1, cy)
        south = np.zeros((cells, cells), bool)     # passage between (cx, cy) and (cx, cy + 1)
        seen = np.zeros((cells, cells), bool)
        stack = [(0, 0)]
        seen[0, 0] = True
        steps = ((1, 0), (-1, 0), (0, 1), (0, -1))
        while stack:
            cx, cy = stack[-1]
            nb = [(cx + dx, cy + dy) for dx, dy in steps
                  if 0 <= cx + dx < cells and 0 <= cy + dy < cells and not seen[cx + dx, cy + dy]]
            if not nb:
                stack.pop()
                continue
            nx, ny = nb[rng.integers(len(nb))]
            if nx != cx:
                east[min(cx, nx), cy] = True
            else:
                south[cx, min(cy, ny)] = True
            seen[nx, ny] = True
            stack.append((nx, ny))
        self.east, self.south = east, south

    # ---- the occupancy predicate -------------------------------------------------------
    def free_at(self, x, y):
        """bool array: pixel (x, y) (integer arrays, in range) is free."""
        p, h, n = self.pitch, self.wall // 2, self.cells
        x, y = np.asarray(x, dtype=np.int64), np.asarray(y, dtype=np.int64)
        lx, cx, ly, cy = x % p, x // p, y % p, y // p
        in_x, in_y = (lx >= h) & (lx < p - h), (ly >= h) & (ly < p - h)
        f = in_x & in_y
        f |= (lx >= p - h) & in_y & self.east[cx, cy]
        f |= (lx < h) & in_y & (cx > 0) & self.east[np.maximum(cx - 1, 0), cy]
        f |= in_x & (ly >= p - h) & self.south[cx, cy]
        f |= in_x & (ly < h) & (cy > 0) & self.south[cx, np.maximum(cy - 1, 0)]
        return f

    def free_points(self, P):
        """ImgCollisionChecker.feasible semantics for in-range float points [n][2]."""
        P = np.asarray(P)
        return self.free_at(np.trunc(P[:, 0]).astype(np.int64), np.trunc(P[:, 1]).astype(np.int64))

    def raster_host(self, dtype=np.uint8):
        """The reference checker's `_img` ([W][H], 1 = free) as a host array (1 GiB at 32k^2),
        assembled from the 16 possible cell tiles (which of the four wall strips are open)."""
        p, n = self.pitch, self.cells
        lx = np.arange(p)
        tiles = np.empty((16, p, p), dtype)
        for code in range(16):
            # free_at() of a stand-alone cell whose passages are the bits of `code`
            h = self.wall // 2
            in_a = (lx >= h) & (lx < p - h)
            f = in_a[:, None] & in_a[None, :]
            if code & 1:
                f = f | ((lx >= p - h)[:, None] & in_a[None, :])          # east strip
            if code & 2:
                f = f | ((lx < h)[:, None] & in_a[None, :])               # west strip
            if code & 4:
                f = f | (in_a[:, None] & (lx >= p - h)[None, :])          # south strip
            if code & 8:
                f = f | (in_a[:, None] & (lx < h)[None, :])               # north strip
            tiles[code] = f
        west = np.zeros((n, n), bool)
        west[1:] = self.east[:-1]
        north = np.zeros((n, n), bool)
        north[:, 1:] = self.south[:, :-1]
        code = self.east.astype(np.int64) | (west << 1) | (self.south.astype(np.int64) << 2) | (north << 3)
        out = np.empty((n, p, n, p), dtype)
        for cx0 in range(0, n, 16):                                       # bounded temporaries
            out[cx0:cx0 + 16] = tiles[code[cx0:cx0 + 16]].transpose(0, 2, 1, 3)
        return out.reshape(self.W, self.H)

    def raster_device(self, dev):
        """The same grid as a CUDA uint8 tensor [W][H] (for ImgCollisionChecker.from_free_mask)."""
        import torch

        p, h, W = self.pitch, self.wall // 2, self.W
        free = torch.empty((W, W), dtype=torch.uint8, device=dev)
        east_d, south_d = torch.from_numpy(self.east).to(dev), torch.from_numpy(self.south).to(dev)
        ys = torch.arange(W, device=dev)
        ly, cy = ys % p, ys // p
        in_y = (ly >= h) & (ly < p - h)
        for x0 in range(0, W, 1024):
            xs = torch.arange(x0, x0 + 1024, device=dev)
            lx, cx = xs % p, xs // p
            in_x = (lx >= h) & (lx < p - h)
            f = in_x[:, None] & in_y[None, :]
            e_here = east_d[cx][:, cy]
            e_left = east_d[(cx - 1).clamp(min=0)][:, cy] & (cx > 0)[:, None]
            f |= (lx >= p - h)[:, None] & in_y[None, :] & e_here
            f |= (lx < h)[:, None] & in_y[None, :] & e_left
            s_here = south_d[cx][:, cy]
            s_up = south_d[cx][:, (cy - 1).clamp(min=0)] & (cy > 0)[None, :]
            f |= in_x[:, None] & (ly >= p - h)[None, :] & s_here
            f |= in_x[:, None] & (ly < h)[None, :] & s_up
            free[x0:x0 + 1024] = f.to(torch.uint8)
        return free

    def free_samples(self, n: int, seed: int):
        """First n accepted of default_rng(seed).random((., 2)) * [W, H] filtered by feasible."""
        rng = np.random.default_rng(seed)
        out, have = [], 0
        while have < n:
            c = rng.random((max(4096, n + n // 4), 2)) * [self.W, self.H]
            c = c[self.free_points(c)]
            out.append(c)
            have += len(c)
        return np.ascontiguousarray(np.concatenate(out)[:n])


CFG5 = dict(cells=256, pitch=128, wall=16, seed=32768, nodes=1_000_000, node_seed=2, k=10,
            queries=4096, query_seed=3)


def cfg5_problem(scale: float = 1.0):
    """(maze, nodes [n][2], start [q][2], goal [q][2]) of cfg5.  `scale` < 1 shrinks the maze's
    cell count and the node / query counts together (tests; density per pixel unchanged)."""
    cells = max(4, int(round(CFG5["cells"] * scale)))
    mz = Maze(cells, CFG5["pitch"], CFG5["wall"], CFG5["seed"])
    area = (cells / CFG5["cells"]) ** 2
    n = max(2048, int(round(CFG5["nodes"] * area)))
    q = max(16, int(round(CFG5["queries"] * area))) if scale < 1.0 else CFG5["queries"]
    nodes = mz.free_samples(n, CFG5["node_seed"])
    sg = mz.free_samples(2 * q, CFG5["query_seed"])
    return mz, nodes, sg[:q].copy(), sg[q:].copy()


def cfg5_on_device(dev, scale: float = 1.0):
    """cfg5 on a CUDA device: (maze, checker, node store, nodes [n][2] CUDA, S, T host arrays)."""
    import os
    import sys

    import torch

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import sbp_env_b200 as sg

    mz, nodes, S, T = cfg5_problem(scale)
    free = mz.raster_device(dev)
    cc = sg.ImgCollisionChecker.from_free_mask(free, device=dev.index or 0)
    del free
    torch.cuda.empty_cache()
    nd = torch.from_numpy(nodes).to(dev)
    tree = sg.NodeStore(2, capacity=len(nodes), device=dev.index or 0)
    tree.extend(nd)
    return mz, cc, tree, nd, S, T
