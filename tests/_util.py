"""Shared helpers for the test-suite: golden fixture access and the oracle."""
import io
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def golden(name):
    if name not in _cache:
        _cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return _cache[name]


def load_map(name):
    """The reference's `_img == 1` for a shipped map, bool [W][H] (tests/golden/maps.npz)."""
    g = golden("maps")
    W, H = (int(v) for v in g[name + "_shape"])
    return np.unpackbits(g[name + "_bits"])[: W * H].reshape(W, H).astype(bool)


def unpack(bits, n):
    return np.unpackbits(bits)[:n].astype(bool)


def map_png(name):
    """A PNG (file-like) whose reference binarisation equals the golden occupancy grid:
    free -> 255, blocked -> 0 (collisionChecker.py:101-108 keeps only grey == 255)."""
    from PIL import Image

    free = load_map(name)                      # [W][H]
    arr = np.where(free.T, 255, 0).astype(np.uint8)   # image is [H][W]
    f = io.BytesIO()
    Image.fromarray(arr).save(f, "png")
    f.name = name + ".png"
    f.seek(0)
    return f


def kat_png():
    """The 4x6 image of the reference's own tests (tests/common_vars.py:13-29)."""
    from PIL import Image

    arr = np.array([[255, 255, 0, 0, 100, 200]] * 4, dtype=np.uint8)
    f = io.BytesIO()
    Image.fromarray(arr).save(f, "png")
    f.name = "test.png"
    f.seek(0)
    return f


def draw_free_samples(free, n, seed, feasible_fn):
    """First n accepted of default_rng(seed).random((., 2)) * [W, H] filtered by feasible
    (SURVEY.md section 8(d), cfg2)."""
    W, H = free.shape
    rng = np.random.default_rng(seed)
    out, have = [], 0
    while have < n:
        c = rng.random((2 * n, 2)) * [W, H]
        c = c[feasible_fn(c)]
        out.append(c)
        have += len(c)
    return np.ascontiguousarray(np.concatenate(out)[:n])


def reference_present():
    return os.path.isdir("/root/reference/sbp_env")
