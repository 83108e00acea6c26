"""The free-run planes of the tile codes (sbp_env_b200/csrc/map.cu:k_tile_runs0 / k_tile_runs_next,
lookup device_fns.cuh:tiles_rect_free), restated in numpy with the SAME layout and the same two-word
test, against the definition "every 8x8 tile of the rectangle is all free" on random maps -- the
argument DESIGN.md 3a rests on (windows of 32 tile rows every 16, two runs of 2^j tiles covering a span
of up to 15 columns), and the line property that makes the box test sound: every pixel of the
reference's Bresenham line (collisionChecker.py:155-215) lies in the bounding box of its end pixels.
Host logic only; the CUDA path is compared with the oracle in tests/test_gpu_prm_connect.py."""
import numpy as np


def build_planes(tile_free):
    """tile_free [TX][TY] bool -> planes [4][RK][TX] uint32 (bit b of [j][k][u] = tiles (u..u+2^j-1, 16k+b) free)."""
    TX, TY = tile_free.shape
    RK = (TY + 15) // 16
    planes = np.zeros((4, RK, TX), np.uint64)
    for k in range(RK):
        for b in range(32):
            v = 16 * k + b
            if v < TY:
                planes[0, k, :] |= tile_free[:, v].astype(np.uint64) << np.uint64(b)
    for j in range(1, 4):
        half = 1 << (j - 1)
        planes[j, :, :TX - half] = planes[j - 1, :, :TX - half] & planes[j - 1, :, half:]
        planes[j, :, TX - half:] = 0
    return planes.astype(np.uint32)


def rect_free(planes, tu0, tu1, tv0, tv1):
    """device_fns.cuh:tiles_rect_free."""
    w, h = tu1 - tu0 + 1, tv1 - tv0 + 1
    if w > 15 or h > 16:
        return False
    j = w.bit_length() - 1
    mask = ((1 << h) - 1) << (tv0 & 15)
    row = planes[j, tv0 >> 4]
    a, b = int(row[tu0]), int(row[tu1 + 1 - (1 << j)])
    return (a & b & mask) == mask


def test_two_word_test_equals_the_definition():
    rng = np.random.default_rng(5)
    for TX, TY, p in ((37, 53, 0.9), (64, 64, 0.97), (130, 17, 0.8), (16, 16, 1.0)):
        tile_free = rng.random((TX, TY)) < p
        planes = build_planes(tile_free)
        decided = 0
        for _ in range(4000):
            tu0, tv0 = int(rng.integers(TX)), int(rng.integers(TY))
            tu1 = min(TX - 1, tu0 + int(rng.integers(0, 18)))
            tv1 = min(TY - 1, tv0 + int(rng.integers(0, 19)))
            want = bool(tile_free[tu0:tu1 + 1, tv0:tv1 + 1].all())
            got = rect_free(planes, tu0, tu1, tv0, tv1)
            if tu1 - tu0 + 1 <= 15 and tv1 - tv0 + 1 <= 16:
                assert got == want, (TX, TY, tu0, tu1, tv0, tv1)
                decided += 1
            else:
                assert got is False                     # spans the planes do not cover: never "free"
        assert decided > 2000


def reference_line(x1, y1, x2, y2):
    """ImgCollisionChecker.get_line (collisionChecker.py:155-215), restated."""
    dx, dy = x2 - x1, y2 - y1
    steep = abs(dy) > abs(dx)
    if steep:
        x1, y1, x2, y2 = y1, x1, y2, x2
    if x1 > x2:
        x1, x2, y1, y2 = x2, x1, y2, y1
    dx, dy = x2 - x1, y2 - y1
    error = int(dx / 2.0)
    ystep = 1 if y1 < y2 else -1
    y, pts = y1, []
    for x in range(x1, x2 + 1):
        pts.append((y, x) if steep else (x, y))
        error -= abs(dy)
        if error < 0:
            y += ystep
            error += dx
    return pts


def test_line_stays_in_the_bounding_box_of_its_end_pixels():
    rng = np.random.default_rng(9)
    for _ in range(3000):
        x1, y1, x2, y2 = (int(v) for v in rng.integers(0, 200, 4))
        pts = np.array(reference_line(x1, y1, x2, y2))
        assert pts[:, 0].min() >= min(x1, x2) and pts[:, 0].max() <= max(x1, x2)
        assert pts[:, 1].min() >= min(y1, y2) and pts[:, 1].max() <= max(y1, y2)
        ends = {tuple(pts[0]), tuple(pts[-1])}
        assert ends == {(x1, y1), (x2, y2)}
