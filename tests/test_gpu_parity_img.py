"""-m gpu: the CUDA path (through the C ABI) against the golden fixtures and the
oracle for the 2-D image-space checker.  Bit-exact booleans and pixels."""
import copy
import pickle
from unittest.mock import MagicMock

import numpy as np
import pytest

import c_oracle as orc
from _util import golden, kat_png, load_map, map_png, unpack
from sbp_env_b200.runtime import tune  # noqa: E402

pytestmark = pytest.mark.gpu
EPS = 1e-5


def _pair(a, b):
    return np.array(a, float) + EPS, np.array(b, float) - EPS


@pytest.fixture(scope="module")
def sg():
    import sbp_env_b200

    return sbp_env_b200


@pytest.fixture(scope="module")
def kat(sg):
    return sg.ImgCollisionChecker(kat_png())


def test_reference_kat(sg, kat):
    # tests/test_image_space_collision_checker.py:33-60 of the reference, verbatim cases
    target = (np.array([[255, 255, 0, 0, 100, 200]] * 4) == 255).astype(kat.image.dtype)
    assert np.isclose(kat.image, target).all()
    assert kat.get_coor_before_collision(*_pair([0, 2], [2, 6])) == (1, 4)
    assert kat.get_coor_before_collision(*_pair([0, 2], [2, 4])) == (1, 3)
    assert kat.visible(*_pair([0, 2], [2, 4]))
    assert kat.visible(*_pair([0, 2], [0, 4]))
    assert not kat.visible(*_pair([0, 2], [2, 6]))
    assert not kat.visible(*_pair([0, 0], [4, 0]))
    for p in [(1.5, 0.5), (1.5, 3.5), (1.78, 2.5)]:
        assert kat.feasible(p)
    for p in [(3.5, 0.5), (-1.5, 3.5), (1.78, 4.5)]:
        assert not kat.feasible(p)
    assert sg.ImgCollisionChecker.get_line((0, 0), (3, 4)) == [(0, 0), (1, 1), (1, 2), (2, 3), (3, 4)]
    assert sg.ImgCollisionChecker.get_line((3, 4), (0, 0)) == [(3, 4), (2, 3), (1, 2), (1, 1), (0, 0)]


def test_wraparound_and_errors(kat):
    # numpy negative-index wrap-around, NaN / inf behaviour
    assert kat.feasible((-6.0, 0.5))          # wraps to x = 0
    assert not kat.feasible((-7.0, 0.5))      # IndexError -> False
    with pytest.raises(ValueError):
        kat.feasible((float("nan"), 0.5))
    with pytest.raises(OverflowError):
        kat.feasible((float("inf"), 0.5))
    with pytest.raises(OverflowError):
        kat.feasible((2.0 ** 63, 0.5))        # numpy quirk pinned by the golden fixture
    assert not kat.feasible((2.0 ** 64, 0.5))
    assert kat.visible((float("nan"), 0.5), (1.0, 1.0)) is False
    with pytest.raises(OverflowError):
        kat.visible((float("inf"), 0.5), (1.0, 1.0))
    with pytest.raises(ValueError):
        kat.get_coor_before_collision((float("nan"), 0.5), (1.0, 1.0))


@pytest.mark.parametrize("name", ["room1", "intel_lab", "maze1", "noise", "kat4x6"])
def test_golden_img(sg, name):
    g = golden("img_cc")
    cc = sg.ImgCollisionChecker(kat_png() if name == "kat4x6" else map_png(name))
    assert cc._img.shape == load_map(name).shape
    P, Q = g[name + "_P"], g[name + "_Q"]
    assert np.array_equal(cc.visible_batch(P, Q), unpack(g[name + "_visible"], len(P)))
    pts, st = g[name + "_pts"], g[name + "_feasible_status"]
    ok = st == 0
    assert np.array_equal(cc.feasible_batch(pts[ok]), unpack(g[name + "_feasible"], len(pts))[ok])
    with pytest.raises(OverflowError):
        cc.feasible_batch(pts[~ok])
    sel = g[name + "_cbc_sel"]
    assert np.array_equal(cc.first_blocked_batch(P[sel], Q[sel]), g[name + "_cbc"])


@pytest.mark.parametrize("group,coop", [(4, 8), (8, 8), (16, 8), (32, 8), (0, 1), (0, 4), (0, 8), (0, 16), (0, 33)])
@pytest.mark.parametrize("force_global", [0, 1])
def test_all_kernel_variants_agree(sg, group, coop, force_global):
    """Every kernel variant -- adaptive (group 0) from pure lane-per-edge (coop 1) to pure
    warp-per-edge (coop 33), the lane-group kernels, map in smem / in L2 -- gives the
    same bits."""
    from sbp_env_b200 import _lib

    g = golden("img_cc")
    cc = sg.ImgCollisionChecker(map_png("room1"))
    P, Q = g["room1_P"], g["room1_Q"]
    P, Q = np.tile(P, (8, 1)), np.tile(Q, (8, 1))
    lib = _lib.load()
    try:
        tune("visible_group", group)
        tune("coop_below", coop)
        tune("force_global", force_global)
        got = cc.visible_batch(P, Q)
    finally:
        tune("visible_group", 0)
        tune("coop_below", 12)
        tune("force_global", 0)
    assert np.array_equal(got, np.tile(unpack(g["room1_visible"], len(g["room1_P"])), 8))


@pytest.mark.parametrize("name", ["room1", "intel_lab", "maze1", "maze2", "4d", "noise", "test", "blank"])
def test_random_edges_vs_oracle(sg, name):
    free = load_map(name)
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png(name))
    om = orc.Map(free)
    rng = np.random.default_rng(hash(name) % 2 ** 31)
    n = 300_000
    P = (rng.random((n, 2)) * 1.2 - 0.1) * [W, H]
    Q = (rng.random((n, 2)) * 1.2 - 0.1) * [W, H]
    short = slice(0, n // 2)
    Q[short] = P[short] + rng.standard_normal((n // 2, 2)) * 8
    P[-1000:] -= [W, H]                       # negative (wrapping) region
    assert np.array_equal(cc.visible_batch(P, Q), om.visible2d(P, Q))
    assert np.array_equal(cc.feasible_batch(P), om.feasible2d(P))
    sel = slice(0, 20000)
    assert np.array_equal(cc.first_blocked_batch(P[sel], Q[sel]), om.coor_before_collision(P[sel], Q[sel]))


def test_device_pointer_path_and_counter(sg):
    import torch

    free = load_map("maze1")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("maze1"))
    om = orc.Map(free)
    rng = np.random.default_rng(5)
    n = 100_000
    P, Q = rng.random((n, 2)) * [W, H], rng.random((n, 2)) * [W, H]
    dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda()
    exp, full, tested = om.visible2d(P, Q, return_counts=True)
    assert np.array_equal(cc.visible_batch(dP, dQ).cpu().numpy(), exp)
    assert np.array_equal(cc.feasible_batch(dP).cpu().numpy(), om.feasible2d(P))
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    cc._visible_device(dP, dQ, px_tested=cnt)
    got = int(cnt.item())
    # the kernel tests pixels in chunks of G lanes: between the reference's
    # until-first-hit count and the full line length
    assert got <= int(full.sum())
    assert got > 0


def test_synthetic_large_grid_l2_path(sg):
    """A 4096^2 grid (SURVEY.md 8(d) style blocky noise, plus a region of per-pixel noise and thin
    diagonal walls so that many 8x8 tiles are MIXED): too big for shared memory, so the L2 / global
    paths are exercised -- the two-level walk on the tile codes (default) and the pixel walk --
    built on the device, checked against the oracle, adversarial endpoints included."""
    import torch

    rng = np.random.default_rng(4096)
    coarse = rng.random((64, 64)) < 0.75
    free = np.kron(coarse, np.ones((64, 64), bool))
    free[1024:1536, 2048:2560] &= rng.random((512, 512)) < 0.97          # per-pixel noise
    i = np.arange(1500)
    free[2000 + i, 500 + i // 2] = False                                  # thin diagonal walls
    free[300 + i // 3, 1000 + i] = False
    W, H = free.shape
    cc = sg.ImgCollisionChecker.from_free_mask(torch.from_numpy(free.astype(np.uint8)).cuda())
    assert not cc.device_map.info()["smem_resident"]
    om = orc.Map(free)
    n = 300_000
    P, Q = rng.random((n, 2)) * [W, H], rng.random((n, 2)) * [W, H]
    Q[: n // 2] = P[: n // 2] + rng.standard_normal((n // 2, 2)) * 40
    P[n // 2: n // 2 + 40000] = rng.random((40000, 2)) * 512 + [1024, 2048]     # inside the noisy region
    Q[n // 2: n // 2 + 40000] = P[n // 2: n // 2 + 40000] + rng.standard_normal((40000, 2)) * 12
    P[-20000:] = (rng.random((20000, 2)) * 2.4 - 1.2) * [W, H]            # wrap-around / out of range
    Q[-10000:] = (rng.random((10000, 2)) * 2.4 - 1.2) * [W, H]
    P[-30000:-20000] = np.floor(P[-30000:-20000])                         # integer endpoints
    Q[-30000:-25000] = P[-30000:-25000]                                   # zero-length edges
    want = om.visible2d(P, Q)
    assert 0.02 < want.mean() < 0.9
    try:
        for tiles, rect in ((1, 1), (1, 0), (0, 1)):
            sg.runtime.tune("visible_tiles", tiles)
            sg.runtime.tune("tile_rect", rect)       # (free-run planes: bounding box of tiles all free)
            assert np.array_equal(cc.visible_batch(P, Q), want), (tiles, rect)
            dP, dQ = torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda()
            assert np.array_equal(cc.visible_batch(dP, dQ).cpu().numpy(), want), (tiles, rect)
    finally:
        sg.runtime.tune("visible_tiles", 1)
        sg.runtime.tune("tile_rect", 1)
    assert np.array_equal(cc.visible_batch(Q, P), want)                   # direction symmetry
    assert np.array_equal(cc.feasible_batch(P), om.feasible2d(P))


def test_properties_at_scale(sg):
    """Size-independent properties on 4M edges: direction symmetry of the 2-D line
    (SURVEY.md section 7), blank map => visible iff both endpoints in range, and
    idempotence."""
    free = load_map("intel_lab")
    W, H = free.shape
    cc = sg.ImgCollisionChecker(map_png("intel_lab"))
    rng = np.random.default_rng(11)
    n = 4_000_000
    P, Q = rng.random((n, 2)) * [W, H], rng.random((n, 2)) * [W, H]
    a = cc.visible_batch(P, Q)
    assert np.array_equal(a, cc.visible_batch(Q, P))
    assert np.array_equal(a, cc.visible_batch(P, Q))
    # visible => both endpoints feasible
    fp, fq = cc.feasible_batch(P), cc.feasible_batch(Q)
    assert not (a & ~(fp & fq)).any()
    # zero-length edges == feasible
    assert np.array_equal(cc.visible_batch(P, P), fp)
    allfree = sg.ImgCollisionChecker.from_free_mask(np.ones((W, H), np.uint8))
    P2 = (rng.random((100000, 2)) * 1.4 - 0.2) * [W, H]
    Q2 = (rng.random((100000, 2)) * 1.4 - 0.2) * [W, H]
    inr = lambda A: (A[:, 0] > -W - 1) & (A[:, 0] < W) & (A[:, 1] > -H - 1) & (A[:, 1] < H)
    assert np.array_equal(allfree.visible_batch(P2, Q2), inr(P2) & inr(Q2))


def test_stats_mock_deepcopy_pickle(sg):
    sg.Stats.clear_instance()
    cc = sg.ImgCollisionChecker(kat_png())
    st = sg.Stats.get_instance()
    assert cc.stats is st
    cc.visible((0.5, 2.5), (1.5, 3.5))
    cc.feasible((1.5, 0.5))
    assert (st.visible_cnt, st.feasible_cnt) == (1, 1)
    cc.visible_batch(np.zeros((7, 2)), np.ones((7, 2)))
    cc.feasible_batch(np.zeros((5, 2)))
    assert (st.visible_cnt, st.feasible_cnt) == (8, 6)
    cc2 = copy.deepcopy(cc)                                    # tests/test_rrtPlanner.py:17
    assert cc2.visible((0.5, 2.5), (1.5, 3.5))
    assert cc2.device_map is cc.device_map
    cc3 = pickle.loads(pickle.dumps(cc))
    assert cc3.visible((0.5, 2.5), (1.5, 3.5)) and not cc3.feasible((3.5, 0.5))
    cc2.visible = MagicMock(return_value=False)                # tests/test_rrtPlanner.py:31-33
    assert not cc2.visible_batch(np.zeros((3, 2)), np.ones((3, 2))).any()
    assert cc2.visible.call_count == 3


def test_empty_batches(kat):
    assert kat.visible_batch(np.zeros((0, 2)), np.zeros((0, 2))).shape == (0,)
    assert kat.feasible_batch(np.zeros((0, 2))).shape == (0,)


def test_prescreen_bucket_answers_feasible_without_launch(sg):
    """cc.prescreen(bucket) is one feasible_batch launch; feasible(p) of a
    row is then answered from the remembered booleans (same value, same Stats count), other
    points still go to the device, and a bucket with a NaN row is refused like feasible(NaN)."""
    from sbp_env_b200 import samplers

    name = "room1"
    eng = sg.ImageEngine(None, map_png(name))
    cc = eng.cc
    om = orc.Map(load_map(name))
    rng = np.random.default_rng(11)
    W, H = load_map(name).shape
    bucket = rng.random((10000, 2)) * [W, H]
    want = om.feasible2d(bucket)
    launches = []
    inner = cc._feasible_host
    cc._feasible_host = lambda P: (launches.append(len(P)), inner(P))[1]
    try:
        got = cc.prescreen(bucket)
        assert np.array_equal(got, want) and launches == [10000]
        before = cc.stats.feasible_cnt
        for i in (0, 17, 9999, 4242):
            assert bool(cc.feasible(bucket[i])) == bool(want[i])
        assert launches == [10000] and cc.stats.feasible_cnt == before + 4
        assert bool(cc.feasible((100.5, 100.25))) == bool(om.feasible2d(np.array([[100.5, 100.25]]))[0])
        assert launches == [10000, 1]
        bad = bucket.copy()
        bad[3, 0] = np.nan
        with pytest.raises(ValueError):
            cc.prescreen(bad)
        with pytest.raises(ValueError):
            cc.feasible(bad[3])

        # the hook on a RandomnessManager look-alike: every redraw pre-screens the new bucket
        class RM:
            def __init__(self):
                self.random_draws = {}

            def redraw(self, method):
                self.random_draws[method] = rng.random((500, 2))

        assert eng.upper.tolist() == [W, H]
        rm = samplers.attach_prescreen(RM(), eng)
        rm.redraw("pseudo_random")
        assert launches[-1] == 500
        n0 = len(launches)
        p = eng.transform(rm.random_draws["pseudo_random"][-1])
        assert bool(cc.feasible(p)) == bool(om.feasible2d(p.reshape(1, 2))[0]) and len(launches) == n0
    finally:
        del cc._feasible_host
