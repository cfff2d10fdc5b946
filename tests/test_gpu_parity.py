"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes), against the oracle.

Bar: bit-exact for neighbour lists (all five arrays + segments/counts) and raw RDF bin counts, per
arithmetic flavour; 1e-5 for q_l.  Oracles: ``oracle/_ref`` (the unmodified reference compiled here,
travels as a prebuilt .so) where it exists, else ``oracle/port`` (plain-C restatement, itself pinned to
``_ref`` by tests/test_oracle_port.py).
"""
import os

import numpy as np
import pytest

from oracle import port, ref
from tests.golden.make_golden import CASES as GOLDEN_CASES
from tests.util import BOXES, assert_nlist_equal, bits, random_points

pytestmark = pytest.mark.gpu

WRAP, IMAGE = 0, 1


def _capi():
    from freud_b200 import _capi

    return _capi


def oracle_ball(flavour, box, pts, q, r_max, r_min=0.0, exclude_ii=False, sort_by_distance=False):
    """The reference itself when its compiled library travelled with the repo, else the pinned port."""
    if ref.available():
        eng = "linkcell" if flavour == WRAP else "aabb"
        query = ref.Query(eng, box, pts, is2d=box.is2D, cell_width=min(r_max, 0.4 * float(min(box.Lx, box.Ly))))
        return query.nlist(q, r_max=r_max, r_min=r_min, exclude_ii=exclude_ii, sort_by_distance=sort_by_distance)
    return port.ball_nlist(flavour, box, box.is2D, pts, q, r_max, r_min, exclude_ii, sort_by_distance)


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _assert_gold(got, gold, prefix):
    for key in ("neighbors", "segments", "counts"):
        assert np.array_equal(got[key], gold[f"{prefix}_{key}"]), f"{prefix} {key}"
    for key in ("distances", "vectors"):
        assert np.array_equal(bits(got[key]), bits(gold[f"{prefix}_{key}"])), f"{prefix} {key} differ bitwise"


@pytest.mark.parametrize("name", list(GOLDEN_CASES))
def test_golden_vectors_from_the_reference(ctx, name):
    """The committed outputs of the reference itself (tests/golden/make_golden.py), both flavours + kNN + RDF."""
    capi = _capi()
    box, n, nq, r_max, r_min, excl, seed = GOLDEN_CASES[name]
    gold = np.load(os.path.join(GOLD, f"nl_{name}.npz"))
    pts = random_points(box, n, seed)
    q = None if nq == 0 else random_points(box, nq, seed + 1000)
    dp = capi.DevicePoints(ctx, box, pts)
    _assert_gold(dp.ball_query(q, WRAP, r_max, r_min, excl).to_host(), gold, "wrap")
    _assert_gold(dp.ball_query(q, IMAGE, r_max, r_min, excl).to_host(), gold, "image")
    _assert_gold(dp.ball_query(q, IMAGE, r_max, r_min, excl, True).to_host(), gold, "image_bydist")
    _assert_gold(dp.knn_query(q, 6, exclude_ii=excl).to_host(), gold, "knn6")
    for flavour, tag in ((WRAP, "wrap"), (IMAGE, "image")):
        rdf = capi.DeviceRDF(ctx, 40, r_max, r_min)
        rdf.accumulate(dp, q, flavour, r_max, 0.0, excl)
        assert np.array_equal(rdf.read(), gold[f"rdf_{tag}_bin_counts"])


def test_golden_rdf_config0(ctx):
    """BASELINE.json configs[0] through the C ABI: raw bin counts bit-exact, one frame and two (reset=False)."""
    from freud_b200 import data

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "rdf_config0.npz"))
    rdf = capi.DeviceRDF(ctx, 100, 5.0)
    for seed, key in ((0, "bin_counts"), (1, "two_frames_bin_counts")):
        box, pts = data.make_random_system(50, 10000, seed=seed)
        rdf.accumulate(capi.DevicePoints(ctx, box, pts), None, IMAGE, 5.0, 0.0, True)
        assert np.array_equal(rdf.read(), gold[key])


def test_golden_steinhardt(ctx):
    from freud_b200 import data

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "steinhardt_fcc.npz"))
    box, pts = data.make_fcc_system(4, scale=1.2, sigma_noise=0.06, seed=7)
    dp = capi.DevicePoints(ctx, box, pts)
    nl = dp.knn_query(None, 12, exclude_ii=True)
    for ls in ([6], [4, 6], [2, 8], [12]):
        tag = "_".join(str(l) for l in ls)
        out = dp.steinhardt(nl, ls)
        assert np.allclose(out["ql"], gold[f"knn12_ql_{tag}"], rtol=1e-5, atol=1e-6)
        assert np.allclose(out["order"], gold[f"knn12_order_{tag}"], rtol=1e-4, atol=1e-6)
        for l, qlm in zip(ls, out["qlm"]):
            assert np.allclose(qlm, gold[f"knn12_qlm_{tag}_l{l}"], atol=1e-5)
    nlb = dp.ball_query(None, IMAGE, 1.05, 0.0, True)
    assert np.allclose(dp.steinhardt(nlb, [6])["ql"], gold["ball_ql_6"], rtol=1e-5, atol=1e-6)


def test_hand_built_queries_of_the_reference_tests(ctx):
    """tests/test_locality_neighbor_query.py:94-155, :218-300 through the C ABI."""
    capi = _capi()
    from freud_b200.box import Box

    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0], [2, 0, 0]], np.float32)
    dp = capi.DevicePoints(ctx, box, pts)

    def sets(nl):
        h = nl.to_host()
        return [set(int(j) for j in h["neighbors"][h["neighbors"][:, 0] == i, 1]) for i in range(4)]

    for flavour in (WRAP, IMAGE):
        assert list(dp.ball_query(None, flavour, 2.01).to_host()["counts"]) == [3, 4, 3, 4]
        assert dp.ball_query(None, flavour, 2.01, 0.0, True).num_bonds == 10
        assert sets(dp.ball_query(None, flavour, 2.9, 1.1, True)) == [{3}, {2}, {1}, {0}]
    assert sets(dp.knn_query(None, 3)) == [{0, 1, 3}, {0, 1, 3}, {1, 2, 3}, {1, 2, 3}]
    assert sets(dp.knn_query(None, 3, exclude_ii=True)) == [{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}]
    assert sets(dp.knn_query(None, 5, exclude_ii=True)) == [{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}]
    assert sets(dp.knn_query(None, 3, r_max=1.9, exclude_ii=True)) == [{1}, {0, 3}, {3}, {1, 2}]
    assert sets(dp.knn_query(None, 3, r_min=1.1, exclude_ii=True)) == [{2, 3}, {2}, {0, 1}, {0}]


@pytest.mark.parametrize("name", list(BOXES))
@pytest.mark.parametrize("flavour", [WRAP, IMAGE])
def test_ball_nlist_self_query(ctx, name, flavour):
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=11)
    dp = _capi().DevicePoints(ctx, box, pts)
    for exclude_ii in (True, False):
        got = dp.ball_query(None, flavour, r, 0.0, exclude_ii).to_host()
        want = oracle_ball(flavour, box, pts, pts, r, 0.0, exclude_ii)
        assert_nlist_equal(got, want, f"{name} flavour={flavour} exclude_ii={exclude_ii}")


@pytest.mark.parametrize("name", list(BOXES))
@pytest.mark.parametrize("flavour", [WRAP, IMAGE])
def test_ball_nlist_separate_queries_rmin_sort_by_distance(ctx, name, flavour):
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=12)
    q = random_points(box, 700, seed=13)
    dp = _capi().DevicePoints(ctx, box, pts)
    for sort_by_distance in (False, True):
        got = dp.ball_query(q, flavour, r, 1.0, False, sort_by_distance).to_host()
        want = oracle_ball(flavour, box, pts, q, r, 1.0, False, sort_by_distance)
        assert_nlist_equal(got, want, f"{name} flavour={flavour} sort_by_distance={sort_by_distance}")


@pytest.mark.parametrize("name", ["cubic", "tri1", "sq2d"])
def test_ball_points_outside_box(ctx, name):
    """Un-wrapped inputs.  AABB has no restriction (images +-1 only); LinkCell's own cell index is UB upstream
    for such points, so the WRAP flavour is checked against the port, whose
    definition is E1: brute force with Box::wrap."""
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=14, spill=0.2)
    q = random_points(box, 500, seed=15, spill=0.2)
    dp = _capi().DevicePoints(ctx, box, pts)
    got = dp.ball_query(q, IMAGE, r).to_host()
    assert_nlist_equal(got, oracle_ball(IMAGE, box, pts, q, r), f"{name} image outside")
    got = dp.ball_query(q, WRAP, r).to_host()
    assert_nlist_equal(got, port.ball_nlist(WRAP, box, box.is2D, pts, q, r), f"{name} wrap outside")


def test_ball_query_shard_offsets(ctx):
    """Contiguous query shards with q_index_offset concatenate to the single-GPU list."""
    box, n, r = BOXES["tri1"]
    pts = random_points(box, n, seed=16)
    dp = _capi().DevicePoints(ctx, box, pts)
    full = dp.ball_query(None, IMAGE, r, 0.0, True).to_host()
    parts, off = [], 0
    for lo, hi in ((0, 1000), (1000, 1001), (1001, n)):
        part = dp.ball_query(pts[lo:hi], IMAGE, r, 0.0, True, q_index_offset=lo).to_host()
        part["neighbors"][:, 0] += lo
        parts.append(part)
    for key in ("neighbors", "distances", "vectors"):
        assert np.array_equal(np.concatenate([p[key] for p in parts]), full[key]), key
    assert np.array_equal(np.concatenate([p["counts"] for p in parts]), full["counts"])


def test_empty_and_error_behaviour(ctx):
    capi = _capi()
    box, n, r = BOXES["cubic"]
    pts = random_points(box, 100, seed=17)
    dp = capi.DevicePoints(ctx, box, pts)
    # zero query points -> empty (0, 2) list, tests/test_locality_neighbor_list.py:253-256
    nl = dp.ball_query(np.zeros((0, 3), np.float32), IMAGE, r)
    assert nl.num_bonds == 0 and nl.to_host()["neighbors"].shape == (0, 2)
    # r_max <= 0, r_max <= r_min -> ValueError (NeighborQuery.h:321-328)
    with pytest.raises(ValueError):
        dp.ball_query(None, IMAGE, -1.0)
    with pytest.raises(ValueError):
        dp.ball_query(None, WRAP, 1.0, 2.0)
    # AABB flavour: r_max >= half the box -> RuntimeError (NeighborQuery.h:503-510)
    with pytest.raises(RuntimeError):
        dp.ball_query(None, IMAGE, 10.1)
    # zero particles -> ValueError (NeighborQuery.h:97-100)
    with pytest.raises(ValueError):
        capi.DevicePoints(ctx, box, np.zeros((0, 3), np.float32))
    # 2-D box with z != 0 -> ValueError (NeighborQuery.h:103-112)
    from freud_b200.box import Box

    with pytest.raises(ValueError):
        capi.DevicePoints(ctx, Box.square(10), np.array([[0, 0, 0.1]], np.float32))
    # RDF constructor validation (RDF.cc:27-42)
    for args in ((0, 5.0, 0.0), (10, -1.0, 0.0), (10, 5.0, -0.5), (10, 2.0, 3.0)):
        with pytest.raises(ValueError):
            capi.DeviceRDF(ctx, *args)


def test_cell_list_is_a_partition(ctx):
    """K1-K3: every point appears exactly once, cells are contiguous and ascending."""
    box, n, r = BOXES["tri1"]
    pts = random_points(box, n, seed=18)
    dp = _capi().DevicePoints(ctx, box, pts)
    dims = dp.build_cells(r)
    cell_start, order = dp.read_cells(dims)
    assert cell_start[0] == 0 and cell_start[-1] == n and np.all(np.diff(cell_start.astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.uint32))
    # cell thickness covers r along every axis; plane distances and fractional coordinates from the ORACLE's Box
    # (Box::getNearestPlaneDistance, Box::makeFractional: freud/box/Box.h:243-255, 468-487), not from this package's own
    _, pd = port.box_info(box, False)
    assert np.all(pd / dims > r)
    # membership: fractional coordinate of every point lies inside its cell (up to float rounding)
    frac = port.box_apply(box, False, "fractional", pts[order]).astype(np.float64)
    cell_of_slot = np.searchsorted(cell_start, np.arange(n), side="right") - 1
    cx = cell_of_slot % dims[0]
    cy = (cell_of_slot // dims[0]) % dims[1]
    cz = cell_of_slot // (dims[0] * dims[1])
    for c, f, d in ((cx, frac[:, 0], dims[0]), (cy, frac[:, 1], dims[1]), (cz, frac[:, 2], dims[2])):
        assert np.all(np.abs(f * d - (c + 0.5)) <= 0.5 + 1e-3)


@pytest.mark.parametrize("name", list(BOXES))
@pytest.mark.parametrize("flavour", [WRAP, IMAGE])
def test_rdf_bin_counts(ctx, name, flavour):
    capi = _capi()
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=21)
    dp = capi.DevicePoints(ctx, box, pts)
    rdf = capi.DeviceRDF(ctx, 50, r, 0.5)
    rdf.accumulate(dp, None, flavour, r, 0.0, True)
    want = port.rdf_accumulate(flavour, box, box.is2D, pts, pts, 50, r, 0.5, True)
    assert np.array_equal(rdf.read(), want)
    if ref.available() and flavour == IMAGE:
        R = ref.RDF(50, r, 0.5)
        R.accumulate(ref.Query("raw", box, pts, is2d=box.is2D), pts, mode="ball", r_max=r, exclude_ii=True)
        assert np.array_equal(rdf.read(), R.results()["bin_counts"])
    # reset=False semantics: a second frame adds; separate query points; then reset
    q = random_points(box, 500, seed=22)
    rdf.accumulate(dp, q, flavour, r, 0.0, False)
    want2 = port.rdf_accumulate(flavour, box, box.is2D, pts, q, 50, r, 0.5, False, counts=want.copy())
    assert np.array_equal(rdf.read(), want2)
    rdf.reset()
    assert not rdf.read().any()
    # from an existing neighbour list: one increment per stored distance
    nl = dp.ball_query(None, flavour, r, 0.0, True)
    rdf.accumulate_nlist(nl)
    assert np.array_equal(rdf.read(), want)


def test_rdf_many_bins_global_histogram(ctx):
    capi = _capi()
    box, n, r = BOXES["cubic"]
    pts = random_points(box, n, seed=23)
    dp = capi.DevicePoints(ctx, box, pts)
    bins = 20000  # > 48 KB of counters: the kernel bins straight into global memory
    rdf = capi.DeviceRDF(ctx, bins, r)
    rdf.accumulate(dp, None, IMAGE, r, 0.0, True)
    assert np.array_equal(rdf.read(), port.rdf_accumulate(IMAGE, box, False, pts, pts, bins, r, 0.0, True))


@pytest.mark.parametrize("name", list(BOXES))
def test_knn(ctx, name):
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=31)
    dp = _capi().DevicePoints(ctx, box, pts)
    q = pts[:400]
    for k, sbd in ((12, False), (6, True), (1, False)):
        got = dp.knn_query(q, k, exclude_ii=True, sort_by_distance=sbd).to_host()
        if ref.available():
            want = ref.Query("aabb", box, pts, is2d=box.is2D).nlist(q, num_neighbors=k, exclude_ii=True,
                                                                     sort_by_distance=sbd)
        else:
            want = port.knn_nlist(box, box.is2D, pts, q, k, exclude_ii=True, sort_by_distance=sbd)
        assert_nlist_equal(got, want, f"{name} knn k={k}")


GHOST = 2


@pytest.mark.parametrize("blob", [600, 3000])
def test_clustered_system_dense_tiles(ctx, blob):
    """A tile with far more candidates than the uniform estimate: the warp-cooperative kernels are retried with a hit
    buffer sized for it (blob = 600), and hand over to the general family only beyond the largest buffer (3000)."""
    from freud_b200.box import Box

    capi = _capi()
    box = Box.cube(20)
    rs = np.random.RandomState(5)
    cluster = (np.float32([3.0, -2.0, 1.0]) + 0.4 * rs.standard_normal((blob, 3))).astype(np.float32)
    pts = np.concatenate([random_points(box, 4000, seed=71), box.wrap(cluster)]).astype(np.float32)
    dp = capi.DevicePoints(ctx, box, pts)
    for flavour in (WRAP, IMAGE):
        got = dp.ball_query(None, flavour, 2.5, 0.0, True).to_host()
        assert_nlist_equal(got, port.ball_nlist(flavour, box, False, pts, pts, 2.5, 0.0, True), f"blob {blob} fl {flavour}")
    q = pts[::7]
    got = dp.knn_query(q, 8, exclude_ii=False).to_host()
    assert_nlist_equal(got, port.knn_nlist(box, False, pts, q, 8), f"blob {blob} knn")
    rdf = capi.DeviceRDF(ctx, 60, 2.5)
    rdf.accumulate(dp, None, IMAGE, 2.5, 0.0, True)
    assert np.array_equal(rdf.read(), port.rdf_accumulate(port.IMAGE, box, False, pts, pts, 60, 2.5, 0.0, True))


def test_pmftxy_over_a_neighbor_list(ctx):
    """fgpu_pmftxy_* (PMFTXY.cc:25-87) over device NeighborLists against the committed outputs of the reference: bin
    counts bit for bit (the rotation's cos/sin come from the host libm), reset=False accumulation, a histogram too
    large for shared memory."""
    from freud_b200.box import Box
    from tests.golden.make_golden import pmftxy_inputs

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "pmftxy.npz"))
    r = float(np.sqrt(3.0 ** 2 + 2.5 ** 2))
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        dp = capi.DevicePoints(ctx, box, pts)
        pm = capi.DevicePMFTXY(ctx, 3.0, 2.5, 30, 24)
        pm.accumulate_nlist(dp.ball_query(q, IMAGE, r, 0.0, False), th_q)
        assert np.array_equal(pm.read(), gold[f"{name}_query_counts"])
        pm.accumulate_nlist(dp.ball_query(q, IMAGE, r, 0.0, False), th_q)
        assert np.array_equal(pm.read(), 2 * gold[f"{name}_query_counts"])
        pm.reset()
        pm.accumulate_nlist(dp.ball_query(None, IMAGE, r, 0.0, True), th_p)
        assert np.array_equal(pm.read(), gold[f"{name}_self_counts"])
    big = capi.DevicePMFTXY(ctx, 3.0, 2.5, 150, 120)  # 72 KB of counters: global atomics
    nl_dev = dp.ball_query(None, IMAGE, r, 0.0, True)
    big.accumulate_nlist(nl_dev, th_p)
    nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, r, 0.0, True)
    assert np.array_equal(big.read(), port.pmftxy(box, 3000, nl, th_p, 3.0, 2.5, 150, 120)[0])
    with pytest.raises(ValueError):
        capi.DevicePMFTXY(ctx, 3.0, 2.5, 0, 4)
    with pytest.raises(ValueError):
        capi.DevicePMFTXY(ctx, -1.0, 2.5, 4, 4)


def test_pmft3_over_a_neighbor_list(ctx):
    """fgpu_pmft_* (PMFTXYZ.cc:111-147, PMFTXYT.cc:77-101, PMFTR12.cc:91-113) over device NeighborLists against the
    committed outputs of the reference: bin counts bit for bit.  XYZ is float arithmetic only; the angle axes of XYT / R12
    hang on libm's atan2f, which the kernel brackets, leaving bonds next to a bin edge to the host -- all of them on the
    lattice case.  Accumulation over two calls, a histogram too large for shared memory, constructor errors."""
    from freud_b200.box import Box
    from tests.golden.make_golden import PMFT3_EQUIV, pmft3_lattice, pmft3_quats, pmftxy_inputs

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "pmft3.npz"))
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 1500, 31), random_points(box, 400, 32)
    r = float(np.sqrt(2.0 ** 2 + 2.5 ** 2 + 3.0 ** 2))
    dp = capi.DevicePoints(ctx, box, pts)
    xyz = capi.DevicePMFT(ctx, capi.PMFT_XYZ, (2.0, 2.5, 3.0), (12, 10, 8))
    xyz.accumulate_nlist(dp.ball_query(q, IMAGE, r, 0.0, False), None, pmft3_quats(400, 6), PMFT3_EQUIV)
    assert np.array_equal(xyz.read(), gold["xyz_query_counts"])
    xyz.accumulate_nlist(dp.ball_query(q, IMAGE, r, 0.0, False), None, pmft3_quats(400, 6), PMFT3_EQUIV)
    assert np.array_equal(xyz.read(), 2 * gold["xyz_query_counts"]) and xyz.host_binned_bonds == 0
    xyz.reset()
    xyz.accumulate_nlist(dp.ball_query(None, IMAGE, r, 0.0, True), None, pmft3_quats(1500, 7), PMFT3_EQUIV[:1])
    assert np.array_equal(xyz.read(), gold["xyz_self_counts"])
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        dp = capi.DevicePoints(ctx, box, pts)
        xyt = capi.DevicePMFT(ctx, capi.PMFT_XYT, (3.0, 2.5), (14, 12, 9))
        xyt.accumulate_nlist(dp.ball_query(q, IMAGE, float(np.sqrt(3.0 ** 2 + 2.5 ** 2)), 0.0, False), th_p, th_q)
        assert np.array_equal(xyt.read(), gold[f"{name}_xyt_query_counts"])
        r12 = capi.DevicePMFT(ctx, capi.PMFT_R12, (4.0,), (10, 11, 12))
        r12.accumulate_nlist(dp.ball_query(None, IMAGE, 4.0, 0.0, True), th_p, th_p)
        assert np.array_equal(r12.read(), gold[f"{name}_r12_self_counts"])
    big = capi.DevicePMFT(ctx, capi.PMFT_R12, (4.0,), (20, 36, 36))  # 104 KB of counters: global atomics
    big.accumulate_nlist(dp.ball_query(None, IMAGE, 4.0, 0.0, True), th_p, th_p)
    nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, 4.0, 0.0, True)
    want_big = port.pmft3(port.PMFT_R12, box, 3000, nl, th_p, th_p, (4.0,), (20, 36, 36))[0]
    assert np.array_equal(big.read(), want_big)
    # the same histogram in slices over the shared memories of a thread-block cluster (the alternative to global
    # atomics, pmft.cu), through a NeighborList and straight from the search's bag
    ctx.set_tuning("pmft_cluster", 1)
    try:
        for route in ("nlist", "bag"):
            big = capi.DevicePMFT(ctx, capi.PMFT_R12, (4.0,), (20, 36, 36))
            if route == "nlist":
                big.accumulate_nlist(dp.ball_query(None, IMAGE, 4.0, 0.0, True), th_p, th_p)
            else:
                big.accumulate(dp, None, IMAGE, 4.0, th_p, th_p, exclude_ii=True)
            assert np.array_equal(big.read(), want_big), route
    finally:
        ctx.set_tuning("pmft_cluster", 0)
    box, pts, th = pmft3_lattice()
    dp = capi.DevicePoints(ctx, box, pts)
    xyt = capi.DevicePMFT(ctx, capi.PMFT_XYT, (3.0, 3.0), (6, 6, 8))
    xyt.accumulate_nlist(dp.ball_query(None, IMAGE, float(np.sqrt(18.0)), 0.0, True), th, th)
    assert np.array_equal(xyt.read(), gold["lattice_xyt_counts"]) and xyt.host_binned_bonds > 1000
    r12 = capi.DevicePMFT(ctx, capi.PMFT_R12, (3.0,), (6, 8, 8))
    r12.accumulate_nlist(dp.ball_query(None, IMAGE, 3.0, 0.0, True), th, th)
    assert np.array_equal(r12.read(), gold["lattice_r12_counts"])
    for kind, maxes, bins in ((capi.PMFT_XYZ, (1, 1, 1), (4, 0, 4)), (capi.PMFT_XYT, (-1, 1), (4, 4, 4)),
                              (capi.PMFT_R12, (-1,), (4, 4, 4)), (7, (1, 1, 1), (4, 4, 4))):
        with pytest.raises(ValueError):
            capi.DevicePMFT(ctx, kind, maxes, bins)


def test_pmft_query_and_histogram_in_one_call(ctx):
    """fgpu_pmft_accumulate / fgpu_pmftxy_accumulate: the ball query and the histogram without a NeighborList in between
    (the bonds are read from the search's bag) -- the committed outputs of the reference again, bit for bit, for all
    four classes, both engines' arithmetic, separate query points and self queries, two accumulated frames, the lattice
    whose every bond goes to the host (its list outgrows the first guess and the frame is repeated), and the tiny box
    that the warp-cooperative search does not take (fallback through a NeighborList)."""
    from freud_b200.box import Box
    from tests.golden.make_golden import PMFT3_EQUIV, pmft3_lattice, pmft3_quats, pmftxy_inputs

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "pmft3.npz"))
    gold_xy = np.load(os.path.join(GOLD, "pmftxy.npz"))
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 1500, 31), random_points(box, 400, 32)
    r = float(np.sqrt(2.0 ** 2 + 2.5 ** 2 + 3.0 ** 2))
    dp = capi.DevicePoints(ctx, box, pts)
    xyz = capi.DevicePMFT(ctx, capi.PMFT_XYZ, (2.0, 2.5, 3.0), (12, 10, 8))
    xyz.accumulate(dp, q, IMAGE, r, None, pmft3_quats(400, 6), PMFT3_EQUIV)
    assert np.array_equal(xyz.read(), gold["xyz_query_counts"])
    xyz.accumulate(dp, q, IMAGE, r, None, pmft3_quats(400, 6), PMFT3_EQUIV)
    assert np.array_equal(xyz.read(), 2 * gold["xyz_query_counts"])
    xyz.reset()
    xyz.accumulate(dp, None, IMAGE, r, None, pmft3_quats(1500, 7), PMFT3_EQUIV[:1], exclude_ii=True)
    assert np.array_equal(xyz.read(), gold["xyz_self_counts"])
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        dp = capi.DevicePoints(ctx, box, pts)
        r_xy = float(np.sqrt(3.0 ** 2 + 2.5 ** 2))
        xy = capi.DevicePMFTXY(ctx, 3.0, 2.5, 30, 24)
        xy.accumulate(dp, q, IMAGE, r_xy, th_q)
        assert np.array_equal(xy.read(), gold_xy[f"{name}_query_counts"])
        xy.reset()
        xy.accumulate(dp, None, IMAGE, r_xy, th_p, exclude_ii=True)
        assert np.array_equal(xy.read(), gold_xy[f"{name}_self_counts"])
        xyt = capi.DevicePMFT(ctx, capi.PMFT_XYT, (3.0, 2.5), (14, 12, 9))
        xyt.accumulate(dp, q, IMAGE, r_xy, th_p, th_q)
        assert np.array_equal(xyt.read(), gold[f"{name}_xyt_query_counts"])
        r12 = capi.DevicePMFT(ctx, capi.PMFT_R12, (4.0,), (10, 11, 12))
        r12.accumulate(dp, None, IMAGE, 4.0, th_p, th_p, exclude_ii=True)
        assert np.array_equal(r12.read(), gold[f"{name}_r12_self_counts"])
        # the LinkCell arithmetic gives the same bonds up to the last place of their vectors: same totals here
        r12w = capi.DevicePMFT(ctx, capi.PMFT_R12, (4.0,), (10, 11, 12))
        r12w.accumulate(dp, None, WRAP, 4.0, th_p, th_p, exclude_ii=True)
        nl = port.ball_nlist(port.WRAP, box, True, pts, pts, 4.0, 0.0, True)
        assert np.array_equal(r12w.read(), port.pmft3(port.PMFT_R12, box, 3000, nl, th_p, th_p, (4.0,), (10, 11, 12))[0])
    box, pts, th = pmft3_lattice()
    dp = capi.DevicePoints(ctx, box, pts)
    xyt = capi.DevicePMFT(ctx, capi.PMFT_XYT, (3.0, 3.0), (6, 6, 8))
    xyt.accumulate(dp, None, IMAGE, float(np.sqrt(18.0)), th, th, exclude_ii=True)
    assert np.array_equal(xyt.read(), gold["lattice_xyt_counts"]) and xyt.host_binned_bonds > 1000
    # 2 cells per axis: not the warp-cooperative search's case
    tiny = Box.square(7)
    tp = random_points(tiny, 60, 3)
    ta = np.linspace(0, 6, 60).astype(np.float32)
    tdp = capi.DevicePoints(ctx, tiny, tp)
    small = capi.DevicePMFT(ctx, capi.PMFT_XYT, (2.0, 2.0), (5, 5, 4))
    small.accumulate(tdp, None, IMAGE, 3.0, ta, ta, exclude_ii=True)
    nl = port.ball_nlist(port.IMAGE, tiny, True, tp, tp, 3.0, 0.0, True)
    assert np.array_equal(small.read(), port.pmft3(port.PMFT_XYT, tiny, 60, nl, ta, ta, (2.0, 2.0), (5, 5, 4))[0])


def test_bond_order_query_and_histogram_in_one_call(ctx):
    """fgpu_bondorder_accumulate: ball query and diagram without a NeighborList (bonds read from the search's bag) -- the
    oracle's counts over the same bonds, bit for bit, in all four modes, for separate query points and a self query, two
    accumulated frames; a perfect FCC lattice sends every bond to the host and the frame is repeated with a longer list."""
    from freud_b200 import data
    from freud_b200.box import Box
    from tests.golden.make_golden import pmft3_quats

    capi = _capi()
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 800, 41), random_points(box, 300, 42)
    o, qo = pmft3_quats(800, 1), pmft3_quats(300, 2)
    dp = capi.DevicePoints(ctx, box, pts)
    for flavour, pflav in ((IMAGE, port.IMAGE), (WRAP, port.WRAP)):
        nl_q = port.ball_nlist(pflav, box, False, pts, q, 3.0)
        nl_s = port.ball_nlist(pflav, box, False, pts, pts, 3.0, 0.0, True)
        for mode in ("bod", "lbod", "obcd", "oocd"):
            bo = capi.DeviceBondOrder(ctx, 12, 9, mode)
            bo.accumulate(dp, q, flavour, 3.0, o, qo)
            want = port.bond_order(mode, nl_q, o, qo, (12, 9))[0]
            assert np.array_equal(bo.read(), want), mode
            bo.accumulate(dp, q, flavour, 3.0, o, qo)
            assert np.array_equal(bo.read(), 2 * want), mode
            bo.reset()
            bo.accumulate(dp, None, flavour, 3.0, o, o, exclude_ii=True)
            assert np.array_equal(bo.read(), port.bond_order(mode, nl_s, o, o, (12, 9))[0]), mode
    fbox, fpts = data.UnitCell.fcc().generate_system(8)
    ident = np.tile(np.float32([1, 0, 0, 0]), (len(fpts), 1))
    fcc = capi.DeviceBondOrder(ctx, 8, 4)
    fcc.accumulate(capi.DevicePoints(ctx, fbox, fpts), None, IMAGE, 0.8, exclude_ii=True)  # the 12 nearest: a / sqrt 2
    nl = port.ball_nlist(port.IMAGE, fbox, False, fpts, fpts, 0.8, 0.0, True)
    assert len(nl.distances) == 12 * len(fpts)
    assert np.array_equal(fcc.read(), port.bond_order("bod", nl, ident, ident, (8, 4))[0])
    assert fcc.host_binned_bonds > len(nl.distances) // 16  # more than the first list held: the frame was repeated


def test_local_density_and_correlation_in_one_call(ctx):
    """fgpu_local_density_query / fgpu_corr_accumulate: the ball query and the sums without a NeighborList in between (the
    bonds are read from the search's bag).  Correlation bin counts are identical to the list route's and the oracle's;
    the float / double sums run in bag order and agree to rounding -- the bar every on-the-fly query is held to; rows
    without bonds stay 0; the tiny box falls back to the list route."""
    from freud_b200.box import Box
    from tests.golden.make_golden import correlation_inputs

    capi = _capi()
    for box, n, flavour in ((Box.cube(12), 3000, IMAGE), (Box(30, 26, 0, 0.35, 0, 0, is2D=True), 2500, WRAP),
                            (Box(14, 15, 16, 0.3, 0.2, 0.1), 2500, IMAGE)):
        pts, q = random_points(box, n, 7), random_points(box, 700, 8)
        v, qv = correlation_inputs(n, 700, 3)
        dp = capi.DevicePoints(ctx, box, pts)
        for qpts, qvals, excl in ((q, qv, False), (None, v, True)):
            nl_dev = dp.ball_query(qpts, flavour, 3.0, 0.0, excl)
            num_l, den_l = nl_dev.local_density(2.5, 1.0, is2d=box.is2D)
            num_f, den_f = dp.local_density(qpts, flavour, 3.0, 2.5, 1.0, exclude_ii=excl)
            assert np.allclose(num_f, num_l, rtol=2e-6, atol=1e-6) and np.allclose(den_f, den_l, rtol=2e-6, atol=1e-7)
            assert np.array_equal(num_f == 0, num_l == 0)
            listed = capi.DeviceCorrelation(ctx, 40, 3.0)
            listed.accumulate_nlist(nl_dev, v, qvals)
            fused = capi.DeviceCorrelation(ctx, 40, 3.0)
            fused.accumulate(dp, qpts, flavour, 3.0, v, qvals, exclude_ii=excl)
            fused.accumulate(dp, qpts, flavour, 3.0, v, qvals, exclude_ii=excl)  # a second frame accumulates
            (c_l, s_l), (c_f, s_f) = listed.read(), fused.read()
            assert np.array_equal(c_f, 2 * c_l)
            assert np.allclose(s_f, 2 * s_l, rtol=1e-12, atol=1e-9 * np.abs(s_l).max())
    sparse = Box.cube(40)
    far = random_points(sparse, 50, 1)  # hardly any bonds within 1.5: empty rows
    num, den = capi.DevicePoints(ctx, sparse, far).local_density(None, IMAGE, 1.5, 1.0, 1.0, exclude_ii=True)
    want = port.local_density(port.ball_nlist(port.IMAGE, sparse, False, far, far, 1.5, 0.0, True), 1.0, 1.0)
    assert np.allclose(num, want[0], rtol=2e-6) and np.allclose(den, want[1], rtol=2e-6) and (num == 0).sum() > 10
    tiny = Box.cube(5)
    tp = random_points(tiny, 80, 2)
    num, den = capi.DevicePoints(ctx, tiny, tp).local_density(None, IMAGE, 2.4, 2.0, 0.8, exclude_ii=True)
    want = port.local_density(port.ball_nlist(port.IMAGE, tiny, False, tp, tp, 2.4, 0.0, True), 2.0, 0.8)
    assert np.array_equal(bits(num), bits(want[0]))  # the list route (2 cells per axis): list order, bit for bit


def test_bond_order_over_a_neighbor_list(ctx):
    """fgpu_bondorder_* (BondOrder.cc:100-153) over device NeighborLists against the committed outputs of the reference:
    bin counts bit for bit in all four modes, on the FCC lattice whose bond directions sit on bin edges, accumulated over
    two calls, with a histogram too large for shared memory; constructor errors."""
    from freud_b200 import data
    from freud_b200.box import Box
    from tests.golden.make_golden import pmft3_quats

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "bond_order.npz"))
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 800, 41), random_points(box, 300, 42)
    o, qo = pmft3_quats(800, 1), pmft3_quats(300, 2)
    dp = capi.DevicePoints(ctx, box, pts)
    for mode in ("bod", "lbod", "obcd", "oocd"):
        bo = capi.DeviceBondOrder(ctx, 12, 9, mode)
        bo.accumulate_nlist(dp.knn_query(q, 8), o, qo)
        assert np.array_equal(bo.read(), gold[f"tri_{mode}_counts"]), mode
        bo.accumulate_nlist(dp.knn_query(q, 8), o, qo)
        assert np.array_equal(bo.read(), 2 * gold[f"tri_{mode}_counts"]), mode
        bo.reset()
        assert bo.read().sum() == 0
    plain = capi.DeviceBondOrder(ctx, 12, 9)
    plain.accumulate_nlist(dp.knn_query(q, 8))  # bod reads no orientations
    assert np.array_equal(plain.read(), gold["tri_bod_counts"])
    big = capi.DeviceBondOrder(ctx, 160, 80, "obcd")  # 51 KB of counters: global atomics
    big.accumulate_nlist(dp.knn_query(q, 8), o, qo)
    nl = port.knn_nlist(box, False, pts, q, 8)
    assert np.array_equal(big.read(), port.bond_order("obcd", nl, o, qo, (160, 80))[0])
    fbox, fpts = data.UnitCell.fcc().generate_system(4)
    fdp = capi.DevicePoints(ctx, fbox, fpts)
    for bins in ((8, 4), (7, 5)):
        fcc = capi.DeviceBondOrder(ctx, *bins)
        fcc.accumulate_nlist(fdp.knn_query(None, 12, exclude_ii=True))
        assert np.array_equal(fcc.read(), gold[f"fcc_{bins[0]}x{bins[1]}_counts"]) and fcc.host_binned_bonds > 0
    for n_theta, n_phi, mode in ((1, 4, "bod"), (4, 1, "bod"), (4, 4, "nope")):
        with pytest.raises(ValueError):
            capi.DeviceBondOrder(ctx, n_theta, n_phi, mode)


def test_context_trim_keeps_results_identical(ctx):
    """fgpu_ctx_trim drops the grow-only scratch and the spare NeighborList arrays; the next queries re-grow them and give
    the same lists; a list that is alive across the trim is untouched."""
    from freud_b200.box import Box

    capi = _capi()
    box = Box(14, 15, 16, 0.3, 0.2, 0.1)
    pts = random_points(box, 3000, 5)
    dp = capi.DevicePoints(ctx, box, pts)
    first = dp.ball_query(None, IMAGE, 3.0, 0.0, True).to_host()
    alive = dp.knn_query(None, 6, exclude_ii=True)
    ctx.trim()
    again = dp.ball_query(None, IMAGE, 3.0, 0.0, True).to_host()
    for key in ("neighbors", "distances", "weights", "vectors", "segments", "counts"):
        assert np.array_equal(first[key], again[key]), key
    want = port.knn_nlist(box, False, pts, pts, 6, exclude_ii=True)
    assert np.array_equal(alive.to_host()["neighbors"], want.neighbors)
    ctx.trim()
    ctx.trim()
    assert np.array_equal(dp.knn_query(None, 6, exclude_ii=True).to_host()["neighbors"], want.neighbors)


def test_correlation_function_over_a_neighbor_list(ctx):
    """fgpu_corr_* (CorrelationFunction.cc:26-95) over device NeighborLists against the committed outputs of the
    reference: bin counts identical, complex<double> sums to double rounding; accumulation over two calls."""
    from freud_b200.box import Box
    from tests.golden.make_golden import correlation_inputs

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "correlation_function.npz"))
    for name, box, n in (("cube", Box.cube(12), 3000), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True), 2500)):
        pts, q = random_points(box, n, 7), random_points(box, 700, 8)
        v, qv = correlation_inputs(n, 700, 3)
        dp = capi.DevicePoints(ctx, box, pts)
        cf = capi.DeviceCorrelation(ctx, 40, 3.0)
        cf.accumulate_nlist(dp.ball_query(q, IMAGE, 3.0, 0.0, False), v, qv)
        counts, sums = cf.read()
        assert np.array_equal(counts, gold[f"{name}_complex_counts"])
        np.testing.assert_allclose(sums / np.maximum(counts, 1), gold[f"{name}_complex_corr"], rtol=1e-11, atol=1e-12)
        cf.accumulate_nlist(dp.ball_query(q, IMAGE, 3.0, 0.0, False), v, qv)  # reset=False: a second frame adds
        counts2, sums2 = cf.read()
        assert np.array_equal(counts2, 2 * counts)
        np.testing.assert_allclose(sums2, 2 * sums, rtol=1e-11, atol=1e-11)
        cf.reset()
        cf.accumulate_nlist(dp.ball_query(None, IMAGE, 3.0, 0.0, True), v.real, v.real)
        counts, sums = cf.read()
        assert np.array_equal(counts, gold[f"{name}_real_counts"])
        np.testing.assert_allclose(sums / np.maximum(counts, 1), gold[f"{name}_real_corr"], rtol=1e-11, atol=1e-12)
    big = capi.DeviceCorrelation(ctx, 5000, 3.0)  # accumulators beyond shared memory: global atomics
    big.accumulate_nlist(dp.ball_query(None, IMAGE, 3.0, 0.0, True), v.real, v.real)
    nl = port.ball_nlist(port.IMAGE, box, box.is2D, pts, pts, 3.0, 0.0, True)
    want_corr, want_counts = port.correlation_function(nl, v.real, v.real, 5000, 3.0)
    counts, sums = big.read()
    assert np.array_equal(counts, want_counts)
    np.testing.assert_allclose(sums / np.maximum(counts, 1), want_corr, rtol=1e-11, atol=1e-12)


def test_local_density_over_a_neighbor_list(ctx):
    """fgpu_local_density (LocalDensity.cc:38-84) over device NeighborLists: the same bits as the oracle over the same
    list, in 3-D and 2-D, and the committed outputs of the reference."""
    from freud_b200.box import Box

    capi = _capi()
    gold = np.load(os.path.join(GOLD, "local_density.npz"))
    for name, box, n in (("cube", Box.cube(10), 3000), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True), 2500)):
        pts, q = random_points(box, n, 123), random_points(box, 500, 124)
        dp = capi.DevicePoints(ctx, box, pts)
        for r_max, diameter in ((3.0, 1.0), (2.0, 0.5)):
            key = f"{name}_{r_max:g}_{diameter:g}"
            nl = dp.ball_query(q, IMAGE, r_max + 0.5 * diameter, 0.0, False)
            num, den = nl.local_density(r_max, diameter, box.is2D)
            assert np.array_equal(bits(num), bits(gold[f"{key}_nlist_num"])), key
            assert np.array_equal(bits(den), bits(gold[f"{key}_nlist_density"])), key
            want = port.local_density(port.ball_nlist(port.WRAP, box, box.is2D, pts, pts, r_max + 0.5 * diameter, 0.0, True),
                                      r_max, diameter, box.is2D)
            got = dp.ball_query(None, WRAP, r_max + 0.5 * diameter, 0.0, True).local_density(r_max, diameter, box.is2D)
            assert np.array_equal(bits(got[0]), bits(want[0])) and np.array_equal(bits(got[1]), bits(want[1])), key
    with pytest.raises(ValueError):
        nl.local_density(-1.0, 1.0)
    with pytest.raises(ValueError):
        nl.local_density(1.0, -1.0)


@pytest.mark.parametrize("name", list(BOXES))
def test_ghost_flavour_ball_and_rdf(ctx, name):
    """CellQuery's arithmetic r = (p_j + shift) - q (CellQuery.cc:107, CellIterator.h:167; E5 in oracle/port.c):
    NeighborList in both orders and fused RDF, self query and a separate query set."""
    capi = _capi()
    box, n, r = BOXES[name]
    r = min(r, 0.49 * float(min(box.Lx, box.Ly)))
    pts = random_points(box, n, seed=51)
    dp = capi.DevicePoints(ctx, box, pts)
    for q, excl, r_min in ((None, True, 0.0), (random_points(box, 400, seed=52), False, 0.5)):
        qq = pts if q is None else q
        for sbd in (False, True):
            got = dp.ball_query(q, GHOST, r, r_min, excl, sbd).to_host()
            want = port.ball_nlist(port.GHOST, box, box.is2D, pts, qq, r, r_min, excl, sbd)
            assert_nlist_equal(got, want, f"{name} ghost self={q is None} sbd={sbd}")
        rdf = capi.DeviceRDF(ctx, 50, r)
        rdf.accumulate(dp, q, GHOST, r, r_min, excl)
        assert np.array_equal(rdf.read(), port.rdf_accumulate(port.GHOST, box, box.is2D, pts, qq, 50, r, 0.0, excl,
                                                              query_r_min=r_min))
    with pytest.raises(RuntimeError, match="CellQuery only supports"):
        dp.knn_query(None, 4, flavour=GHOST)


@pytest.mark.parametrize("name", list(BOXES))
def test_knn_wrap_flavour(ctx, name):
    """LinkCell's nearest-neighbour iterator (LinkCell.cc:575-679): the k smallest WRAPPED distances."""
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=33)
    dp = _capi().DevicePoints(ctx, box, pts)
    for q, excl in ((None, True), (random_points(box, 300, seed=34), False)):
        qq = pts if q is None else q
        for k, sbd, kw in ((12, False, {}), (5, True, {}), (7, False, dict(r_max=2.2, r_min=0.7))):
            got = dp.knn_query(q, k, exclude_ii=excl, sort_by_distance=sbd, flavour=WRAP, **kw).to_host()
            want = port.knn_nlist(box, box.is2D, pts, qq, k, kw.get("r_max", np.inf), kw.get("r_min", 0.0), excl, sbd,
                                  flavour=port.WRAP)
            assert_nlist_equal(got, want, f"{name} wrap knn k={k} self={q is None}")


@pytest.mark.parametrize("k", [4, 12])
def test_knn_short_rows_are_searched_again(ctx, k):
    """Uniform random points: the first window (2(k+1) expected points) leaves a few rows with fewer than k hits;
    those rows -- only those -- are searched again with a wider window into a second bag (knn2.cu).  Self query and
    a separate query set, both orders."""
    box = __import__("freud_b200.box", fromlist=["Box"]).Box(26, 24, 22, 0.2, -0.1, 0.15)
    pts = random_points(box, 5000, seed=77)
    dp = _capi().DevicePoints(ctx, box, pts)
    for q, excl in ((None, True), (random_points(box, 1500, seed=78), False)):
        for sbd in (False, True):
            got = dp.knn_query(q, k, exclude_ii=excl, sort_by_distance=sbd).to_host()
            qq = pts if q is None else q
            want = port.knn_nlist(box, False, pts, qq, k, exclude_ii=excl, sort_by_distance=sbd)
            assert_nlist_equal(got, want, f"knn k={k} self={q is None} sbd={sbd}")
            assert (got["counts"] == k).all()


@pytest.mark.parametrize("name", ["cubic", "tri2", "sq2d", "tilt2d"])
@pytest.mark.parametrize("flavour", [WRAP, IMAGE])
def test_rdf_home_tiles_sharded(ctx, name, flavour):
    """fgpu_points_set_shard: the home tiles of a self-query RDF dealt to S ranks, each with a cell list built for
    its slab only.  One process plays the ranks in turn; the summed counts are the single-GPU counts bit for bit."""
    capi = _capi()
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=91)
    want = port.rdf_accumulate(flavour, box, box.is2D, pts, pts, 64, r, 0.0, True)
    for shards in (2, 3, 5):
        rdf = capi.DeviceRDF(ctx, 64, r)
        for s in range(shards):
            dp = capi.DevicePoints(ctx, box, pts)
            dp.set_shard(s, shards)
            rdf.accumulate(dp, None, flavour, r, 0.0, True)
        assert np.array_equal(rdf.read(), want), f"{name} flavour={flavour} shards={shards}"
    # sharded points serve that call only
    dp = capi.DevicePoints(ctx, box, pts)
    dp.set_shard(1, 2)
    with pytest.raises(RuntimeError):
        dp.ball_query(None, flavour, r, 0.0, True)
    with pytest.raises(RuntimeError):
        capi.DeviceRDF(ctx, 64, r).accumulate(dp, pts[:10], flavour, r, 0.0, False)
    dp.set_shard(0, 1)
    assert dp.ball_query(None, flavour, r, 0.0, True).num_bonds == int(want.sum())
    # un-wrapped input: the sharded path has no general-kernel fallback and says so when the counts are read
    dq = capi.DevicePoints(ctx, box, random_points(box, n, seed=92, spill=0.05))
    dq.set_shard(0, 2)
    bad = capi.DeviceRDF(ctx, 64, r)
    bad.accumulate(dq, None, flavour, r, 0.0, True)
    with pytest.raises(RuntimeError):
        bad.read()


def test_knn_r_max_r_min(ctx):
    box, n, r = BOXES["cubic"]
    pts = random_points(box, n, seed=32)
    dp = _capi().DevicePoints(ctx, box, pts)
    got = dp.knn_query(pts[:300], 12, r_max=1.6, r_min=0.7, exclude_ii=True).to_host()
    want = ref.Query("aabb", box, pts).nlist(pts[:300], num_neighbors=12, r_max=1.6, r_min=0.7, exclude_ii=True) \
        if ref.available() else port.knn_nlist(box, False, pts, pts[:300], 12, 1.6, 0.7, True)
    assert_nlist_equal(got, want, "knn r_max r_min")


def test_steinhardt_fcc_known_answer(ctx):
    """PERFECT_FCC_Q6 = 0.57452416 (reference tests/test_order_steinhardt.py:17, :101-166), k = 12."""
    from freud_b200 import data

    capi = _capi()
    box, pts = data.make_fcc_system(4)
    dp = capi.DevicePoints(ctx, box, pts)
    nl = dp.knn_query(None, 12, exclude_ii=True)
    out = dp.steinhardt(nl, [6])
    assert np.allclose(out["ql"], 0.57452416, atol=1e-5)
    assert abs(out["order"][0] - 0.57452416) < 1e-5
    # ball query r_max = 1.5 * nn distance gives the same shell
    nl = dp.ball_query(None, IMAGE, 0.8, 0.0, True)
    assert np.allclose(dp.steinhardt(nl, [6])["ql"], 0.57452416, atol=1e-5)


@pytest.mark.parametrize("ls", [[6], [4], [4, 6], [2, 8, 12], [3], [0, 1, 5], [20]])
def test_steinhardt_vs_oracle(ctx, ls):
    from freud_b200 import data

    capi = _capi()
    box, pts = data.make_fcc_system(5, scale=1.3, sigma_noise=0.08, seed=3)
    dp = capi.DevicePoints(ctx, box, pts)
    nl = dp.knn_query(None, 12, exclude_ii=True)
    got = dp.steinhardt(nl, ls)
    h = nl.to_host()

    class _NL:
        neighbors, distances, weights, segments, counts = (h["neighbors"], h["distances"], h["weights"],
                                                           h["segments"], h["counts"])

    want = port.steinhardt(box, False, pts, _NL, ls)
    assert np.allclose(got["ql"], want["ql"], rtol=1e-5, atol=1e-6)
    for a, b in zip(got["qlm"], want["qlm"]):
        assert np.allclose(a, b, atol=1e-5)
    assert np.allclose(got["order"], want["order"], rtol=1e-4, atol=1e-6)


def test_steinhardt_no_neighbours_is_nan(ctx):
    """tests/test_order_steinhardt.py:361-369."""
    capi = _capi()
    from freud_b200.box import Box

    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [0.5, 0, 0], [4, 4, 4]], np.float32)
    dp = capi.DevicePoints(ctx, box, pts)
    nl = dp.ball_query(None, IMAGE, 1.0, 0.0, True)
    ql = dp.steinhardt(nl, [6])["ql"][:, 0]
    assert np.isfinite(ql[0]) and np.isfinite(ql[1]) and np.isnan(ql[2])


@pytest.mark.parametrize("flavour", [WRAP, IMAGE])
def test_medium_system_properties(ctx, flavour):
    """N = 200k: counts vs the port, reciprocity of the pair set, RDF == histogram of the list."""
    from freud_b200 import data

    capi = _capi()
    box, pts = data.make_random_system(135.72, 200000, seed=5)
    dp = capi.DevicePoints(ctx, box, pts)
    nl = dp.ball_query(None, flavour, 3.0, 0.0, True)
    got = nl.to_host()
    want = port.ball_nlist(flavour, box, False, pts, pts, 3.0, 0.0, True)
    assert_nlist_equal(got, want, "200k")
    i, j = got["neighbors"][:, 0].astype(np.int64), got["neighbors"][:, 1].astype(np.int64)
    fwd = np.sort(i * len(pts) + j)
    bwd = np.sort(j * len(pts) + i)
    # the pair set is symmetric up to bonds sitting on the r_max rounding edge: neither wrap(p_j - p_i) nor
    # p_j - (p_i + image) is bitwise antisymmetric in float32 (the reference shows the same asymmetry)
    assert len(np.setxor1d(fwd, bwd)) <= 1e-4 * len(fwd)
    rdf = capi.DeviceRDF(ctx, 100, 3.0)
    rdf.accumulate(dp, None, flavour, 3.0, 0.0, True)
    assert np.array_equal(rdf.read(), port.rdf_accumulate_distances(got["distances"], 100, 3.0))


@pytest.mark.parametrize("flavour", [IMAGE, WRAP])
def test_steinhardt_knn_fused_matches_list_route(ctx, flavour):
    """fgpu_steinhardt_knn: the k nearest of every row picked out of the window search's bag and their Y_lm sums in
    one kernel must give what the NeighborList route gives (q_l to summation order: the list is sorted by j, the bag is
    not) and what the oracle gives (1e-5); ties, short rows (k > neighbours in r_max) and k = n - 1 included."""
    from freud_b200 import data

    capi = _capi()
    box, pts = data.make_fcc_system(7, sigma_noise=0.05, seed=4)
    dp = capi.DevicePoints(ctx, box, pts)
    for l, k in ((6, 12), (4, 6), (12, 16)):
        fused = dp.steinhardt_knn(k, [l], exclude_ii=True, flavour=flavour, want_qlm=True)
        nl = dp.knn_query(None, k, exclude_ii=True, flavour=flavour)
        listed = dp.steinhardt(nl, [l])
        assert np.allclose(fused["ql"], listed["ql"], rtol=2e-6, atol=1e-7), (l, k)
        assert np.allclose(fused["qlm"][0], listed["qlm"][0], atol=2e-6)
        assert np.allclose(fused["order"], listed["order"], rtol=1e-5)
    pnl = port.knn_nlist(box, False, pts, pts, 12, exclude_ii=True, flavour=port.IMAGE if flavour == IMAGE else port.WRAP)
    want = port.steinhardt(box, False, pts, pnl, [6])["ql"]
    assert np.allclose(dp.steinhardt_knn(12, [6], exclude_ii=True, flavour=flavour)["ql"], want, rtol=1e-5, atol=1e-6)
    # a perfect lattice: twelve exactly tied nearest neighbours, then a gap -- k = 8 cuts through the tie group
    lbox, lpts = data.make_fcc_system(5)
    dl = capi.DevicePoints(ctx, lbox, lpts)
    f8 = dl.steinhardt_knn(8, [6], exclude_ii=True, flavour=flavour)["ql"]
    l8 = dl.steinhardt(dl.knn_query(None, 8, exclude_ii=True, flavour=flavour), [6])["ql"]
    assert np.allclose(f8, l8, rtol=2e-6, atol=1e-7)
    # r_max so small that most rows hold fewer than k neighbours; unsupported l / k fall back to the list inside the call
    short = dp.steinhardt_knn(12, [6], r_max=0.72, exclude_ii=True, flavour=flavour)["ql"]
    short_l = dp.steinhardt(dp.knn_query(None, 12, r_max=0.72, exclude_ii=True, flavour=flavour), [6])["ql"]
    assert np.array_equal(np.isnan(short), np.isnan(short_l))
    assert np.allclose(short[~np.isnan(short)], short_l[~np.isnan(short_l)], rtol=2e-6, atol=1e-7)
    odd = dp.steinhardt_knn(20, [5, 7], exclude_ii=True, flavour=flavour)["ql"]
    odd_l = dp.steinhardt(dp.knn_query(None, 20, exclude_ii=True, flavour=flavour), [5, 7])["ql"]
    assert np.allclose(odd, odd_l, rtol=2e-6, atol=1e-7)


def test_kernel_timeline_of_a_frame():
    """fgpu_ctx_profile + fgpu_ctx_kernel_time / fgpu_ctx_kernel_timeline (tracing): every launch
    of a NeighborList frame is named and bracketed on the device clock, in stream order, and the summed durations of
    the two views agree."""
    from freud_b200.box import Box

    capi = _capi()
    c = capi.Context(0)
    try:
        box = Box.cube(40.0)
        pts = random_points(box, 20000, 5)
        dp = capi.DevicePoints(c, box, pts)
        dp.ball_query(None, WRAP, 3.0, 0.0, True)  # warm: allocations, kernel attributes
        c.profile(True)
        c.kernel_time("", reset=True)
        dp.build_cells(3.0)
        nl = dp.ball_query(None, WRAP, 3.0, 0.0, True)
        tl = c.kernel_timeline()
        total_ms, launches = c.kernel_time("", reset=True)
        names = [t[0] for t in tl]
        assert launches == len(tl) >= 6
        for must in ("cell_assign", "scan", "cell_scatter", "search_nl", "emit"):
            assert must in names, names
        assert names.index("cell_assign") < names.index("search_nl") < names.index("emit")
        assert tl[0][1] == 0.0 and all(e >= b for _, b, e in tl)
        assert all(tl[k + 1][1] >= tl[k][2] - 1e-3 for k in range(len(tl) - 1))  # one stream: no overlap
        assert abs(sum(e - b for _, b, e in tl) - total_ms * 1e3) < 1.0  # microseconds
        assert nl.num_bonds > 0
        c.profile(False)
    finally:
        c.close()
