"""Pins the oracle (CPU, no GPU needed).

1. ``oracle/port.c`` (plain-C restatement) against the golden vectors produced by the reference itself
   (tests/golden/*.npz, generator committed beside them) -- bit for bit.
2. the same port against ``oracle/_ref`` (the unmodified reference compiled here) on more boxes, when that
   library exists (it does wherever /root/reference was available at build time).
3. the known answers the reference's own tests hold for this path (SURVEY.md section 8c list).
"""
import os

import numpy as np
import pytest

from freud_b200 import data
from freud_b200.box import Box
from oracle import port, ref
from tests.golden.make_golden import CASES
from tests.util import BOXES, bits, random_points

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def same_nl(got, gold, prefix):
    assert np.array_equal(got.neighbors, gold[f"{prefix}_neighbors"])
    assert np.array_equal(bits(got.distances), bits(gold[f"{prefix}_distances"]))
    assert np.array_equal(bits(got.vectors), bits(gold[f"{prefix}_vectors"]))
    assert np.array_equal(got.segments, gold[f"{prefix}_segments"])
    assert np.array_equal(got.counts, gold[f"{prefix}_counts"])


@pytest.mark.parametrize("name", list(CASES))
def test_port_matches_golden_neighbor_lists(name):
    box, n, nq, r_max, r_min, excl, seed = CASES[name]
    gold = np.load(os.path.join(GOLD, f"nl_{name}.npz"))
    pts = random_points(box, n, seed)
    q = pts if nq == 0 else random_points(box, nq, seed + 1000)
    same_nl(port.ball_nlist(port.WRAP, box, box.is2D, pts, q, r_max, r_min, excl), gold, "wrap")
    same_nl(port.ball_nlist(port.IMAGE, box, box.is2D, pts, q, r_max, r_min, excl), gold, "image")
    same_nl(port.ball_nlist(port.IMAGE, box, box.is2D, pts, q, r_max, r_min, excl, True), gold, "image_bydist")
    same_nl(port.knn_nlist(box, box.is2D, pts, q, 6, exclude_ii=excl), gold, "knn6")
    for flavour, tag in ((port.WRAP, "wrap"), (port.IMAGE, "image")):
        counts = port.rdf_accumulate(flavour, box, box.is2D, pts, q, 40, r_max, r_min, excl)
        assert np.array_equal(counts, gold[f"rdf_{tag}_bin_counts"])
        red = port.rdf_reduce(counts, r_max, r_min, box, box.is2D, n, len(q))
        for key in ("rdf", "n_r", "bin_edges", "bin_centers"):
            assert np.array_equal(bits(red[key]), bits(gold[f"rdf_{tag}_{key}"])), key


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref is built only where /root/reference exists")
@pytest.mark.parametrize("name", list(BOXES))
def test_port_knn_wrap_matches_the_reference_linkcell(name):
    """fport_knn_nlist_wrap against the reference's own LinkCellQueryIterator (LinkCell.cc:575-679), bit for bit."""
    box, n, _ = BOXES[name]
    pts, q = random_points(box, min(n, 800), 5), random_points(box, 150, 6)
    width = min(2.0, 0.4 * float(min(box.Lx, box.Ly)))
    for k, sbd, kw in ((6, False, {}), (12, True, {}), (5, False, dict(r_max=2.0, r_min=0.6))):
        want = ref.Query("linkcell", box, pts, is2d=box.is2D, cell_width=width).nlist(q, num_neighbors=k,
                                                                                      sort_by_distance=sbd, **kw)
        got = port.knn_nlist(box, box.is2D, pts, q, k, kw.get("r_max", np.inf), kw.get("r_min", 0.0), False, sbd,
                             flavour=port.WRAP)
        assert np.array_equal(got.neighbors, want.neighbors), (name, k)
        assert np.array_equal(bits(got.distances), bits(want.distances)), (name, k)
        assert np.array_equal(bits(got.vectors), bits(want.vectors)), (name, k)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref is built only where /root/reference exists")
@pytest.mark.parametrize("name", list(BOXES))
def test_port_ghost_flavour_matches_the_reference_cellquery(name):
    """E5: the port's GHOST flavour against the reference's own CellQuery (CellQuery.cc, CellIterator.h), bit for bit."""
    box, n, r = BOXES[name]
    r = min(r, 0.49 * float(min(box.Lx, box.Ly)))
    pts, q = random_points(box, min(n, 1500), 5), random_points(box, 300, 6)
    for qq, excl, r_min, sbd in ((q, False, 0.4, False), (pts, True, 0.0, True)):
        want = ref.Query("cell", box, pts, is2d=box.is2D).nlist(qq, r_max=r, r_min=r_min, exclude_ii=excl,
                                                                sort_by_distance=sbd)
        got = port.ball_nlist(port.GHOST, box, box.is2D, pts, qq, r, r_min, excl, sbd)
        assert np.array_equal(got.neighbors, want.neighbors), name
        assert np.array_equal(bits(got.distances), bits(want.distances)), name
        assert np.array_equal(bits(got.vectors), bits(want.vectors)), name
        assert np.array_equal(got.segments, want.segments) and np.array_equal(got.counts, want.counts)


def test_port_matches_golden_rdf_config0():
    """BASELINE.json configs[0]: RDF bins=100 r_max=5 on make_random_system(50, 10000), one and two frames."""
    gold = np.load(os.path.join(GOLD, "rdf_config0.npz"))
    box, pts = data.make_random_system(50, 10000, seed=0)
    counts = port.rdf_accumulate(port.IMAGE, box, False, pts, pts, 100, 5.0, 0.0, True)
    assert np.array_equal(counts, gold["bin_counts"])
    red = port.rdf_reduce(counts, 5.0, 0.0, box, False, 10000, 10000)
    assert np.array_equal(bits(red["rdf"]), bits(gold["rdf"])) and np.array_equal(bits(red["n_r"]), bits(gold["n_r"]))
    b1, p1 = data.make_random_system(50, 10000, seed=1)
    port.rdf_accumulate(port.IMAGE, b1, False, p1, p1, 100, 5.0, 0.0, True, counts=counts)
    assert np.array_equal(counts, gold["two_frames_bin_counts"])
    red = port.rdf_reduce(counts, 5.0, 0.0, box, False, 10000, 10000, frames=2)
    assert np.array_equal(bits(red["rdf"]), bits(gold["two_frames_rdf"]))
    # statistical sanity the reference tests use (tests/test_density_rdf.py:94-127): g(r) -> 1
    assert abs(float(np.mean(red["rdf"][50:])) - 1.0) < 0.05


def test_port_matches_golden_steinhardt():
    gold = np.load(os.path.join(GOLD, "steinhardt_fcc.npz"))
    box, pts = data.make_fcc_system(4, scale=1.2, sigma_noise=0.06, seed=7)
    nl = port.knn_nlist(box, False, pts, pts, 12, exclude_ii=True, sort_by_distance=True)
    for ls in ([6], [4, 6], [2, 8], [12]):
        tag = "_".join(str(l) for l in ls)
        out = port.steinhardt(box, False, pts, nl, ls)
        assert np.allclose(out["ql"], gold[f"knn12_ql_{tag}"], rtol=1e-6, atol=1e-7)
        assert np.allclose(out["order"], gold[f"knn12_order_{tag}"], rtol=1e-5)
        for l, qlm in zip(ls, out["qlm"]):
            assert np.allclose(qlm, gold[f"knn12_qlm_{tag}_l{l}"], atol=1e-6)
    nlb = port.ball_nlist(port.IMAGE, box, False, pts, pts, 1.05, 0.0, True)
    assert np.allclose(port.steinhardt(box, False, pts, nlb, [6])["ql"], gold["ball_ql_6"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("tag", ["ave", "wl", "wln", "ave_wl", "ave_wln"])
def test_port_matches_golden_steinhardt_options(tag):
    """computeAve / aggregatewl / normalizeSystem restated in oracle/port.c against outputs of the reference itself
    (tests/golden/steinhardt_options.npz): per-particle values bit for bit, the system order to summation order."""
    from tests.golden.make_golden import STEINHARDT_OPTIONS

    gold = np.load(os.path.join(GOLD, "steinhardt_options.npz"))
    box, pts = data.make_fcc_system(4, scale=1.2, sigma_noise=0.06, seed=7)
    nl = port.knn_nlist(box, False, pts, pts, 12, exclude_ii=True, sort_by_distance=True)
    for ls in ([6], [4, 6], [3, 10]):
        key = tag + "_" + "_".join(str(l) for l in ls)
        out = port.steinhardt(box, False, pts, nl, ls, **STEINHARDT_OPTIONS[tag])
        assert np.array_equal(bits(out["particle_order"]), bits(gold[f"{key}_particle_order"])), key
        assert np.array_equal(bits(out["ql"]), bits(gold[f"{key}_ql"])), key
        np.testing.assert_allclose(out["order"], gold[f"{key}_order"], rtol=1e-5, atol=1e-8)


def test_port_local_density_matches_golden():
    """LocalDensity::compute restated in oracle/port.c against outputs of the reference (tests/golden/local_density.npz):
    bit for bit when the bonds are summed in list order, to summation order against the on-the-fly query."""
    gold = np.load(os.path.join(GOLD, "local_density.npz"))
    for name, box, n in (("cube", Box.cube(10), 3000), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True), 2500)):
        pts, q = random_points(box, n, 123), random_points(box, 500, 124)
        for r_max, diameter in ((3.0, 1.0), (2.0, 0.5)):
            key = f"{name}_{r_max:g}_{diameter:g}"
            nl = port.ball_nlist(port.IMAGE, box, box.is2D, pts, q, r_max + 0.5 * diameter, 0.0, False)
            num, den = port.local_density(nl, r_max, diameter, box.is2D)
            assert np.array_equal(bits(num), bits(gold[f"{key}_nlist_num"])), key
            assert np.array_equal(bits(den), bits(gold[f"{key}_nlist_density"])), key
            np.testing.assert_allclose(num, gold[f"{key}_query_num"], rtol=1e-5)
            np.testing.assert_allclose(den, gold[f"{key}_query_density"], rtol=1e-5)


def test_port_correlation_function_matches_golden():
    """CorrelationFunction restated in oracle/port.c against outputs of the reference
    (tests/golden/correlation_function.npz): identical bin counts, sums to double rounding."""
    from tests.golden.make_golden import correlation_inputs

    gold = np.load(os.path.join(GOLD, "correlation_function.npz"))
    for name, box, n in (("cube", Box.cube(12), 3000), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True), 2500)):
        pts, q = random_points(box, n, 7), random_points(box, 700, 8)
        v, qv = correlation_inputs(n, 700, 3)
        nl = port.ball_nlist(port.IMAGE, box, box.is2D, pts, q, 3.0, 0.0, False)
        corr, counts = port.correlation_function(nl, v, qv, 40, 3.0)
        assert np.array_equal(counts, gold[f"{name}_complex_counts"])
        np.testing.assert_allclose(corr, gold[f"{name}_complex_corr"], rtol=1e-12, atol=1e-13)
        nl = port.ball_nlist(port.IMAGE, box, box.is2D, pts, pts, 3.0, 0.0, True)
        corr, counts = port.correlation_function(nl, v.real, v.real, 40, 3.0)
        assert np.array_equal(counts, gold[f"{name}_real_counts"])
        np.testing.assert_allclose(corr, gold[f"{name}_real_corr"], rtol=1e-12, atol=1e-13)


def test_port_pmftxy_matches_golden():
    """PMFTXY restated in oracle/port.c against outputs of the reference (tests/golden/pmftxy.npz): bin counts and PCF
    bit for bit (cosf / sinf are the same libm's on both sides)."""
    from tests.golden.make_golden import pmftxy_inputs

    gold = np.load(os.path.join(GOLD, "pmftxy.npz"))
    r = float(np.sqrt(3.0 ** 2 + 2.5 ** 2))
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        nl = port.ball_nlist(port.IMAGE, box, True, pts, q, r, 0.0, False)
        counts, pcf = port.pmftxy(box, 3000, nl, th_q, 3.0, 2.5, 30, 24)
        assert np.array_equal(counts, gold[f"{name}_query_counts"])
        assert np.array_equal(bits(pcf), bits(gold[f"{name}_query_pcf"]))
        nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, r, 0.0, True)
        counts, pcf = port.pmftxy(box, 3000, nl, th_p, 3.0, 2.5, 30, 24)
        assert np.array_equal(counts, gold[f"{name}_self_counts"])
        assert np.array_equal(bits(pcf), bits(gold[f"{name}_self_pcf"]))


def test_port_pmft3_matches_golden():
    """PMFTXYZ / PMFTXYT / PMFTR12 restated in oracle/port.c against outputs of the reference (tests/golden/pmft3.npz):
    bin counts and PCF bit for bit (same libm), incl. the lattice whose bond angles all sit on bin edges."""
    from tests.golden.make_golden import PMFT3_EQUIV, pmft3_lattice, pmft3_quats, pmftxy_inputs

    gold = np.load(os.path.join(GOLD, "pmft3.npz"))

    def same(got, tag):
        assert np.array_equal(got[0], gold[f"{tag}_counts"]), tag
        assert np.array_equal(bits(got[1]), bits(gold[f"{tag}_pcf"])), tag

    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 1500, 31), random_points(box, 400, 32)
    mx, bins = (2.0, 2.5, 3.0), (12, 10, 8)
    r = float(np.sqrt(sum(m * m for m in mx)))
    nl = port.ball_nlist(port.IMAGE, box, False, pts, q, r)
    same(port.pmft3(port.PMFT_XYZ, box, len(pts), nl, None, pmft3_quats(400, 6), mx, bins, equiv=PMFT3_EQUIV), "xyz_query")
    nl = port.ball_nlist(port.IMAGE, box, False, pts, pts, r, exclude_ii=True)
    same(port.pmft3(port.PMFT_XYZ, box, len(pts), nl, None, pmft3_quats(1500, 7), mx, bins, equiv=PMFT3_EQUIV[:1]),
         "xyz_self")
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        r_xyt = float(np.sqrt(3.0 ** 2 + 2.5 ** 2))
        nl = port.ball_nlist(port.IMAGE, box, True, pts, q, r_xyt)
        same(port.pmft3(port.PMFT_XYT, box, len(pts), nl, th_p, th_q, (3.0, 2.5), (14, 12, 9)), f"{name}_xyt_query")
        nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, r_xyt, exclude_ii=True)
        same(port.pmft3(port.PMFT_XYT, box, len(pts), nl, th_p, th_p, (3.0, 2.5), (14, 12, 9)), f"{name}_xyt_self")
        nl = port.ball_nlist(port.IMAGE, box, True, pts, q, 4.0)
        same(port.pmft3(port.PMFT_R12, box, len(pts), nl, th_p, th_q, (4.0,), (10, 11, 12)), f"{name}_r12_query")
        nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, 4.0, exclude_ii=True)
        same(port.pmft3(port.PMFT_R12, box, len(pts), nl, th_p, th_p, (4.0,), (10, 11, 12)), f"{name}_r12_self")
    box, pts, th = pmft3_lattice()
    nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, float(np.sqrt(18.0)), exclude_ii=True)
    same(port.pmft3(port.PMFT_XYT, box, len(pts), nl, th, th, (3.0, 3.0), (6, 6, 8)), "lattice_xyt")
    nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, 3.0, exclude_ii=True)
    same(port.pmft3(port.PMFT_R12, box, len(pts), nl, th, th, (3.0,), (6, 8, 8)), "lattice_r12")


def test_port_bond_order_matches_golden():
    """BondOrder restated in oracle/port.c against outputs of the reference (tests/golden/bond_order.npz): bin counts and
    the diagram bit for bit (same libm), all four modes and the FCC lattice whose bond directions sit on bin edges."""
    from tests.golden.make_golden import pmft3_quats

    gold = np.load(os.path.join(GOLD, "bond_order.npz"))
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 800, 41), random_points(box, 300, 42)
    o, qo = pmft3_quats(800, 1), pmft3_quats(300, 2)
    nl = port.knn_nlist(box, False, pts, q, 8)
    for mode in ("bod", "lbod", "obcd", "oocd"):
        counts, bo = port.bond_order(mode, nl, o, qo, (12, 9))
        assert np.array_equal(counts, gold[f"tri_{mode}_counts"]), mode
        assert np.array_equal(bits(bo), bits(gold[f"tri_{mode}_bo"])), mode
    box, pts = data.UnitCell.fcc().generate_system(4)
    ident = np.tile(np.float32([1, 0, 0, 0]), (len(pts), 1))
    nl = port.knn_nlist(box, False, pts, pts, 12, exclude_ii=True)
    for bins in ((8, 4), (7, 5)):
        counts, bo = port.bond_order("bod", nl, ident, ident, bins)
        assert np.array_equal(counts, gold[f"fcc_{bins[0]}x{bins[1]}_counts"])
        assert np.array_equal(bits(bo), bits(gold[f"fcc_{bins[0]}x{bins[1]}_bo"]))


def test_port_wigner3j_known_values():
    """(0 0 0; 0 0 0) = 1; (1 1 1; m1 m2 m3) = +-1/sqrt(6) or 0 in the table order of Wigner3j.cc:43-55; and, where
    the reference is present, every tabulated l <= 20 as float."""
    assert np.array_equal(port.wigner3j(0), np.float32([1.0]))
    s = np.float32(1.0 / np.sqrt(6.0))
    assert np.array_equal(port.wigner3j(1), np.float32([s, -s, -s, 0.0, s, s, -s]))
    path = "/root/reference/freud/order/Wigner3j.cc"
    if os.path.exists(path):
        import re

        src = open(path).read()
        for l in range(21):
            body = re.search(r"case %d: \{\s*return \{(.*?)\};" % l, src, re.S).group(1)
            table = np.array([float(x) for x in body.replace("\n", " ").split(",") if x.strip()])
            assert np.array_equal(table.astype(np.float32), port.wigner3j(l)), l


# ---- known answers held by the reference's own tests ---------------------------------------------------
def test_known_answer_perfect_fcc_q6():
    """PERFECT_FCC_Q6 = 0.57452416, tests/test_order_steinhardt.py:17, :101-166 (k = 12 and ball)."""
    box, pts = data.make_fcc_system(4)
    nl = port.knn_nlist(box, False, pts, pts, 12, exclude_ii=True)
    out = port.steinhardt(box, False, pts, nl, [6])
    assert np.allclose(out["ql"], 0.57452416, atol=1e-5) and abs(out["order"][0] - 0.57452416) < 1e-5
    box, pts = data.make_fcc_system(4, scale=2.0)
    nl = port.ball_nlist(port.IMAGE, box, False, pts, pts, 1.5, 0.0, True)
    assert np.allclose(port.steinhardt(box, False, pts, nl, [6])["ql"], 0.57452416, atol=1e-5)


def test_known_answer_axis_aligned_ql():
    """q_l of one bond along z: Y_lm vanishes for m != 0, so q_l = 1 for every l; tests/test_order_steinhardt.py:79-99."""
    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [0, 0, 1]], np.float32)
    nl = port.ball_nlist(port.IMAGE, box, False, pts, pts, 1.5, 0.0, True)
    for l in range(0, 20):
        ql = port.steinhardt(box, False, pts, nl, [l])["ql"][:, 0]
        assert np.allclose(ql, 1.0, atol=1e-5), l


def test_known_answer_box_wrap_and_round_trip():
    """tests/test_box_box.py:118-121 (wrap in a tilted box), :316-353 (fractional/absolute round trips)."""
    for impl in (port.box_apply, lambda b, d, op, v: getattr(Box.from_box(b), {"wrap": "wrap", "fractional":
                 "make_fractional", "absolute": "make_absolute"}[op])(v)):
        assert np.array_equal(impl(Box(2, 2, 2, 1, 0, 0), False, "wrap", [[10, -5, -5]]), [[-2, -1, -1]])
    box = Box(2, 2, 2, 1, 0, 0)
    assert np.allclose(port.box_apply(box, False, "absolute", [[0.5, 0.5, 0.5]]), [[0, 0, 0]])
    f = np.array([[0.1, 0.9, 0.4], [0.6, 0.2, 0.7]], np.float32)
    back = port.box_apply(box, False, "fractional", port.box_apply(box, False, "absolute", f))
    assert np.allclose(back, f, atol=1e-6)
    # the numpy Box (host utility) and the port agree bit for bit on random inputs in a triclinic box
    tb = Box(7, 8, 9, 0.3, -0.2, 0.1)
    v = (np.random.RandomState(0).random_sample((500, 3)) * 40 - 20).astype(np.float32)
    for op, fn in (("wrap", tb.wrap), ("fractional", tb.make_fractional), ("absolute", tb.make_absolute)):
        assert np.array_equal(bits(port.box_apply(tb, False, op, v)), bits(fn(v))), op


def test_known_answer_hand_built_ball_queries():
    """tests/test_locality_neighbor_query.py:94-155 (bond counts) and :218-251 (r_min)."""
    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0], [2, 0, 0]], np.float32)
    for flavour in (port.WRAP, port.IMAGE):
        nl = port.ball_nlist(flavour, box, False, pts, pts, 2.01)
        assert list(nl.counts) == [3, 4, 3, 4]
        assert len(port.ball_nlist(flavour, box, False, pts, pts, 2.01, 0.0, True)) == 10
        moved = pts.copy()
        moved[0] = 5
        assert list(port.ball_nlist(flavour, box, False, moved, moved, 2.01).counts) == [1, 3, 3, 3]
        nl = port.ball_nlist(flavour, box, False, pts, pts, 2.9, 1.1, True)
        assert [set(nl.neighbors[nl.neighbors[:, 0] == i, 1]) for i in range(4)] == [{3}, {2}, {1}, {0}]


def test_known_answer_hand_built_nearest_queries():
    """tests/test_locality_neighbor_query.py:253-300."""
    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0], [2, 0, 0]], np.float32)

    def sets(nl):
        return [set(int(j) for j in nl.neighbors[nl.neighbors[:, 0] == i, 1]) for i in range(4)]

    assert sets(port.knn_nlist(box, False, pts, pts, 3)) == [{0, 1, 3}, {0, 1, 3}, {1, 2, 3}, {1, 2, 3}]
    assert sets(port.knn_nlist(box, False, pts, pts, 3, exclude_ii=True)) == [{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}]
    assert sets(port.knn_nlist(box, False, pts, pts, 5, exclude_ii=True)) == [{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}]
    assert sets(port.knn_nlist(box, False, pts, pts, 3, r_max=1.9, exclude_ii=True)) == [{1}, {0, 3}, {3}, {1, 2}]
    assert sets(port.knn_nlist(box, False, pts, pts, 3, r_min=1.1, exclude_ii=True)) == [{2, 3}, {2}, {0, 1}, {0}]


def test_known_answer_rdf_bin_edges_and_2d_ring():
    """tests/test_density_rdf.py:242-251 (bin edges) and :188-224 (exact n(r) for a 2-D ring system)."""
    red = port.rdf_reduce(np.zeros(10, np.uint32), 5.0, 0.0, Box.cube(20), False, 10, 10)
    assert np.allclose(red["bin_edges"], np.linspace(0, 5, 11), atol=1e-6)
    assert np.allclose(red["bin_centers"], np.linspace(0.25, 4.75, 10), atol=1e-6)
    # one point in the middle, `num` points on a ring of radius r: n(r) jumps from 0 to num at the ring
    num, radius = 20, 2.0
    box = Box.square(10)
    ang = np.linspace(0, 2 * np.pi, num, endpoint=False)
    ring = np.stack([radius * np.cos(ang), radius * np.sin(ang), 0 * ang], 1).astype(np.float32)
    centre = np.zeros((1, 3), np.float32)
    for flavour in (port.WRAP, port.IMAGE):
        counts = port.rdf_accumulate(flavour, box, True, ring, centre, 50, 3.0, 0.1, False)
        red = port.rdf_reduce(counts, 3.0, 0.1, box, True, num, 1)
        edges = red["bin_edges"]
        want = np.where(edges[1:] > radius, num, 0).astype(np.float32)
        assert np.allclose(red["n_r"], want, atol=1e-5)


@needs_ref
@pytest.mark.parametrize("name", list(BOXES))
def test_port_matches_compiled_reference(name):
    box, n, r = BOXES[name]
    n = min(n, 1500)
    pts = random_points(box, n, seed=41)
    q = random_points(box, 300, seed=42)
    cw = min(r, 0.4 * float(min(box.Lx, box.Ly)))
    lc = ref.Query("linkcell", box, pts, is2d=box.is2D, cell_width=cw)
    aq = ref.Query("aabb", box, pts, is2d=box.is2D)
    for excl in (True, False):
        for flavour, query in ((port.WRAP, lc), (port.IMAGE, aq)):
            a = query.nlist(pts, r_max=r, exclude_ii=excl)
            b = port.ball_nlist(flavour, box, box.is2D, pts, pts, r, 0.0, excl)
            assert np.array_equal(a.neighbors, b.neighbors) and np.array_equal(bits(a.distances), bits(b.distances))
            assert np.array_equal(bits(a.vectors), bits(b.vectors)) and np.array_equal(a.segments, b.segments)
    a = aq.nlist(q, r_max=r, r_min=1.0, sort_by_distance=True)
    b = port.ball_nlist(port.IMAGE, box, box.is2D, pts, q, r, 1.0, False, True)
    assert np.array_equal(a.neighbors, b.neighbors) and np.array_equal(bits(a.distances), bits(b.distances))
    a = aq.nlist(q, num_neighbors=12, exclude_ii=False)
    b = port.knn_nlist(box, box.is2D, pts, q, 12)
    assert np.array_equal(a.neighbors, b.neighbors) and np.array_equal(bits(a.distances), bits(b.distances))
    R = ref.RDF(50, r, 0.5, finite_size=True)
    R.accumulate(ref.Query("raw", box, pts, is2d=box.is2D), pts, mode="ball", r_max=r, exclude_ii=True)
    want = R.results()
    counts = port.rdf_accumulate(port.IMAGE, box, box.is2D, pts, pts, 50, r, 0.5, True)
    got = port.rdf_reduce(counts, r, 0.5, box, box.is2D, n, n, finite_size=True)
    for key in want:
        assert np.array_equal(bits(want[key]), bits(got[key])), key
    # Steinhardt from a neighbour list: identical libm on the same host -> identical bits
    nl = aq.nlist(pts, num_neighbors=8, exclude_ii=True)
    want = ref.Steinhardt([4, 6]).compute(aq, nlist=nl)
    got = port.steinhardt(box, box.is2D, pts, port.knn_nlist(box, box.is2D, pts, pts, 8, exclude_ii=True), [4, 6])
    assert np.array_equal(bits(want["ql"]), bits(got["ql"]))


@needs_ref
def test_port_matches_compiled_reference_large():
    """SURVEY.md section 8c, plan (ii): before the grid-based port may stand in for the (quadratic) reference LinkCell at
    N = 1e6 it is pinned to the compiled reference at the largest sizes the reference finishes in seconds: AABBQuery at
    N = 1e5 (r = 3 and r = 5, RDF counts too) and the unmodified LinkCell at N = 2e4, bit for bit."""
    ref.set_num_threads(0)
    n = 100_000
    box, pts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=3)
    aq = ref.Query("aabb", box, pts)
    for r in (3.0, 5.0):
        a = aq.nlist(pts, mode="ball", r_max=r, exclude_ii=True)
        b = port.ball_nlist(port.IMAGE, box, False, pts, pts, r, 0.0, True)
        assert np.array_equal(a.neighbors, b.neighbors) and np.array_equal(bits(a.distances), bits(b.distances))
        assert np.array_equal(bits(a.vectors), bits(b.vectors))
        assert np.array_equal(a.segments, b.segments) and np.array_equal(a.counts, b.counts)
    R = ref.RDF(100, 5.0)
    R.accumulate(ref.Query("raw", box, pts), pts, mode="ball", r_max=5.0, exclude_ii=True)
    assert np.array_equal(R.results()["bin_counts"], port.rdf_accumulate(port.IMAGE, box, False, pts, pts, 100, 5.0, 0.0, True))
    # triclinic at the same size (configs[3]'s tilts)
    tbox, tpts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=4, tilt=(0.3, 0.2, 0.1))
    a = ref.Query("aabb", tbox, tpts).nlist(tpts, mode="ball", r_max=5.0, exclude_ii=True)
    b = port.ball_nlist(port.IMAGE, tbox, False, tpts, tpts, 5.0, 0.0, True)
    assert np.array_equal(a.neighbors, b.neighbors) and np.array_equal(bits(a.vectors), bits(b.vectors))
    # the reference's own LinkCell (deep-copies its cell list per visited cell: N = 2e4 is what finishes in seconds)
    n = 20_000
    box, pts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=5)
    a = ref.Query("linkcell", box, pts, cell_width=3.0).nlist(pts, mode="ball", r_max=3.0, exclude_ii=True)
    b = port.ball_nlist(port.WRAP, box, False, pts, pts, 3.0, 0.0, True)
    assert len(a) > 150_000
    assert np.array_equal(a.neighbors, b.neighbors) and np.array_equal(bits(a.distances), bits(b.distances))
    assert np.array_equal(bits(a.vectors), bits(b.vectors)) and np.array_equal(a.segments, b.segments)
    # 2-D at configs[4]'s density
    n = 100_000
    box, pts = data.make_random_system((n / 0.5) ** 0.5, n, is2D=True, seed=6)
    a = ref.Query("aabb", box, pts, is2d=True).nlist(pts, mode="ball", r_max=5.0, exclude_ii=True)
    b = port.ball_nlist(port.IMAGE, box, True, pts, pts, 5.0, 0.0, True)
    assert np.array_equal(a.neighbors, b.neighbors) and np.array_equal(bits(a.vectors), bits(b.vectors))


@needs_ref
def test_reference_reproduces_its_own_constants():
    """The compiled reference itself: Q6 = 0.57452416, W6 = -0.00262604 (tests/test_order_steinhardt.py:17-18)."""
    box, pts = data.make_fcc_system(4)
    q = ref.Query("raw", box, pts)
    assert np.allclose(ref.Steinhardt(6).compute(q, num_neighbors=12, exclude_ii=True)["ql"], 0.57452416, atol=1e-5)
    w6 = ref.Steinhardt(6, wl=True).compute(q, num_neighbors=12, exclude_ii=True)["particle_order"]
    assert np.allclose(w6, -0.00262604, atol=1e-5)
    assert np.array_equal(ref.box_apply((2, 2, 2, 1, 0, 0), False, "wrap", [[10, -5, -5]]), [[-2, -1, -1]])
