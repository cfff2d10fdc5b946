"""The reference-facing API (freud_b200.locality / density / order -> C++ host classes -> C ABI -> CUDA) against the
oracle.  Cases follow the reference's own suite: tests/test_locality_neighbor_query.py, test_density_rdf.py,
test_order_steinhardt.py, test_managedarray.py (lines cited per test)."""
import os

import numpy as np
import pytest

from freud_b200 import data, density, environment, locality, order, pmft
from freud_b200.box import Box
from oracle import port, ref
from tests.util import BOXES, bits, random_points

pytestmark = pytest.mark.gpu

# upstream builds LinkCell with cell_width = r_max (tests/test_locality_neighbor_query.py:662-666); the default width
# (10 particles per cell) is rejected for tiny systems exactly as in the reference (LinkCell.cc:241-246)
ENGINES = {"aabb": locality.AABBQuery, "linkcell": lambda box, pts: locality.LinkCell(box, pts, cell_width=2.0),
           "raw": lambda box, pts: (box, pts)}


def make_nq(engine, box, pts):
    obj = ENGINES[engine](box, pts)
    return locality.NeighborQuery.from_system(obj)


def assert_matches_oracle(nl, want, what):
    assert len(nl) == len(want), what
    assert np.array_equal(nl[:], want.neighbors), what
    assert np.array_equal(bits(nl.distances), bits(want.distances)), f"{what}: distances differ bitwise"
    assert np.array_equal(bits(nl.vectors), bits(want.vectors)), f"{what}: vectors differ bitwise"
    assert np.array_equal(nl.segments, want.segments) and np.array_equal(nl.neighbor_counts, want.counts), what


@pytest.mark.parametrize("engine", list(ENGINES))
def test_query_ball_hand_built(engine):
    """tests/test_locality_neighbor_query.py:94-155, :218-251 upstream."""
    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0], [2, 0, 0]], np.float32)
    nq = make_nq(engine, box, pts)
    nl = nq.query(pts, dict(r_max=2.01)).toNeighborList()
    assert list(nl.neighbor_counts) == [3, 4, 3, 4]
    nl = nq.query(pts, dict(r_max=2.01, exclude_ii=True)).toNeighborList()
    assert len(nl) == 10
    bonds = {(int(i), int(j)) for i, j, d in nq.query(pts, dict(r_max=2.9, r_min=1.1, exclude_ii=True))}
    assert bonds == {(0, 3), (1, 2), (2, 1), (3, 0)}
    # zero query points -> empty (0, 2) list (tests/test_locality_neighbor_list.py:253-256)
    nl = nq.query(np.zeros((0, 3), np.float32), dict(r_max=2.0)).toNeighborList()
    assert len(nl) == 0 and nl[:].shape == (0, 2)


@pytest.mark.parametrize("engine", ["aabb", "linkcell"])
def test_query_nearest_hand_built(engine):
    """tests/test_locality_neighbor_query.py:253-300 upstream."""
    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0], [2, 0, 0]], np.float32)
    nq = make_nq(engine, box, pts)
    nl = nq.query(pts, dict(num_neighbors=3, exclude_ii=True)).toNeighborList()
    rows = [set(int(j) for j in nl.point_indices[nl.query_point_indices == i]) for i in range(4)]
    assert rows == [{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}]
    nl = nq.query(pts, dict(num_neighbors=3, r_max=1.9, exclude_ii=True)).toNeighborList()
    rows = [set(int(j) for j in nl.point_indices[nl.query_point_indices == i]) for i in range(4)]
    assert rows == [{1}, {0, 3}, {3}, {1, 2}]


@pytest.mark.parametrize("engine", ["aabb", "linkcell"])
@pytest.mark.parametrize("seed", range(4))
def test_exhaustive_search_vs_box_wrap(engine, seed):
    """tests/test_locality_neighbor_query.py:395-448 upstream: the pair SET equals brute force with box.wrap."""
    box, n, r = Box.cube(11), 400, 2.1
    pts = random_points(box, n, seed)
    q = random_points(box, 150, seed + 100)
    nq = make_nq(engine, box, pts)
    nl = nq.query(q, dict(r_max=r)).toNeighborList()
    d = box.wrap((pts[None, :, :] - q[:, None, :]).reshape(-1, 3)).reshape(len(q), n, 3)
    dist = np.sqrt((d.astype(np.float64) ** 2).sum(-1))
    want = {(i, j) for i, j in zip(*np.nonzero(dist < r - 1e-4))}
    maybe = {(i, j) for i, j in zip(*np.nonzero(dist < r + 1e-4))}
    got = {(int(i), int(j)) for i, j in nl[:]}
    assert want <= got <= maybe


@pytest.mark.parametrize("name", ["cubic", "tri1", "sq2d"])
@pytest.mark.parametrize("engine", ["aabb", "linkcell", "raw"])
def test_engines_match_the_reference_bit_for_bit(name, engine):
    box, n, r = BOXES[name]
    pts = random_points(box, n, seed=41)
    q = random_points(box, 300, seed=42)
    nq = make_nq(engine, box, pts)
    flavour = port.WRAP if engine == "linkcell" else port.IMAGE
    for sbd in (False, True):
        nl = nq.query(q, dict(r_max=r, r_min=0.5)).toNeighborList(sort_by_distance=sbd)
        want = port.ball_nlist(flavour, box, box.is2D, pts, q, r, 0.5, False, sbd)
        assert_matches_oracle(nl, want, f"{name} {engine} sbd={sbd}")
    nl = nq.query(pts, dict(r_max=r, exclude_ii=True)).toNeighborList()
    assert_matches_oracle(nl, port.ball_nlist(flavour, box, box.is2D, pts, pts, r, 0.0, True), f"{name} {engine} self")
    # arrays are read-only (tests/test_locality_neighbor_list.py:26-39)
    with pytest.raises(ValueError):
        nl.distances[0] = 0


@pytest.mark.parametrize("system_kind", ["tuple", "aabb", "linkcell"])
def test_rdf_matches_the_reference(system_kind):
    """density.RDF.compute on BASELINE.json configs[0]-shaped input: counts bit-exact, g(r) / n(r) within 1e-5."""
    box, pts = data.make_random_system(30, 4000, seed=3)
    system = {"tuple": (box, pts), "aabb": locality.AABBQuery(box, pts), "linkcell": locality.LinkCell(box, pts, 5.0)}[
        system_kind]
    flavour = port.WRAP if system_kind == "linkcell" else port.IMAGE
    rdf = density.RDF(bins=100, r_max=5.0)
    rdf.compute(system)
    counts = port.rdf_accumulate(flavour, box, False, pts, pts, 100, 5.0, 0.0, True)
    want = port.rdf_reduce(counts, 5.0, 0.0, box, False, len(pts), len(pts))
    assert np.array_equal(rdf.bin_counts, counts)
    np.testing.assert_allclose(rdf.rdf, want["rdf"], rtol=1e-5, atol=0)
    np.testing.assert_allclose(rdf.n_r, want["n_r"], rtol=1e-5, atol=0)
    assert np.array_equal(rdf.bin_edges, want["bin_edges"]) and np.array_equal(rdf.bin_centers, want["bin_centers"])
    if ref.available() and system_kind != "linkcell":
        R = ref.RDF(100, 5.0)
        R.accumulate(ref.Query("raw", box, pts), pts, mode="ball", r_max=5.0, exclude_ii=True)
        res = R.results()
        assert np.array_equal(rdf.bin_counts, res["bin_counts"])
        np.testing.assert_allclose(rdf.rdf, res["rdf"], rtol=1e-5, atol=0)
        np.testing.assert_allclose(rdf.n_r, res["n_r"], rtol=1e-5, atol=0)
    # statistical sanity of the reference suite (tests/test_density_rdf.py:94-127): g(r) ~ 1 for an ideal gas
    assert abs(float(np.mean(rdf.rdf[20:])) - 1.0) < 0.02 and np.all(np.abs(rdf.rdf[40:] - 1.0) < 0.1)


def test_rdf_accumulate_reset_false_query_points_and_nlist():
    """tests/test_density_rdf.py:129-165 upstream (+ separate query points, a precomputed NeighborList, 2-D)."""
    box, pts = data.make_random_system(40, 3000, is2D=True, seed=5)
    _, pts2 = data.make_random_system(40, 3000, is2D=True, seed=6)
    q = random_points(box, 500, seed=7)
    rdf = density.RDF(50, 4.0, 0.5, normalization_mode="finite_size")
    rdf.compute((box, pts), reset=False)
    first = rdf.bin_counts  # a view handed out BEFORE the second frame
    first_copy = first.copy()
    rdf.compute((box, pts2), query_points=q, reset=False)
    c = port.rdf_accumulate(port.IMAGE, box, True, pts, pts, 50, 4.0, 0.5, True)
    assert np.array_equal(first_copy, c)
    c2 = port.rdf_accumulate(port.IMAGE, box, True, pts2, q, 50, 4.0, 0.5, False, counts=c.copy())
    assert np.array_equal(rdf.bin_counts, c2)
    want = port.rdf_reduce(c2, 4.0, 0.5, box, True, len(pts2), len(q), frames=2, finite_size=True)
    np.testing.assert_allclose(rdf.rdf, want["rdf"], rtol=1e-5)
    np.testing.assert_allclose(rdf.n_r, want["n_r"], rtol=1e-5)
    # reset=True starts over; a NeighborList as `neighbors` is binned bond by bond
    nl = locality.AABBQuery(box, pts).query(pts, dict(r_max=4.0, exclude_ii=True)).toNeighborList()
    rdf.compute((box, pts), neighbors=nl)
    assert np.array_equal(rdf.bin_counts, c)
    # output-lifetime contract (tests/test_managedarray.py:25-53): reset() hands out new arrays
    old = rdf.rdf
    old_copy = old.copy()
    rdf.compute((box, pts2))
    assert np.array_equal(old, old_copy) and not np.array_equal(rdf.rdf, old_copy)


def test_rdf_empty_histogram():
    """tests/test_density_rdf.py:226-237 upstream: far-apart points give zeros, no NaN in n(r)."""
    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [4, 4, 4]], np.float32)
    rdf = density.RDF(10, 1.0).compute((box, pts))
    assert not rdf.bin_counts.any() and not rdf.rdf.any() and not rdf.n_r.any()


def test_steinhardt_fcc_known_answers():
    """PERFECT_FCC_Q6 = 0.57452416 (tests/test_order_steinhardt.py:17, :101-166 upstream) for k = 12 and a ball
    query, through every system kind."""
    box, pts = data.make_fcc_system(4)
    for system in ((box, pts), locality.AABBQuery(box, pts), locality.LinkCell(box, pts, 1.0)):
        st = order.Steinhardt(6).compute(system, neighbors=dict(num_neighbors=12))
        np.testing.assert_allclose(st.particle_order, 0.57452416, atol=1e-5)
        assert abs(st.order - 0.57452416) < 1e-5
        st = order.Steinhardt(6).compute(system, neighbors=dict(r_max=0.8))
        np.testing.assert_allclose(st.ql, 0.57452416, atol=1e-5)
    # a NeighborList as neighbours, several l at once, harmonics shape
    nl = locality.AABBQuery(box, pts).query(pts, dict(num_neighbors=12, exclude_ii=True)).toNeighborList()
    st = order.Steinhardt([4, 6]).compute((box, pts), neighbors=nl)
    assert st.particle_order.shape == (len(pts), 2) and [h.shape[1] for h in st.particle_harmonics] == [9, 13]
    np.testing.assert_allclose(st.particle_order[:, 1], 0.57452416, atol=1e-5)


def test_steinhardt_vs_reference_noisy_fcc():
    box, pts = data.make_fcc_system(5, scale=1.2, sigma_noise=0.07, seed=9)
    st = order.Steinhardt([4, 6]).compute((box, pts), neighbors=dict(num_neighbors=12))
    if ref.available():
        want = ref.Steinhardt([4, 6]).compute(ref.Query("raw", box, pts), num_neighbors=12, exclude_ii=True)
        np.testing.assert_allclose(st.ql, want["ql"], rtol=1e-5, atol=1e-6)
        for a, b in zip(st.particle_harmonics, want["qlm"]):
            np.testing.assert_allclose(a, b, atol=1e-5)
        np.testing.assert_allclose(st.order, want["order"], rtol=1e-4)
    else:
        pnl = port.knn_nlist(box, False, pts, pts, 12, exclude_ii=True)
        want = port.steinhardt(box, False, pts, pnl, [4, 6])
        np.testing.assert_allclose(st.ql, want["ql"], rtol=1e-5, atol=1e-6)


def test_pmftxy_api():
    """freud.pmft.PMFTXY (tests/test_pmft.py upstream): bin counts and PCF bit for bit against the reference's committed
    outputs, pmft = -log(pcf), histogram properties, reset=False, 3-D boxes refused."""
    from tests.golden.make_golden import pmftxy_inputs

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pmftxy.npz"))
    box = Box(30, 26, 0, 0.35, 0, 0, is2D=True)
    pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
    th_p, th_q = pmftxy_inputs(3000, 800, 5)
    pm = pmft.PMFTXY(3.0, 2.5, (30, 24))
    assert pm.nbins == (30, 24) and np.isclose(pm.r_max, np.sqrt(3.0 ** 2 + 2.5 ** 2))
    assert pm.bounds == [(-3.0, 3.0), (-2.5, 2.5)] and [len(e) for e in pm.bin_edges] == [31, 25]
    pm.compute((box, pts), th_q, query_points=q)
    assert np.array_equal(pm.bin_counts, gold["tilt2d_query_counts"])
    assert np.array_equal(bits(pm._pcf), bits(gold["tilt2d_query_pcf"]))
    with np.errstate(divide="ignore"):
        assert np.array_equal(pm.pmft, -np.log(gold["tilt2d_query_pcf"]))
    pm.compute((box, pts), th_q, query_points=q, reset=False)  # a second identical frame: same PCF
    assert np.array_equal(pm.bin_counts, 2 * gold["tilt2d_query_counts"])
    assert np.array_equal(bits(pm._pcf), bits(gold["tilt2d_query_pcf"]))
    pm.compute((box, pts), th_p)
    assert np.array_equal(pm.bin_counts, gold["tilt2d_self_counts"])
    assert np.array_equal(bits(pm._pcf), bits(gold["tilt2d_self_pcf"])) and pm.box == box
    # quaternions about +z are reduced to their angle (freud/pmft.py:58-84: rowan's axis-angle, so the angle is taken
    # in [0, 2 pi) -- a quaternion with a negative z component is a rotation about -z and is refused, as upstream)
    th_pos = np.mod(th_p.astype(np.float64), 2 * np.pi)
    quats = np.stack([np.cos(th_pos / 2), np.zeros_like(th_pos), np.zeros_like(th_pos), np.sin(th_pos / 2)], axis=1)
    counts_q = pmft.PMFTXY(3.0, 2.5, (30, 24)).compute((box, pts), quats).bin_counts
    assert abs(int(counts_q.sum()) - int(gold["tilt2d_self_counts"].sum())) < 50  # angles differ by rounding only
    with pytest.raises(ValueError):
        pmft.PMFTXY(3.0, 2.5, (30, 24)).compute((box, pts), quats * np.array([1.0, 1.0, 1.0, -1.0]))
    with pytest.raises(ValueError):
        pmft.PMFTXY(3.0, 2.5, 10).compute((Box.cube(10), random_points(Box.cube(10), 100, 1)), np.zeros(100))
    with pytest.raises(ValueError):
        pmft.PMFTXY(3.0, 2.5, (0, 4))


def test_pmft3_api():
    """freud.pmft.PMFTXYZ / PMFTXYT / PMFTR12 (tests/test_pmft.py upstream): bin counts and PCF bit for bit against the
    reference's committed outputs (tests/golden/pmft3.npz) -- incl. the lattice whose bond angles all sit on bin edges,
    where the bin hangs on the last place of libm's atan2f and the host decides -- reset=False, properties, errors."""
    from tests.golden.make_golden import PMFT3_EQUIV, pmft3_lattice, pmft3_quats, pmftxy_inputs

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pmft3.npz"))

    def same(pm, tag, frames=1):
        assert np.array_equal(pm.bin_counts, frames * gold[f"{tag}_counts"]), tag
        assert np.array_equal(bits(pm._pcf), bits(gold[f"{tag}_pcf"])), tag

    # XYZ
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 1500, 31), random_points(box, 400, 32)
    xyz = pmft.PMFTXYZ(2.0, 2.5, 3.0, (12, 10, 8))
    assert xyz.nbins == (12, 10, 8) and xyz.bounds == [(-2.0, 2.0), (-2.5, 2.5), (-3.0, 3.0)]
    assert np.isclose(xyz.r_max, np.sqrt(2.0 ** 2 + 2.5 ** 2 + 3.0 ** 2)) and [len(c) for c in xyz.bin_centers] == [12, 10, 8]
    xyz.compute((box, pts), pmft3_quats(400, 6), query_points=q, equiv_orientations=PMFT3_EQUIV)
    same(xyz, "xyz_query")
    assert xyz.bin_counts.shape == (12, 10, 8) and xyz.box == box
    xyz.compute((box, pts), pmft3_quats(400, 6), query_points=q, equiv_orientations=PMFT3_EQUIV, reset=False)
    same(xyz, "xyz_query", frames=2)
    with pytest.raises(RuntimeError):  # the number of equivalent orientations may not change while accumulating
        xyz.compute((box, pts), pmft3_quats(400, 6), query_points=q, reset=False)
    xyz.compute((box, pts), pmft3_quats(1500, 7))  # default: the identity only; self query excludes i == j
    same(xyz, "xyz_self")
    shifted = pmft.PMFTXYZ(2.0, 2.5, 3.0, (12, 10, 8), shiftvec=[0.5, 0, 0])
    shifted.compute((box, pts), pmft3_quats(400, 6), query_points=q + np.float32([0.5, 0, 0]), equiv_orientations=PMFT3_EQUIV)
    assert abs(int(shifted.bin_counts.sum()) - int(gold["xyz_query_counts"].sum())) < 200  # same bonds up to rounding
    with pytest.raises(ValueError):
        sq = Box.square(20)
        pmft.PMFTXYZ(1, 1, 1, 4).compute((sq, random_points(sq, 50, 1)), pmft3_quats(50, 1))
    with pytest.raises(ValueError):
        pmft.PMFTXYZ(1, 1, -1, 4)

    # XYT and R12
    for name, box in (("sq2d", Box.square(40)), ("tilt2d", Box(30, 26, 0, 0.35, 0, 0, is2D=True))):
        pts, q = random_points(box, 3000, 11), random_points(box, 800, 12)
        th_p, th_q = pmftxy_inputs(3000, 800, 5)
        xyt = pmft.PMFTXYT(3.0, 2.5, (14, 12, 9))
        xyt.compute((box, pts), th_p, query_points=q, query_orientations=th_q)
        same(xyt, f"{name}_xyt_query")
        xyt.compute((box, pts), th_p, query_points=q, query_orientations=th_q, reset=False)
        same(xyt, f"{name}_xyt_query", frames=2)
        xyt.compute((box, pts), th_p)
        same(xyt, f"{name}_xyt_self")
        r12 = pmft.PMFTR12(4.0, (10, 11, 12))
        r12.compute((box, pts), th_p, query_points=q, query_orientations=th_q)
        same(r12, f"{name}_r12_query")
        r12.compute((box, pts), th_p)
        same(r12, f"{name}_r12_self")
        # random angles: only a sliver of the bonds sits close enough to a bin edge to need the host
        assert 0 <= r12.host_binned_bonds < 0.01 * int(gold[f"{name}_r12_self_counts"].sum()) + 50
    assert xyt.bounds[2] == (0.0, float(np.float32(2 * np.pi))) and r12.bounds[0] == (0.0, 4.0)
    box, pts, th = pmft3_lattice()
    xyt = pmft.PMFTXYT(3.0, 3.0, (6, 6, 8)).compute((box, pts), th)
    same(xyt, "lattice_xyt")
    r12 = pmft.PMFTR12(3.0, (6, 8, 8)).compute((box, pts), th)
    same(r12, "lattice_r12")
    assert xyt.host_binned_bonds > 1000 and r12.host_binned_bonds > 1000  # every bond angle is on a bin edge here
    with pytest.raises(ValueError):
        cube = Box.cube(10)
        pmft.PMFTXYT(3.0, 2.5, 5).compute((cube, random_points(cube, 100, 1)), np.zeros(100))
    with pytest.raises(ValueError):
        pmft.PMFTR12(3.0, (4, 0, 4))
    with pytest.raises(ValueError):
        pmft.PMFTR12(3.0, 4).compute((box, pts), th[:10])


def test_pmft3_matches_the_port_at_scale():
    """100 k particles (2.5 M bonds in 2-D, 3 M in 3-D): bin counts of all three classes bit for bit against oracle/port.c
    over the same NeighborList; the host's share of the angle bins stays a sliver."""
    from tests.golden.make_golden import pmft3_quats

    rs = np.random.RandomState(9)
    box, pts = data.make_random_system(450.0, 100_000, is2D=True, seed=4)
    th = (rs.random_sample(len(pts)) * 4 * np.pi - 2 * np.pi).astype(np.float32)  # also outside [0, 2 pi)
    th[::7] *= np.float32(300.0)  # many turns: the kernel's own remainder ...
    th[::41] *= np.float32(1.0e4)  # ... and the library's for huge angles
    th[5], th[6] = np.float32(2 * np.pi), np.float32(-4 * np.pi)
    nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, 4.0, exclude_ii=True)
    xyt = pmft.PMFTXYT(3.0, 2.5, (40, 30, 36)).compute((box, pts), th, neighbors=dict(mode="ball", r_max=4.0))
    want, want_pcf = port.pmft3(port.PMFT_XYT, box, len(pts), nl, th, th, (3.0, 2.5), (40, 30, 36))
    assert np.array_equal(xyt.bin_counts, want) and np.array_equal(bits(xyt._pcf), bits(want_pcf))
    r12 = pmft.PMFTR12(4.0, (20, 36, 36)).compute((box, pts), th, neighbors=dict(mode="ball", r_max=4.0))
    want, want_pcf = port.pmft3(port.PMFT_R12, box, len(pts), nl, th, th, (4.0,), (20, 36, 36))
    assert np.array_equal(r12.bin_counts, want) and np.array_equal(bits(r12._pcf), bits(want_pcf))
    n_bonds = len(nl.distances)
    assert xyt.host_binned_bonds < 2e-3 * n_bonds and r12.host_binned_bonds < 4e-3 * n_bonds
    box, pts = data.make_random_system(60.0, 100_000, seed=5)
    quats = pmft3_quats(len(pts), 8)
    equiv = pmft3_quats(6, 9)
    nl = port.ball_nlist(port.IMAGE, box, False, pts, pts, 2.5, exclude_ii=True)
    xyz = pmft.PMFTXYZ(2.0, 2.0, 2.0, (24, 20, 16)).compute((box, pts), quats, equiv_orientations=equiv,
                                                           neighbors=dict(mode="ball", r_max=2.5))
    want, want_pcf = port.pmft3(port.PMFT_XYZ, box, len(pts), nl, None, quats, (2.0, 2.0, 2.0), (24, 20, 16), equiv=equiv)
    assert np.array_equal(xyz.bin_counts, want) and np.array_equal(bits(xyz._pcf), bits(want_pcf))


def test_bond_order_api():
    """freud.environment.BondOrder (tests/test_environment_BondOrder.py upstream): bin counts and the diagram bit for bit
    against the reference's committed outputs in all four modes, the FCC lattice whose bond directions sit on bin edges
    (the host's libm decides those), reset=False, default orientations, properties, errors."""
    from tests.golden.make_golden import pmft3_quats

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bond_order.npz"))
    box = Box(12, 13, 14, 0.2, -0.1, 0.15)
    pts, q = random_points(box, 800, 41), random_points(box, 300, 42)
    o, qo = pmft3_quats(800, 1), pmft3_quats(300, 2)
    knn = dict(mode="nearest", num_neighbors=8)
    for mode in ("bod", "lbod", "obcd", "oocd"):
        bo = environment.BondOrder((12, 9), mode=mode)
        bo.compute((box, pts), o, query_points=q, query_orientations=qo, neighbors=knn)
        assert np.array_equal(bo.bin_counts, gold[f"tri_{mode}_counts"]), mode
        assert np.array_equal(bits(bo.bond_order), bits(gold[f"tri_{mode}_bo"])), mode
        assert bo.mode == mode and bo.nbins == (12, 9) and bo.box == box
    bo.compute((box, pts), o, query_points=q, query_orientations=qo, neighbors=knn, reset=False)  # two equal frames
    assert np.array_equal(bo.bin_counts, 2 * gold["tri_oocd_counts"])
    assert np.array_equal(bits(bo.bond_order), bits(gold["tri_oocd_bo"]))
    assert bo.bounds == [(0.0, float(np.float32(2 * np.pi))), (0.0, float(np.float32(np.pi)))]
    assert [len(e) for e in bo.bin_edges] == [13, 10] and [len(c) for c in bo.bin_centers] == [12, 9]
    ident = np.tile(np.float32([1, 0, 0, 0]), (300, 1))  # default orientations are identities, one per point
    plain = environment.BondOrder((12, 9)).compute((box, pts), query_points=q, query_orientations=ident, neighbors=knn)
    assert np.array_equal(plain.bin_counts, gold["tri_bod_counts"])
    with pytest.raises(ValueError):  # as upstream: query_orientations default to orientations, whose length differs
        environment.BondOrder((12, 9)).compute((box, pts), query_points=q, neighbors=knn)
    nlist = locality.AABBQuery(box, pts).query(q, knn).toNeighborList()
    from_list = environment.BondOrder((12, 9), mode="lbod").compute((box, pts), o, query_points=q, query_orientations=qo,
                                                                    neighbors=nlist)
    assert np.array_equal(from_list.bin_counts, gold["tri_lbod_counts"])
    fbox, fpts = data.UnitCell.fcc().generate_system(4)
    for bins in ((8, 4), (7, 5)):
        fcc = environment.BondOrder(bins).compute((fbox, fpts), neighbors=dict(mode="nearest", num_neighbors=12))
        assert np.array_equal(fcc.bin_counts, gold[f"fcc_{bins[0]}x{bins[1]}_counts"])
        assert np.array_equal(bits(fcc.bond_order), bits(gold[f"fcc_{bins[0]}x{bins[1]}_bo"]))
        assert fcc.host_binned_bonds > 0
    with pytest.raises(NotImplementedError):
        environment.BondOrder(4).compute((box, pts))
    with pytest.raises(ValueError):
        environment.BondOrder((1, 4))
    with pytest.raises(ValueError):
        environment.BondOrder(4, mode="nope")
    with pytest.raises(ValueError):
        environment.BondOrder(4).compute((box, pts), o[:10], neighbors=knn)


def test_bond_order_matches_the_port_at_scale():
    """100 k particles, bonds within r = 2 (1.5 M of them), mode obcd: counts bit for bit against oracle/port.c over the
    same NeighborList; the host's share of the bins stays a sliver."""
    from tests.golden.make_golden import pmft3_quats

    box, pts = data.make_random_system(60.0, 100_000, seed=6)
    o = pmft3_quats(len(pts), 3)
    nl = port.ball_nlist(port.IMAGE, box, False, pts, pts, 2.0, exclude_ii=True)
    bo = environment.BondOrder((48, 24), mode="obcd").compute((box, pts), o, neighbors=dict(mode="ball", r_max=2.0))
    want, want_bo = port.bond_order("obcd", nl, o, o, (48, 24))
    assert np.array_equal(bo.bin_counts, want) and np.array_equal(bits(bo.bond_order), bits(want_bo))
    assert 0 < bo.host_binned_bonds < 5e-3 * len(nl.distances)


def test_histogram_clients_reference_suite_cases():
    """Cases restated from the reference's own suite: the two-particle system of tests/test_pmft.py:119-154, 663-706
    (expected bins from the definition, equivalent orientations double the counts and leave the PMFT unchanged), the
    shifted dead pixel (:708-727), PMFTXY against the middle z layer of PMFTXYZ (:992-1026) and against PMFTXYT with one
    angular bin (:1028-1056); empty neighbour lists; results available only after compute()."""
    box = Box.cube(16)
    pts = np.array([[-1.0, 0.0, 0.0], [1.0, 0.1, 0.0]], dtype=np.float32)
    ident = np.array([[1, 0, 0, 0], [1, 0, 0, 0]], dtype=np.float32)
    limits, bins = np.array([3.6, 4.2, 4.8]), (20, 30, 40)

    def xyz_bin(a, b):
        return tuple(np.floor((b - a + limits) * np.asarray(bins) / (2 * limits)).astype(int))

    want = np.zeros(bins, np.uint32)
    want[xyz_bin(pts[0], pts[1])] = 1
    want[xyz_bin(pts[1], pts[0])] = 1
    xyz = pmft.PMFTXYZ(*limits, bins)
    with pytest.raises(AttributeError):
        xyz.bin_counts
    for system in ((box, pts), locality.AABBQuery(box, pts), locality.LinkCell(box, pts, 7.0)):
        xyz.compute(system, ident, reset=False)
        xyz.compute(system, ident)
        assert np.array_equal(xyz.bin_counts, want)
    first = xyz.pmft
    xyz.compute((box, pts), ident, equiv_orientations=[[1, 0, 0, 0]] * 2)
    assert np.array_equal(xyz.bin_counts, 2 * want)
    with np.errstate(invalid="ignore"):
        assert np.allclose(xyz.pmft[np.isfinite(first)], first[np.isfinite(first)], atol=1e-6)
    # r12 / xyt of the same pair in 2-D, bins from the definitions (tests/test_pmft.py:214-232, 641-660)
    sq = Box.square(16)
    two_pi = 2 * np.pi
    r_ij = pts[1] - pts[0]
    r12 = pmft.PMFTR12(5.23, (10, 20, 30)).compute((sq, pts), np.zeros(2))
    want = np.zeros((10, 20, 30), np.uint32)
    for v in (r_ij, -r_ij):
        want[int(np.linalg.norm(v) * 10 / 5.23), int(((0 - np.arctan2(v[1], v[0])) % two_pi) * 20 / two_pi),
             int(((0 - np.arctan2(-v[1], -v[0])) % two_pi) * 30 / two_pi)] = 1
    assert np.array_equal(r12.bin_counts, want)
    # the shifted dead pixel
    cube3, pair = Box.cube(3), np.array([[1, 1, 1], [0, 0, 0]], dtype=np.float32)
    noshift = pmft.PMFTXYZ(0.5, 0.5, 0.5, 3).compute((cube3, pair), ident)
    shift = pmft.PMFTXYZ(0.5, 0.5, 0.5, 3, shiftvec=[1, 1, 1]).compute((cube3, pair), ident)
    with np.errstate(divide="ignore"):
        assert np.isfinite(noshift.pmft).sum() == 0 and np.isfinite(shift.pmft).sum() == 1
    # XY == the middle z layer of XYZ == XYT with a single angular bin
    rs = np.random.RandomState(0)
    flat = rs.random_sample((100, 3)).astype(np.float32)
    flat[:, 2] = 0
    quat0 = np.tile(np.float32([1, 0, 0, 0]), (100, 1))
    xy = pmft.PMFTXY(2.5, 2.5, 4).compute((Box.square(10), flat), quat0)
    xyz4 = pmft.PMFTXYZ(2.5, 2.5, 1, 4).compute((Box.cube(10), flat), quat0)
    assert np.array_equal(xy.bin_counts, xyz4.bin_counts[:, :, 2])
    with np.errstate(divide="ignore", invalid="ignore"):
        a, b = np.exp(xy.pmft), np.exp(xyz4.pmft[:, :, 2]) * (4 / 2) * 10
    assert np.allclose(a, b, atol=1e-6)
    xy3 = pmft.PMFTXY(2.5, 2.5, 3).compute((Box.square(10), flat), np.zeros(100))
    xyt = pmft.PMFTXYT(2.5, 2.5, (3, 3, 1)).compute((Box.square(10), flat), np.zeros(100))
    assert np.array_equal(xy3.bin_counts, xyt.bin_counts.reshape(3, 3))
    with np.errstate(divide="ignore"):
        assert np.allclose(np.exp(xy3.pmft), np.exp(xyt.pmft).reshape(3, 3), atol=1e-6)
    # nothing within reach: all-zero histograms, no error
    far = np.array([[-4, -4, 0], [4, 4, 0]], dtype=np.float32)
    assert pmft.PMFTXYT(1, 1, 4).compute((sq, far), np.zeros(2)).bin_counts.sum() == 0
    assert pmft.PMFTXYZ(1, 1, 1, 4).compute((box, far), ident).bin_counts.sum() == 0
    bo = environment.BondOrder(4).compute((box, far), neighbors=dict(r_max=1.0))
    assert bo.bin_counts.sum() == 0 and not bo.bond_order.any()


def test_pmft_routes_agree_across_engines_and_query_modes():
    """The one-call route (ball query, bonds from the search's bag) and the NeighborList route (nearest-neighbour query
    arguments, or a list handed in) of the same class give what the oracle gives over the engine's own bonds: LinkCell
    (wrap arithmetic), CellQuery (ghost arithmetic) and AABBQuery (image arithmetic), tests/test_pmft.py:455-485 upstream
    for the nearest-neighbour arguments."""
    rs = np.random.RandomState(12)
    box = Box(24, 22, 0, 0.3, 0, 0, is2D=True)
    pts = random_points(box, 2000, 51)
    th = (rs.random_sample(2000) * 2 * np.pi).astype(np.float32)
    for make, flavour in ((lambda: locality.LinkCell(box, pts, 3.0), port.WRAP), (lambda: locality.CellQuery(box, pts), port.GHOST),
                          (lambda: locality.AABBQuery(box, pts), port.IMAGE)):
        nl = port.ball_nlist(flavour, box, True, pts, pts, 3.0, 0.0, True)
        want_xyt = port.pmft3(port.PMFT_XYT, box, 2000, nl, th, th, (2.0, 2.0), (10, 10, 12))[0]
        want_xy = port.pmftxy(box, 2000, nl, th, 2.0, 2.0, 10, 10)[0]
        ball = dict(mode="ball", r_max=3.0)
        assert np.array_equal(pmft.PMFTXYT(2.0, 2.0, (10, 10, 12)).compute(make(), th, neighbors=ball).bin_counts, want_xyt)
        assert np.array_equal(pmft.PMFTXY(2.0, 2.0, 10).compute(make(), th, neighbors=ball).bin_counts, want_xy)
        handed = make().query(pts, dict(ball, exclude_ii=True)).toNeighborList()
        assert np.array_equal(pmft.PMFTXYT(2.0, 2.0, (10, 10, 12)).compute(make(), th, neighbors=handed).bin_counts, want_xyt)
    knn = port.knn_nlist(box, True, pts, pts, 6, exclude_ii=True)
    want = port.pmft3(port.PMFT_R12, box, 2000, knn, th, th, (3.0,), (6, 8, 8))[0]
    got = pmft.PMFTR12(3.0, (6, 8, 8)).compute(locality.AABBQuery(box, pts), th, neighbors=dict(num_neighbors=6))
    assert np.array_equal(got.bin_counts, want)
    cube = Box.cube(14)
    p3 = random_points(cube, 1500, 52)
    q = rs.normal(size=(1500, 4))
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    knn3 = port.knn_nlist(cube, False, p3, p3, 8, exclude_ii=True)
    want = port.pmft3(port.PMFT_XYZ, cube, 1500, knn3, None, q, (2.0, 2.0, 2.0), (8, 8, 8), equiv=np.float32([[1, 0, 0, 0]]))[0]
    got = pmft.PMFTXYZ(2.0, 2.0, 2.0, 8).compute((cube, p3), q, neighbors=dict(num_neighbors=8))
    assert np.array_equal(got.bin_counts, want)


def test_bond_order_reference_suite_case():
    """tests/test_environment_bond_order.py:14-105 upstream: a perfect FCC crystal fills exactly 12 bins of a 6 x 6 diagram;
    lbod and obcd equal bod when all orientations are equal; bod ignores random orientations while obcd smears them over
    every bin; oocd of equal orientations is a single peak at (0, 0) and of random ones is broad."""
    box, pts = data.UnitCell.fcc().generate_system(4)
    quats = np.tile(np.float32([1, 0, 0, 0]), (len(pts), 1))
    nn = dict(num_neighbors=12, r_max=1.5)
    bo = environment.BondOrder(6).compute((box, pts), quats, neighbors=nn)
    ref_bod = bo.bond_order.copy()
    assert np.sum(ref_bod > 0) == 12 and bo.box == box
    assert np.allclose(bo.bin_centers[0], (2 * np.arange(6) + 1) * np.pi / 6)
    assert np.allclose(bo.bin_centers[1], (2 * np.arange(6) + 1) * np.pi / 12)
    rs = np.random.RandomState(10893)
    rq = rs.normal(size=(len(pts), 4))
    rq = (rq / np.linalg.norm(rq, axis=1, keepdims=True)).astype(np.float32)
    nlist = locality.AABBQuery(box, pts).query(pts, dict(nn, exclude_ii=True)).toNeighborList()
    for system, neighbors in (((box, pts), nn), (locality.AABBQuery(box, pts), nn), (locality.LinkCell(box, pts, 1.5), nn),
                              ((box, pts), nlist)):
        lbod = environment.BondOrder(6, mode="lbod").compute(system, quats, neighbors=neighbors, reset=False)
        assert np.allclose(lbod.bond_order, ref_bod)
        assert np.allclose(environment.BondOrder(6, mode="obcd").compute(system, quats, neighbors=neighbors).bond_order,
                           ref_bod)
        assert np.allclose(environment.BondOrder(6).compute(system, rq, neighbors=neighbors).bond_order, ref_bod)
        smeared = environment.BondOrder(6, mode="obcd").compute(system, rq, neighbors=neighbors).bond_order
        assert not np.allclose(smeared, ref_bod) and np.sum(smeared > 0) == smeared.size
        peak = environment.BondOrder(6, mode="oocd").compute(system, quats, neighbors=neighbors, reset=False).bond_order
        assert np.sum(peak > 0) == 1 and peak[0, 0] > 0
        assert np.sum(environment.BondOrder(6, mode="oocd").compute(system, rq, neighbors=neighbors).bond_order > 0) > 30


def test_correlation_function_api():
    """freud.density.CorrelationFunction (tests/test_density_correlation_function.py upstream): complex and real
    inputs, is_complex, reset=False accumulation, histogram properties, the zero-mean random field known answer."""
    from tests.golden.make_golden import correlation_inputs

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "correlation_function.npz"))
    box, n = Box.cube(12), 3000
    pts, q = random_points(box, n, 7), random_points(box, 700, 8)
    v, qv = correlation_inputs(n, 700, 3)
    cf = density.CorrelationFunction(40, 3.0)
    assert cf.default_query_args == dict(mode="ball", r_max=3.0)
    cf.compute((box, pts), v, query_points=q, query_values=qv)
    assert cf.is_complex and cf.correlation.dtype == np.complex128
    assert np.array_equal(cf.bin_counts, gold["cube_complex_counts"])
    np.testing.assert_allclose(cf.correlation, gold["cube_complex_corr"], rtol=1e-11, atol=1e-12)
    cf.compute((box, pts), v, query_points=q, query_values=qv, reset=False)  # a second identical frame: same average
    assert np.array_equal(cf.bin_counts, 2 * gold["cube_complex_counts"])
    np.testing.assert_allclose(cf.correlation, gold["cube_complex_corr"], rtol=1e-11, atol=1e-12)
    cf.compute((box, pts), v.real)  # reset, real values, the points against themselves
    assert not cf.is_complex and cf.correlation.dtype == np.float64
    assert np.array_equal(cf.bin_counts, gold["cube_real_counts"])
    np.testing.assert_allclose(cf.correlation, gold["cube_real_corr"].real, rtol=1e-11, atol=1e-12)
    assert cf.nbins == 40 and cf.bounds == (0.0, 3.0) and len(cf.bin_edges) == 41 and cf.box == box
    # a constant field correlates to its square in every occupied bin (upstream's test_constant_field idea)
    cf.compute((box, pts), np.full(n, 2.0))
    assert np.allclose(cf.correlation[cf.bin_counts > 0], 4.0)
    with pytest.raises(ValueError):
        density.CorrelationFunction(0, 3.0)
    with pytest.raises(ValueError):
        density.CorrelationFunction(10, -1.0)


def test_local_density_api():
    """freud.density.LocalDensity (tests/test_density_local_density.py:21-110 upstream): attribute access, the known
    ranges of the reference's own test, default query arguments, an explicit NeighborList, query points != points."""
    box, pts = data.make_random_system(10, 10000, seed=123)
    ld = density.LocalDensity(3, 1)
    assert ld.r_max == 3 and ld.diameter == 1 and ld.default_query_args == dict(mode="ball", r_max=3.5)
    for attr in ("density", "num_neighbors", "box"):
        with pytest.raises(AttributeError):
            getattr(ld, attr)
    ld.compute(locality.AABBQuery(box, pts), neighbors=dict(mode="ball", r_max=3.5, exclude_ii=True))
    assert ld.box == Box.cube(10)
    assert (np.fabs(ld.density - 10.0) < 1.5).all() and (np.fabs(ld.num_neighbors - 1130.973355292) < 200).all()
    ld.compute((box, pts))  # default arguments, the points against themselves
    assert (np.fabs(ld.density - 10.0) < 1.5).all()
    # against the reference: committed outputs (on-the-fly query: summation order) and a handed-in list (bit for bit)
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "local_density.npz"))
    box, n = Box.cube(10), 3000
    pts, q = random_points(box, n, 123), random_points(box, 500, 124)
    ld = density.LocalDensity(3.0, 1.0).compute((box, pts), query_points=q)
    np.testing.assert_allclose(ld.num_neighbors, gold["cube_3_1_query_num"], rtol=1e-5)
    np.testing.assert_allclose(ld.density, gold["cube_3_1_query_density"], rtol=1e-5)
    ld = density.LocalDensity(3.0, 1.0).compute((box, pts))
    np.testing.assert_allclose(ld.num_neighbors, gold["cube_3_1_self_num"], rtol=1e-5)
    nl = locality.AABBQuery(box, pts).query(q, dict(r_max=3.5)).toNeighborList()
    ld = density.LocalDensity(3.0, 1.0).compute((box, pts), query_points=q, neighbors=nl)
    assert np.array_equal(bits(ld.num_neighbors), bits(gold["cube_3_1_nlist_num"]))
    assert np.array_equal(bits(ld.density), bits(gold["cube_3_1_nlist_density"]))
    with pytest.raises(ValueError):
        density.LocalDensity(0, 1)
    with pytest.raises(ValueError):
        density.LocalDensity(1, -1)
    assert repr(ld) == "freud.density.LocalDensity(r_max=3.0, diameter=1.0)"


def test_cellquery_engine():
    """freud.locality.CellQuery (tests/test_locality_neighbor_query.py:697-811 upstream): ball queries in its own
    arithmetic, nearest-neighbour queries refused with the reference's message, r_max validated against the box."""
    box, n, r = BOXES["tri1"]
    pts, q = random_points(box, n, seed=61), random_points(box, 300, seed=62)
    cq = locality.CellQuery(box, pts)
    for sbd in (False, True):
        nl = cq.query(q, dict(r_max=r, r_min=0.5)).toNeighborList(sort_by_distance=sbd)
        assert_matches_oracle(nl, port.ball_nlist(port.GHOST, box, False, pts, q, r, 0.5, False, sbd), f"cellquery {sbd}")
    nl = cq.query(pts, dict(r_max=r, exclude_ii=True)).toNeighborList()
    assert_matches_oracle(nl, port.ball_nlist(port.GHOST, box, False, pts, pts, r, 0.0, True), "cellquery self")
    if ref.available():
        want = ref.Query("cell", box, pts).nlist(pts, r_max=r, exclude_ii=True)
        assert_matches_oracle(nl, want, "cellquery vs the reference")
    with pytest.raises(RuntimeError, match="CellQuery only supports"):  # raised when the lazy result is consumed
        list(cq.query(q, dict(mode="nearest", num_neighbors=3)))
    with pytest.raises(RuntimeError, match="CellQuery only supports"):
        cq.query(q, dict(num_neighbors=4)).toNeighborList()
    with pytest.raises(RuntimeError, match="too large"):
        cq.query(q, dict(r_max=0.6 * float(box.Lx))).toNeighborList()
    # an RDF over a CellQuery system uses the same arithmetic
    rdf = density.RDF(bins=40, r_max=r).compute(cq)
    assert np.array_equal(rdf.bin_counts, port.rdf_accumulate(port.GHOST, box, False, pts, pts, 40, r, 0.0, True))


def test_steinhardt_w6_known_answers():
    """PERFECT_FCC_W6 = -0.00262604 with wl (and wl + average) on the perfect FCC crystal, k = 12 and a ball query
    (tests/test_order_steinhardt.py:18, :168-214 upstream)."""
    box, pts = data.make_fcc_system(4)
    for neighbors in (dict(num_neighbors=12), dict(r_max=0.8)):
        for kw in (dict(wl=True), dict(wl=True, average=True)):
            st = order.Steinhardt(6, **kw).compute((box, pts), neighbors=neighbors)
            np.testing.assert_allclose(np.average(st.particle_order), -0.00262604, atol=1e-5)
            np.testing.assert_allclose(st.particle_order, st.particle_order[0], atol=1e-5)
            assert abs(st.order - (-0.00262604)) < 1e-5
    # the averaged q_6 of a perfect crystal is q_6 itself; `ql` follows `average` (Steinhardt.h:107-114)
    st = order.Steinhardt(6, average=True).compute((box, pts), neighbors=dict(num_neighbors=12))
    np.testing.assert_allclose(st.particle_order, 0.57452416, atol=1e-5)
    np.testing.assert_allclose(st.ql, 0.57452416, atol=1e-5)
    with pytest.raises(IndexError):  # Wigner3j.cc: "implemented for l <= 20"
        order.Steinhardt(21, wl=True).compute((box, pts), neighbors=dict(num_neighbors=12))


@pytest.mark.parametrize("tag", ["ave", "wl", "wln", "ave_wl", "ave_wln"])
def test_steinhardt_options_golden(tag):
    """Second-shell average, w_l and its normalisation against outputs of the reference itself
    (tests/golden/steinhardt_options.npz, made by tests/golden/make_golden.py)."""
    from tests.golden.make_golden import STEINHARDT_OPTIONS

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "steinhardt_options.npz"))
    box, pts = data.make_fcc_system(4, scale=1.2, sigma_noise=0.06, seed=7)
    for ls in ([6], [4, 6], [3, 10]):
        key = tag + "_" + "_".join(str(l) for l in ls)
        st = order.Steinhardt(ls, **STEINHARDT_OPTIONS[tag]).compute((box, pts), neighbors=dict(num_neighbors=12))
        po, ql, want_po = st.particle_order, st.ql, gold[f"{key}_particle_order"]
        # w_l is a sum of O(l^2) signed terms: the tolerance is relative to the scale of the row, not the element
        scale = np.abs(want_po).max(axis=0)
        np.testing.assert_allclose(po, want_po, rtol=1e-5, atol=2e-5 * float(scale.max()))
        np.testing.assert_allclose(ql, gold[f"{key}_ql"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(st.order, gold[f"{key}_order"], rtol=1e-4, atol=1e-7)
        if ref.available():
            want = ref.Steinhardt(ls, **STEINHARDT_OPTIONS[tag]).compute(ref.Query("raw", box, pts), num_neighbors=12,
                                                                         exclude_ii=True)
            np.testing.assert_allclose(po, want["particle_order"], rtol=1e-5, atol=2e-5 * float(scale.max()))


def test_steinhardt_nan_without_neighbours():
    """tests/test_order_steinhardt.py:361-369 upstream."""
    box = Box.cube(10)
    pts = np.array([[0, 0, 0], [0.5, 0, 0], [4, 4, 4]], np.float32)
    ql = order.Steinhardt(6).compute((box, pts), neighbors=dict(r_max=1.0)).ql
    assert np.isfinite(ql[0]) and np.isfinite(ql[1]) and np.isnan(ql[2])


@pytest.mark.gpu
def test_query_single_matches_the_rows_of_the_full_query():
    """NeighborQuery::querySingle / NeighborQueryPerPointIterator (NeighborQuery.h:144-154, 309-350): the per-point
    iterator that loopOverNeighborsIterator hands a compute yields exactly row i of the full query, in both modes."""
    from freud_b200 import data, locality

    box, pts = data.make_random_system(12.0, 400, seed=9)
    L = locality._ext()._locality
    for cls in (locality.AABBQuery, locality.LinkCell):
        nq = cls(box, pts)
        qa = locality._query_args(dict(mode="ball", r_max=2.5, exclude_ii=True))
        full = nq.query(pts, dict(mode="ball", r_max=2.5, exclude_ii=True)).toNeighborList()
        for i in (0, 17, 399):
            row = nq._cpp_obj.querySingle([float(v) for v in pts[i]], i, qa)
            sel = full.query_point_indices == i
            assert [b[1] for b in row] == list(full.point_indices[sel])
            assert np.array_equal(np.float32([b[2] for b in row]), full.distances[sel])
            assert all(b[0] == i for b in row)
        qk = locality._query_args(dict(mode="nearest", num_neighbors=5, exclude_ii=True))
        row = nq._cpp_obj.querySingle([float(v) for v in pts[3]], 3, qk)
        fullk = nq.query(pts, dict(num_neighbors=5, exclude_ii=True)).toNeighborList(sort_by_distance=True)
        sel = fullk.query_point_indices == 3
        assert [b[1] for b in row] == list(fullk.point_indices[sel])
    del L
