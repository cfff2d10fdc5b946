"""Host-side contract of the reference-facing API (no GPU needed): argument validation, error types, bin edges,
NeighborList container semantics.  Each case restates a test of the reference's own suite (cited)."""
import numpy as np
import pytest

from freud_b200 import density, environment, locality, order, pmft
from freud_b200.box import Box


def test_rdf_constructor_validation():
    """tests/test_density_rdf.py:69-80 upstream."""
    for args in ((0, 5.0), (10, -1.0), (10, 5.0, -0.5), (10, 2.0, 3.0)):
        with pytest.raises(ValueError):
            density.RDF(*args)
    with pytest.raises(ValueError):
        density.RDF(10, 2.0, normalization_mode="bogus")


def test_rdf_bin_edges_centers_bounds():
    """tests/test_density_rdf.py:242-251 upstream: edges/centres of a RegularAxis, atol 1e-6."""
    for r_min in (0.0, 0.1, 3.0):
        rdf = density.RDF(bins=10, r_max=5.16, r_min=r_min)
        edges = np.linspace(r_min, 5.16, 11, dtype=np.float32)
        np.testing.assert_allclose(rdf.bin_edges, edges, atol=1e-6)
        np.testing.assert_allclose(rdf.bin_centers, (edges[:-1] + edges[1:]) / 2, atol=1e-6)
        assert rdf.bounds == pytest.approx((r_min, 5.16)) and rdf.nbins == 10
    # float32 recipe of Histogram.h:126-138: min + float(i) * width, one rounding per operation
    rdf = density.RDF(100, 5.0)
    w = np.float32(np.float32(5.0) / np.float32(100))
    assert np.array_equal(rdf.bin_edges, np.float32(0) + np.arange(101, dtype=np.float32) * w)


def test_query_argument_errors():
    """tests/test_locality_neighbor_query.py:57-78, :199-216, :584-599 upstream: error TYPES."""
    box = Box.cube(10)
    pts = np.zeros((4, 3), np.float32)
    with pytest.raises(ValueError):  # zero particles
        locality.AABBQuery(box, np.zeros((0, 3), np.float32))
    with pytest.raises(ValueError):  # ill-shaped input
        locality.AABBQuery(box, np.zeros((4, 2), np.float32))
    with pytest.raises(ValueError):  # 2-D box with z != 0 (:560-568)
        locality.LinkCell(Box.square(10), np.array([[0, 0, 0.1]], np.float32))
    with pytest.raises(RuntimeError):  # cell_width larger than half the box (LinkCell.cc:241-246)
        locality.LinkCell(box, pts, cell_width=6.0)
    nq = locality.AABBQuery(box, pts)
    for bad in (dict(r_max=0.0), dict(r_max=1.0, r_min=2.0), dict(r_max=1.0, r_min=1.0)):
        with pytest.raises(ValueError):  # NeighborQuery.h:321-328 -> invalid_argument
            nq.query(pts, bad).toNeighborList()
    for bad in (dict(mode="ball"), dict(mode="ball", r_max=1.0, num_neighbors=3), dict(mode="nearest"),
                dict(mode="nearest", num_neighbors=3, scale=0.9), dict()):
        with pytest.raises(RuntimeError):  # NeighborQuery.h:195-228 -> runtime_error
            nq.query(pts, bad).toNeighborList()
    with pytest.raises(ValueError):
        nq.query(pts, dict(r_max=1.0, bogus=1))
    with pytest.raises(ValueError):
        nq.query(pts, dict(mode="sideways", r_max=1.0))
    with pytest.raises(ValueError):
        nq.query(np.zeros((3, 2)), dict(r_max=1.0))


def test_neighborlist_from_arrays_and_container():
    """tests/test_locality_neighbor_list.py:136-218, :298-320 upstream."""
    qi = np.array([0, 0, 1, 2, 3], np.uint32)
    pj = np.array([1, 2, 3, 0, 0], np.uint32)
    vec = np.array([[1, 0, 0], [0, 2, 0], [0, 0, 3], [1, 1, 0], [0.5, 0, 0]], np.float32)
    nl = locality.NeighborList.from_arrays(4, 4, qi, pj, vec)
    assert len(nl) == 5 and nl.num_query_points == 4 and nl.num_points == 4
    assert np.array_equal(nl.query_point_indices, qi) and np.array_equal(nl.point_indices, pj)
    assert np.array_equal(nl.distances, np.sqrt((vec * vec).sum(1)).astype(np.float32))
    assert np.array_equal(nl.weights, np.ones(5, np.float32))
    assert np.array_equal(nl.neighbor_counts, [2, 1, 1, 1]) and np.array_equal(nl.segments, [0, 2, 3, 4])
    assert nl.find_first_index(2) == 3 and nl.find_first_index(5) == 5
    # empty rows keep segment 0 (NeighborList.cc:199-232)
    nl2 = locality.NeighborList.from_arrays(6, 4, np.array([1, 4], np.uint32), np.array([0, 1], np.uint32), vec[:2])
    assert np.array_equal(nl2.neighbor_counts, [0, 1, 0, 0, 1, 0]) and np.array_equal(nl2.segments, [0, 0, 0, 0, 1, 0])
    # arrays are read-only views (:26-39, :276-287)
    for arr in (nl.distances, nl.weights, nl.vectors, nl.segments, nl.neighbor_counts, nl[:]):
        with pytest.raises(ValueError):
            arr[0] = 0
    # validation (:141-218)
    with pytest.raises(ValueError):
        locality.NeighborList.from_arrays(4, 4, qi[::-1].copy(), pj, vec)  # unsorted
    with pytest.raises(ValueError):
        locality.NeighborList.from_arrays(3, 4, qi, pj, vec)  # query index out of range
    with pytest.raises(ValueError):
        locality.NeighborList.from_arrays(4, 3, qi, np.array([1, 2, 3, 0, 0], np.uint32), vec)  # point index out of range
    with pytest.raises(ValueError):
        locality.NeighborList.from_arrays(4, 4, qi, pj[:4], vec)  # length mismatch
    # sort by distance, filter, copy (:298-320, :85-134)
    by_d = nl.copy().sort(by_distance=True)
    assert np.array_equal(by_d.point_indices, [1, 2, 3, 0, 0]) and np.array_equal(by_d.distances[:2], [1, 2])
    kept = nl.copy().filter(nl.distances < 1.5)
    assert np.array_equal(kept.point_indices, [1, 0, 0]) and np.array_equal(kept.neighbor_counts, [1, 0, 1, 1])
    assert len(nl.copy().filter_r(2.5, 0.9)) == 3
    assert len(nl) == 5  # the copies did not touch the original


def test_steinhardt_constructor():
    """Constructor flags are kept as given (freud/order.py:489-520); nothing touches the GPU before compute()."""
    assert order.Steinhardt(6).l == 6 and order.Steinhardt([4, 6]).l == [4, 6]
    st = order.Steinhardt(6, average=True, wl=True, weighted=True, wl_normalize=True)
    assert st.average and st.wl and st.weighted and st.wl_normalize
    st = order.Steinhardt(6)
    assert not (st.average or st.wl or st.weighted or st.wl_normalize)
    with pytest.raises(ValueError):
        order.Steinhardt(-1)
    with pytest.raises(NotImplementedError):  # no default query arguments, as upstream
        order.Steinhardt(6).compute((Box.cube(5), np.zeros((3, 3), np.float32)))


def test_histogram_clients_host_side():
    """Constructors, axes and errors of the PMFT family and BondOrder need no GPU (freud/pmft.py:124-590,
    freud/environment.py:204-391; RegularAxis edges and centres as freud/util/Histogram.h:87-138)."""
    f32 = np.float32
    xyz = pmft.PMFTXYZ(1.0, 2.0, 3.0, (4, 5, 6), shiftvec=[0.5, 0, 0])
    assert xyz.nbins == (4, 5, 6) and xyz.bounds == [(-1.0, 1.0), (-2.0, 2.0), (-3.0, 3.0)]
    assert np.isclose(xyz.r_max, np.sqrt(14.0)) and xyz.default_query_args == dict(mode="ball", r_max=xyz.r_max)
    assert np.array_equal(xyz.bin_edges[0], f32([-1, -0.5, 0, 0.5, 1])) and np.array_equal(xyz.shiftvec, f32([0.5, 0, 0]))
    assert np.array_equal(xyz.bin_centers[0], f32([-0.75, -0.25, 0.25, 0.75]))
    for prop in ("bin_counts", "pmft", "box"):  # _Compute._computed_property, tests/test_pmft.py:93-105 upstream
        with pytest.raises(AttributeError):
            getattr(xyz, prop)
    xyt = pmft.PMFTXYT(2.0, 1.0, 8)
    two_pi = float(f32(2 * np.pi))
    assert xyt.nbins == (8, 8, 8) and xyt.bounds[2] == (0.0, two_pi) and np.isclose(xyt.r_max, np.sqrt(5.0))
    width = f32(two_pi) / f32(8)
    assert np.array_equal(xyt.bin_edges[2], f32(0) + np.arange(9, dtype=f32) * width)
    r12 = pmft.PMFTR12(3.0, (3, 4, 5))
    assert r12.nbins == (3, 4, 5) and r12.bounds == [(0.0, 3.0), (0.0, two_pi), (0.0, two_pi)] and r12.r_max == 3.0
    assert "PMFTR12(r_max=3.0, bins=(3, 4, 5))" in repr(r12) and "shiftvec=[0.5, 0.0, 0.0]" in repr(xyz)
    for make in (lambda: pmft.PMFTXYZ(1, 1, 1, (0, 2, 2)), lambda: pmft.PMFTXYZ(1, -1, 1, 2), lambda: pmft.PMFTXYT(1, 1, (2, 2, 0)),
                 lambda: pmft.PMFTXYT(-1, 1, 2), lambda: pmft.PMFTR12(-1.0, 2), lambda: pmft.PMFTR12(1.0, (2, 0, 2)),
                 lambda: pmft.PMFTXY(1, 1, (0, 2))):
        with pytest.raises(ValueError):
            make()
    bo = environment.BondOrder((6, 3), mode="lbod")
    pi = float(f32(np.pi))
    assert bo.nbins == (6, 3) and bo.mode == "lbod" and bo.bounds == [(0.0, two_pi), (0.0, pi)]
    assert [len(e) for e in bo.bin_edges] == [7, 4] and "mode='lbod'" in repr(bo)
    for prop in ("bond_order", "bin_counts", "box"):
        with pytest.raises(AttributeError):
            getattr(bo, prop)
    with pytest.raises(NotImplementedError):
        bo.default_query_args
    for make in (lambda: environment.BondOrder((1, 3)), lambda: environment.BondOrder((3, 1)),
                 lambda: environment.BondOrder(4, mode="other")):
        with pytest.raises(ValueError):
            make()
    # argument checks that come before any device work
    box, pts = Box.cube(10), np.zeros((6, 3), np.float32)
    with pytest.raises(ValueError):
        pmft.PMFTXYZ(1, 1, 1, 2).compute((box, pts), np.zeros((5, 4)))  # one quaternion per query point
    with pytest.raises(ValueError):
        pmft.PMFTXYZ(1, 1, 1, 2).compute((box, pts), np.zeros((6, 4)), equiv_orientations=np.zeros((2, 3)))
    with pytest.raises(ValueError):
        pmft.PMFTR12(1.0, 2).compute((Box.square(10), pts), np.zeros(5))
    with pytest.raises(ValueError):
        bo.compute((box, pts), np.zeros((6, 3)), neighbors=dict(num_neighbors=2))


def test_results_need_compute_first():
    """``_Compute._computed_property`` (freud/util.py:61-80; tests/test_density_rdf.py:42-59 and its siblings upstream):
    every result raises AttributeError until compute() ran; the histogram geometry does not."""
    cases = [(density.RDF(10, 2.0), ("rdf", "n_r", "bin_counts", "box"), ("bin_edges", "bin_centers", "bounds", "nbins")),
             (density.CorrelationFunction(10, 2.0), ("correlation", "bin_counts", "box"), ("bin_edges", "bounds", "nbins")),
             (density.LocalDensity(2.0, 1.0), ("density", "num_neighbors", "box"), ("r_max", "diameter")),
             (order.Steinhardt(6), ("order", "particle_order", "ql", "particle_harmonics"), ("l", "average", "wl")),
             (pmft.PMFTXY(1, 1, 4), ("pmft", "bin_counts", "box"), ("bin_edges", "bin_centers", "bounds", "nbins"))]
    for obj, results, geometry in cases:
        for name in results:
            with pytest.raises(AttributeError):
                getattr(obj, name)
        for name in geometry:
            getattr(obj, name)


def test_no_cpu_fallback():
    """Without a CUDA device every compute call fails loudly (FGPU_ECUDA -> RuntimeError)."""
    from freud_b200 import _capi

    if _capi.lib().fgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        density.RDF(10, 2.0).compute((Box.cube(10), np.zeros((5, 3), np.float32)))
    with pytest.raises(RuntimeError):
        locality.LinkCell(Box.cube(10), np.zeros((5, 3), np.float32)).query(np.zeros((1, 3)), dict(r_max=2)).toNeighborList()


# ---- system adapters (freud/locality.py:268-383, freud/box.py:751-882) -------------
def _foreign(path, **attrs):
    """Stand-in for a frame of a reader that is not installed: from_system recognises them by class path."""
    module, name = path.rsplit(".", 1)
    return type(name, (), {"__module__": module})(), attrs


def _ns(**kw):
    return type("NS", (), kw)()


def _make(path, **attrs):
    obj, attrs = _foreign(path, **attrs)
    for k, v in attrs.items():
        setattr(obj, k, v)
    return obj


def test_from_box_forms():
    b = Box.from_box([2, 3, 4, 0.1, 0.2, 0.3])
    assert (b.Lx, b.Ly, b.Lz, b.dimensions) == (2, 3, 4, 3) and np.allclose([b.xy, b.xz, b.yz], [0.1, 0.2, 0.3])
    assert Box.from_box([2, 3]).is2D and Box.from_box([2, 3, 0]).is2D and not Box.from_box([2, 3, 4]).is2D
    assert Box.from_box({"Lx": 2, "Ly": 3, "Lz": 4}).Lz == 4
    assert Box.from_box({"Lx": 2, "Ly": 3, "Lz": 1, "dimensions": 2}).is2D
    assert Box.from_box(_ns(Lx=2, Ly=3, Lz=5, xy=0.5)).xy == 0.5
    assert Box.from_box(b) is b
    with pytest.raises(ValueError):
        Box.from_box([1, 2, 3, 4])
    with pytest.raises(ValueError):
        Box.from_box(_ns(Lx=2, Ly=3, Lz=5, dimensions=3), dimensions=2)
    with pytest.raises(ValueError):
        Box.from_box({"Lx": 2, "Ly": 3, "Lz": 5, "dimensions": 3}, dimensions=2)
    # 3x3 matrix of lattice vectors (columns), round trip through to_matrix
    tri = Box(4, 5, 6, 0.3, -0.2, 0.1)
    back = Box.from_box(tri.to_matrix())
    assert np.allclose(back.as_array6(), tri.as_array6(), atol=1e-6) and not back.is2D
    flat = Box.from_matrix([[4, 1, 0], [0, 5, 0], [0, 0, 0]])
    assert flat.is2D and np.isclose(flat.xy, 0.2)
    ang = Box.from_box_lengths_and_angles(2, 3, 4, np.pi / 2, np.pi / 2, np.pi / 2)
    assert np.allclose(ang.as_array6(), [2, 3, 4, 0, 0, 0], atol=1e-6)
    with pytest.raises(ValueError):
        Box.from_box_lengths_and_angles(2, 3, 4, 0, 1, 1)


def test_from_system_adapters():
    pts = np.array([[0, 0, 0], [1, 1, 0], [-1, 2, 0]], dtype=np.float32)
    # pairs, attribute objects, existing engines
    nq = locality.NeighborQuery.from_system((Box.cube(10), pts))
    assert type(nq).__name__ == "_RawPoints" and np.array_equal(nq.points, pts) and nq.box == Box.cube(10)
    assert locality.NeighborQuery.from_system(nq) is nq
    duck = locality.NeighborQuery.from_system(_ns(box=[10, 10, 10], points=pts))
    assert duck.box == Box.cube(10)
    aq = locality.AABBQuery.from_system((Box.cube(10), pts))
    assert isinstance(aq, locality.AABBQuery) and locality.AABBQuery.from_system(aq) is aq
    many = np.random.default_rng(0).uniform(-5, 5, (500, 3)).astype(np.float32)
    lc = locality.LinkCell.from_system(locality.AABBQuery(Box.cube(10), many))  # another engine over the same data
    assert isinstance(lc, locality.LinkCell) and np.array_equal(lc.points, many)
    # MDAnalysis Timestep (both class paths)
    for path in ("MDAnalysis.coordinates.base.Timestep", "MDAnalysis.coordinates.timestep.Timestep"):
        ts = _make(path, triclinic_dimensions=np.diag([10.0, 11.0, 12.0]), positions=pts)
        got = locality.NeighborQuery.from_system(ts)
        assert np.allclose(got.box.L, [10, 11, 12]) and np.array_equal(got.points, pts)
    # GSD / HOOMD-blue 3: Lz = 1 with dimensions = 2 is a 2-D box
    for path in ("gsd.hoomd.Frame", "gsd.hoomd.Snapshot", "hoomd.snapshot.Snapshot"):
        frame = _make(path, configuration=_ns(box=[10, 10, 1, 0.5, 0, 0], dimensions=2), particles=_ns(position=pts))
        got = locality.NeighborQuery.from_system(frame)
        assert got.box.is2D and got.box.Lz == 0 and got.box.xy == 0.5
        frame3 = _make(path, configuration=_ns(box=[10, 10, 8, 0, 0, 0], dimensions=3), particles=_ns(position=pts))
        assert locality.NeighborQuery.from_system(frame3).box.Lz == 8
    # garnett: position (>= 0.7) or positions
    g_new = _make("garnett.trajectory.Frame", box=_ns(Lx=9, Ly=9, Lz=9), position=pts)
    g_old = _make("garnett.trajectory.Frame", box=_ns(Lx=9, Ly=9, Lz=9), positions=pts)
    assert locality.NeighborQuery.from_system(g_new).box == locality.NeighborQuery.from_system(g_old).box == Box.cube(9)
    # OVITO: 3x4 cell matrix (last column is the origin)
    cell = _ns(matrix=np.hstack([np.diag([7.0, 8.0, 9.0]), np.zeros((3, 1))]), is2D=False)
    ov = _make("ovito.data.DataCollection", cell=cell, particles=_ns(positions=pts))
    assert np.allclose(locality.NeighborQuery.from_system(ov).box.L, [7, 8, 9])
    # HOOMD-blue 2 snapshot: box + particles.position
    h2 = _ns(box=_ns(Lx=6, Ly=6, Lz=1, xy=0.25, dimensions=2), particles=_ns(position=pts))
    got = locality.NeighborQuery.from_system(h2)
    assert got.box.is2D and got.box.xy == 0.25
    with pytest.raises(ValueError):
        locality.NeighborQuery.from_system(42)


def test_cellquery_grid_introspection_and_all_pairs():
    """The host-only corners the reference's binding layer exposes (export-NeighborQuery.cc:96-111,
    export-NeighborList.cc:41-51): CellQuery's grid description and NeighborList.all_pairs, the latter's vectors against
    the compiled reference's Box::wrap bit for bit."""
    import numpy as np

    from freud_b200 import locality
    from freud_b200.box import Box
    from oracle import ref

    box = Box(10, 12, 14, 0.2, -0.1, 0.3)
    rs = np.random.RandomState(3)
    pts = box.make_absolute(rs.random_sample((200, 3))).astype(np.float32)
    cq = locality.CellQuery(box, pts)._cpp_obj
    cq.buildGrid(2.0)
    counts, real, starts = np.array(cq.getCounts()), np.array(cq.getCountsReal()), np.array(cq.getCellStarts())
    assert cq.getNx() == int((10 + 12 * 0.2 + 14 * 0.1) / 2.0) + 3 and cq.getNz() == 14 // 2 + 3
    assert len(counts) == cq.getNx() * cq.getNy() * cq.getNz()
    assert real.sum() == 200 and counts.sum() == cq.getNTotal() > 200  # ghosts of the points near a face
    assert np.array_equal(starts, np.concatenate([[0], np.cumsum(counts)[:-1]]))
    assert abs(cq.getCellWidth() - 2.0) < 1e-6 and abs(cq.getCellInverseWidth() - 0.5) < 1e-7
    with pytest.raises(RuntimeError):
        cq.buildGrid(-1.0)

    q = pts[:7] + np.float32(0.25)
    nl = locality.NeighborList.all_pairs((box, pts[:50]), q, exclude_ii=True)
    assert len(nl) == 50 * 7 - 7
    i, j = nl.query_point_indices.astype(int), nl.point_indices.astype(int)
    assert not np.any(i == j) and np.all(np.diff(i) >= 0)
    if ref.available():
        want = ref.box_apply(box, False, "wrap", q[i] - pts[:50][j])
        assert np.array_equal(nl.vectors.view(np.uint32), want.view(np.uint32))
    assert np.allclose(nl.distances, np.linalg.norm(nl.vectors, axis=1), rtol=1e-6)
    assert np.all(nl.weights == 1)


def test_pmft_quaternion_orientations_must_rotate_about_z():
    """freud/pmft.py:58-84 (`_quat_to_z_angle`): quaternions are accepted only as rotations about +z, a 1-D length-4
    input is a quaternion unless there are exactly four points, and the angle is rowan's 2 atan2(|v|, w)."""
    import numpy as np

    from freud_b200.pmft import _angles

    half = np.pi / 6
    qz = np.tile([np.cos(half), 0, 0, np.sin(half)], (5, 1))
    assert np.allclose(_angles(qz, 5), 2 * half)
    assert np.allclose(_angles(np.tile([1.0, 0, 0, 0], (5, 1)), 5), 0)  # identities are fine
    with pytest.raises(ValueError):
        _angles(np.tile([np.cos(half), np.sin(half), 0, 0], (5, 1)), 5)  # about x
    with pytest.raises(ValueError):
        _angles(np.tile([np.cos(half), 0, 0, -np.sin(half)], (5, 1)), 5)  # about -z: rejected upstream too
    assert np.allclose(_angles([0.1, 0.2, 0.3, 0.4], 4), [0.1, 0.2, 0.3, 0.4])  # four points: four angles
    assert np.allclose(_angles(np.array([np.cos(half), 0, 0, np.sin(half)]), 1), 2 * half)  # one point, one quaternion
    assert np.allclose(_angles(np.arange(5) * 0.1, 5), np.arange(5) * 0.1)
