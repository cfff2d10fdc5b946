"""Host-side contract of the reference-facing API (no GPU needed): argument validation, error types, bin edges,
NeighborList container semantics.  Each case restates a test of the reference's own suite (cited)."""
import numpy as np
import pytest

from freud_b200 import density, locality, order
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


def test_no_cpu_fallback():
    """Without a CUDA device every compute call fails loudly (FGPU_ECUDA -> RuntimeError)."""
    from freud_b200 import _capi

    if _capi.lib().fgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        density.RDF(10, 2.0).compute((Box.cube(10), np.zeros((5, 3), np.float32)))
    with pytest.raises(RuntimeError):
        locality.LinkCell(Box.cube(10), np.zeros((5, 3), np.float32)).query(np.zeros((1, 3)), dict(r_max=2)).toNeighborList()
