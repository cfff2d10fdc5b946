"""Parity at the sizes BASELINE.json states (north_star: "bit-exact LinkCell NeighborList and RDF bin counts at N=1M"):
configs[1] (C2), configs[2] (C3), configs[3] (C4) and configs[4] (C5) literally, CUDA path against the oracle.

The oracle's grid-based restatement (oracle/port.c, pinned to the compiled reference bit for bit up to N = 1e5 in
tests/test_oracle_port.py) finishes these sizes in seconds; C3's k-nearest-neighbour list and q_l come from the
compiled reference itself (oracle/_ref).  Inputs follow SURVEY.md section 8(d).  These sizes also exercise what the
small parity tests cannot: the stage-1 filter slack E ~ (Lx + Ly + Lz) at L = 232 / 368 / 1414 and the tile / hit
buffer sizing at the benchmark's cell populations.
"""
import numpy as np
import pytest

from oracle import port, ref
from tests.util import assert_nlist_equal, bits

pytestmark = pytest.mark.gpu
WRAP, IMAGE = 0, 1


@pytest.fixture(scope="module")
def gctx():
    from freud_b200 import _capi

    c = _capi.Context(0)
    yield c
    c.close()


def test_c2_linkcell_nlist_1m(gctx):
    """configs[1]: LinkCell r_max = 3, exclude_ii, 1 M random points, cubic box at rho = 0.08: all five arrays +
    segments / counts bitwise (freud/locality/LinkCell.cc:496-573, NeighborQuery.h:434-481)."""
    from freud_b200 import _capi, data

    n = 1_000_000
    box, pts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=0)
    dp = _capi.DevicePoints(gctx, box, pts)
    got = dp.ball_query(None, WRAP, 3.0, 0.0, True).to_host()
    want = port.ball_nlist(port.WRAP, box, False, pts, pts, 3.0, 0.0, True)
    assert len(want) > 9_000_000
    assert_nlist_equal(got, want, "C2 1M wrap")
    # the same frame in the AABBQuery arithmetic (what (box, points) systems get), sorted by distance
    got = dp.ball_query(None, IMAGE, 3.0, 0.0, True, sort_by_distance=True).to_host()
    want = port.ball_nlist(port.IMAGE, box, False, pts, pts, 3.0, 0.0, True, True)
    assert_nlist_equal(got, want, "C2 1M image by distance")


def test_rdf_1m_r5(gctx):
    """BASELINE metric "RDF frames/sec @1M particles r_max=5": raw bin counts bitwise, both arithmetics
    (freud/density/RDF.cc:101-110, freud/util/Histogram.h:152-174)."""
    from freud_b200 import _capi, data

    n = 1_000_000
    box, pts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=0)
    dp = _capi.DevicePoints(gctx, box, pts)
    for flavour, pf in ((IMAGE, port.IMAGE), (WRAP, port.WRAP)):
        rdf = _capi.DeviceRDF(gctx, 100, 5.0)
        rdf.accumulate(dp, None, flavour, 5.0, 0.0, True)
        want = port.rdf_accumulate(pf, box, False, pts, pts, 100, 5.0, 0.0, True)
        assert int(want.astype(np.uint64).sum()) > 41_000_000
        assert np.array_equal(rdf.read(), want), f"flavour {flavour}"


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_c3_steinhardt_q6_fcc_1m(gctx):
    """configs[2]: Q6, 12 nearest neighbours, 63^3 x 4 = 1 000 188-particle FCC lattice with sigma = 0.05 noise: the kNN
    NeighborList bitwise against the reference's AABBQuery (AABBQuery.cc:152-281), q_l within 1e-5 relative and q_lm
    within 1e-5 absolute of the reference's Steinhardt::compute (Steinhardt.cc:85-222)."""
    from freud_b200 import _capi, data

    box, pts = data.make_fcc_system(63, sigma_noise=0.05, seed=0)
    assert len(pts) == 1_000_188
    dp = _capi.DevicePoints(gctx, box, pts)
    nl = dp.knn_query(None, 12, exclude_ii=True)
    got = dp.steinhardt(nl, [6], want_qlm=True)
    ref.set_num_threads(0)
    q = ref.Query("aabb", box, pts)
    want_nl = q.nlist(pts, mode="nearest", num_neighbors=12, exclude_ii=True)
    assert_nlist_equal(nl.to_host(), want_nl, "C3 kNN")
    want = ref.Steinhardt(6).compute(q, nlist=want_nl)
    assert np.allclose(got["ql"], want["ql"], rtol=1e-5, atol=0)
    assert np.allclose(got["qlm"][0], want["qlm"][0], atol=1e-5)
    assert np.allclose(got["order"], want["order"], rtol=1e-4)


def test_c4_rdf_4m_triclinic_sharded(gctx):
    """configs[3]: RDF bins = 500, r_max = 5, 4 M points, triclinic box: counts bitwise on one GPU and as the sum of 8
    home-tile shards (one process plays the ranks; the slab-restricted cell list of every shard is exercised)."""
    from freud_b200 import _capi, data

    n = 4_000_000
    box, pts = data.make_random_system((n / 0.08) ** (1 / 3), n, seed=0, tilt=(0.3, 0.2, 0.1))
    want = port.rdf_accumulate(port.IMAGE, box, False, pts, pts, 500, 5.0, 0.0, True)
    assert int(want.astype(np.uint64).sum()) > 167_000_000
    dp = _capi.DevicePoints(gctx, box, pts)
    rdf = _capi.DeviceRDF(gctx, 500, 5.0)
    rdf.accumulate(dp, None, IMAGE, 5.0, 0.0, True)
    assert np.array_equal(rdf.read(), want), "single GPU"
    total = np.zeros(500, np.uint64)
    for shard in range(8):
        dp.set_shard(shard, 8)
        rdf.reset()
        rdf.accumulate(dp, None, IMAGE, 5.0, 0.0, True)
        total += rdf.read()
    dp.set_shard(0, 1)
    assert np.array_equal(total.astype(np.uint32), want), "sum of 8 shards"


def test_c5_trajectory_rdf_2d_1m(gctx):
    """configs[4]: 2-D box, 1 M points per frame at areal density 0.5, RDF bins = 100 r_max = 5 accumulated with
    reset=False (4 of the 64 frames here; bench.py checks every rank's 8)."""
    from freud_b200 import _capi, data

    n = 1_000_000
    L = (n / 0.5) ** 0.5
    rdf = _capi.DeviceRDF(gctx, 100, 5.0)
    want = np.zeros(100, np.uint32)
    for seed in range(4):
        box, pts = data.make_random_system(L, n, is2D=True, seed=seed)
        rdf.accumulate(_capi.DevicePoints(gctx, box, pts), None, IMAGE, 5.0, 0.0, True)
        want = port.rdf_accumulate(port.IMAGE, box, True, pts, pts, 100, 5.0, 0.0, True, counts=want)
    assert np.array_equal(rdf.read(), want)
    # the same frames as NeighborList slices: 2-D bond vectors keep z = 0 bit for bit at L = 1414
    box, pts = data.make_random_system(L, n, is2D=True, seed=0)
    got = _capi.DevicePoints(gctx, box, pts).ball_query(None, IMAGE, 5.0, 0.0, True).to_host()
    want_nl = port.ball_nlist(port.IMAGE, box, True, pts, pts, 5.0, 0.0, True)
    assert_nlist_equal(got, want_nl, "C5 frame 0 NeighborList")
