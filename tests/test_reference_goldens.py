"""The golden vectors the reference's own test-suite holds for this path:

* tests/integration/files/lj/rdf.npz -- g(r), 100 bins, r_max = 5, accumulated with reset=False over the ten frames of
  the Lennard-Jones trajectory lj.gsd / lj.dcd (tests/integration/test_reader_integrations.py:27-60, rtol = atol = 1e-5);
* tests/validation/files/steinhardt_average/GC_rc1.4_*.txt -- q4, q6, w4, w6 (and their second-shell averages) of the
  3288 particles of Test_Configuration.gsd from an independent code, ball neighbours r_max = 1.4
  (tests/validation/test_steinhardt_average.py:82-118, atol 2e-5).

The data files are copied verbatim under tests/golden/reference_files/ (test data, not code); they are parsed by
tests/readers.py because neither `gsd` nor `MDAnalysis` is installed here.  The CPU tests pin the oracle (the compiled
reference) to these vectors, the GPU tests pin the product to them through its freud-style Python API.  The
Voronoi-weighted variants (RvD_MSM_*) need freud.locality.Voronoi, which is outside this path.
"""
import os

import numpy as np
import pytest

from oracle import ref
from tests import readers

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_files")
LJ_RDF = np.load(os.path.join(HERE, "rdf.npz"))["rdf"]
needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")


def _gc(name):
    return np.genfromtxt(os.path.join(HERE, f"{name}.txt"))


def test_readers_agree():
    """The two containers of the LJ trajectory hold the same ten frames."""
    dcd = readers.read_dcd(os.path.join(HERE, "lj.dcd"))
    gsd = readers.read_gsd(os.path.join(HERE, "lj.gsd"))
    assert len(dcd) == len(gsd) == 10
    for (bd, pd), (bg, pg) in zip(dcd, gsd):
        assert (bd.Lx, bd.Ly, bd.Lz) == (bg.Lx, bg.Ly, bg.Lz) == (20.0, 20.0, 20.0)
        assert pd.shape == pg.shape == (1000, 3) and np.array_equal(pd, pg)
    conf = readers.read_gsd(os.path.join(HERE, "Test_Configuration.gsd"))
    assert len(conf) == 1 and conf[0][1].shape == (3288, 3)


@needs_ref
@pytest.mark.parametrize("container", ["dcd", "gsd"])
def test_oracle_reproduces_lj_rdf(container):
    frames = (readers.read_dcd if container == "dcd" else readers.read_gsd)(os.path.join(HERE, f"lj.{container}"))
    R = ref.RDF(100, 5.0)
    for box, pts in frames:
        R.accumulate(ref.Query("raw", box, pts), pts, mode="ball", r_max=5.0, exclude_ii=True)
    assert np.allclose(R.results()["rdf"], LJ_RDF, rtol=1e-5, atol=1e-5)


@needs_ref
@pytest.mark.parametrize("average", [False, True])
def test_oracle_reproduces_gc_steinhardt(average):
    box, pts = readers.read_gsd(os.path.join(HERE, "Test_Configuration.gsd"))[-1]
    q = ref.Query("aabb", box, pts)
    nl = q.nlist(pts, mode="ball", r_max=1.4, exclude_ii=True)
    want = _gc("GC_rc1.4_avq4avq6avw4avw6" if average else "GC_rc1.4_q4q6w4w6")
    cols = [ref.Steinhardt(4, average=average).compute(q, nlist=nl)["particle_order"][:, 0],
            ref.Steinhardt(6, average=average).compute(q, nlist=nl)["particle_order"][:, 0],
            ref.Steinhardt(4, average=average, wl=True, wl_normalize=True).compute(q, nlist=nl)["particle_order"][:, 0],
            ref.Steinhardt(6, average=average, wl=True, wl_normalize=True).compute(q, nlist=nl)["particle_order"][:, 0]]
    for k, name in enumerate(("q4", "q6", "w4", "w6")):
        assert np.allclose(want[:, k], cols[k], atol=2e-5), name


@pytest.mark.gpu
@pytest.mark.parametrize("container", ["dcd", "gsd"])
def test_lj_trajectory_rdf_matches_reference_golden(container):
    """tests/integration/test_reader_integrations.py:32-40, run_analyses: RDF(100, 5).compute(system, reset=False) over the
    trajectory, then rdf against the stored g(r); Steinhardt(6) with six nearest neighbours runs on every frame too."""
    from freud_b200 import density, order

    frames = (readers.read_dcd if container == "dcd" else readers.read_gsd)(os.path.join(HERE, f"lj.{container}"))
    rdf = density.RDF(bins=100, r_max=5)
    ql = order.Steinhardt(6)
    for system in frames:
        rdf.compute(system, reset=False)
        ql.compute(system, neighbors={"num_neighbors": 6})
        assert np.all(np.isfinite(ql.particle_order))
    assert np.allclose(rdf.rdf, LJ_RDF, rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("average", [False, True])
def test_gc_steinhardt_reference_values(average):
    """tests/validation/test_steinhardt_average.py:82-118 (test_gc_radius, test_gc_radius_ave)."""
    from freud_b200 import locality, order

    box, pts = readers.read_gsd(os.path.join(HERE, "Test_Configuration.gsd"))[-1]
    nq = locality.AABBQuery.from_system((box, pts))
    nlist = nq.query(pts, dict(mode="ball", r_max=1.4, exclude_ii=True)).toNeighborList()
    want = _gc("GC_rc1.4_avq4avq6avw4avw6" if average else "GC_rc1.4_q4q6w4w6")
    params = [dict(l=4), dict(l=6), dict(l=4, wl=True, wl_normalize=True), dict(l=6, wl=True, wl_normalize=True)]
    for k, (name, p) in enumerate(zip(("q4", "q6", "w4", "w6"), params)):
        op = order.Steinhardt(average=average, weighted=False, **p).compute((box, pts), neighbors=nlist)
        assert np.allclose(want[:, k], op.particle_order, atol=2e-5), ("ave. " if average else "") + name
