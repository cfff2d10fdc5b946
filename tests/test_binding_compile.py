"""INTEGRATION.md's claim, compiled: the reference's binding translation units for this path build UNCHANGED against the
C++ classes of freud_b200/host.

The five files the reference binds the path with -- freud/locality/export-NeighborQuery.cc, export-NeighborList.cc,
export-BondHistogramCompute.cc, freud/density/export-RDF.cc, freud/order/export-Steinhardt.cc -- are copied at test time
into a scratch directory (so that their quoted includes cannot fall back on the stock headers next to them) and
syntax-checked with `-I freud_b200/host` and the nanobind stand-in of tests/nanobind_shim (nanobind is not installed in
this image).  Every member pointer the bindings take and every wrapper body (constructor calls, accumulate /
compute argument lists) therefore type-checks against the replacement classes.  Needs /root/reference, so it runs in
the build container only.
"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/freud"
FILES = ["locality/export-NeighborQuery.cc", "locality/export-NeighborList.cc", "locality/export-BondHistogramCompute.cc",
         "density/export-RDF.cc", "order/export-Steinhardt.cc"]


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference absent (GPU box)")
@pytest.mark.parametrize("rel", FILES)
def test_reference_binding_file_compiles_unchanged(rel, tmp_path):
    src = tmp_path / os.path.basename(rel)
    shutil.copyfile(os.path.join(REF, rel), src)
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "nanobind_shim"),
           "-I", os.path.join(ROOT, "freud_b200", "host"), "-I", os.path.join(ROOT, "include"), str(src)]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert proc.returncode == 0, proc.stdout[-4000:]


def test_host_headers_the_bindings_include_exist():
    for name in ("AABBQuery.h", "Box.h", "CellQuery.h", "LinkCell.h", "NeighborQuery.h", "RawPoints.h", "VectorMath.h",
                 "NeighborBond.h", "NeighborList.h", "BondHistogramCompute.h", "RDF.h", "Steinhardt.h", "ManagedArray.h"):
        assert os.path.exists(os.path.join(ROOT, "freud_b200", "host", name)), name
