"""``freud.density.RDF`` on the GPU path (reference ``freud/density.py:535-696`` + the ``_SpatialHistogram1D``
properties of ``freud/locality.py:1019-1098``)."""

import numpy as np

from .locality import _ext, _PairCompute


class RDF(_PairCompute):
    def __init__(self, bins, r_max, r_min=0, normalization_mode="exact"):
        self._cpp_obj = _ext()._density.RDF(int(bins), float(r_max), float(r_min))
        self.r_max = float(r_max)
        self.mode = normalization_mode

    @property
    def mode(self):
        return "exact" if self._cpp_obj.mode == _ext()._density.NormalizationMode.exact else "finite_size"

    @mode.setter
    def mode(self, value):
        modes = {"exact": _ext()._density.NormalizationMode.exact,
                 "finite_size": _ext()._density.NormalizationMode.finite_size}
        if value not in modes:
            raise ValueError(f"invalid input {value} for normalization_mode")
        self._cpp_obj.mode = modes[value]

    @property
    def default_query_args(self):
        return dict(mode="ball", r_max=self.r_max)  # freud/locality.py:1013-1016

    def compute(self, system, query_points=None, neighbors=None, reset=True):
        if reset:
            self._cpp_obj.reset()
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        self._cpp_obj.accumulateRDF(nq._cpp_obj, qp, nlist, qargs)
        return self

    rdf = property(lambda self: self._cpp_obj.getRDF())
    n_r = property(lambda self: self._cpp_obj.getNr())
    bin_counts = property(lambda self: self._cpp_obj.getBinCounts())
    bin_edges = property(lambda self: np.array(self._cpp_obj.getBinEdges()[0], dtype=np.float32))
    bin_centers = property(lambda self: np.array(self._cpp_obj.getBinCenters()[0], dtype=np.float32))
    bounds = property(lambda self: tuple(self._cpp_obj.getBounds()[0]))
    nbins = property(lambda self: self._cpp_obj.getAxisSizes()[0])

    @property
    def box(self):
        from .box import Box

        b = self._cpp_obj.getBox()
        return Box(b.getLx(), b.getLy(), b.getLz(), b.getTiltFactorXY(), b.getTiltFactorXZ(), b.getTiltFactorYZ(),
                   b.is2D())
