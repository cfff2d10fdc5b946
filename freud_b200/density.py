"""``freud.density.RDF`` (reference ``freud/density.py:535-696`` + the ``_SpatialHistogram1D`` properties of
``freud/locality.py:1019-1098``) ``freud.density.LocalDensity`` (``freud/density.py:418-533``) and ``freud.density.CorrelationFunction``
(``freud/density.py:31-165``) on the GPU path."""

import numpy as np

from .locality import _computed, _ext, _PairCompute


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
        self._called_compute = True
        return self

    rdf = _computed(lambda self: self._cpp_obj.getRDF())
    n_r = _computed(lambda self: self._cpp_obj.getNr())
    bin_counts = _computed(lambda self: self._cpp_obj.getBinCounts())
    bin_edges = property(lambda self: np.array(self._cpp_obj.getBinEdges()[0], dtype=np.float32))
    bin_centers = property(lambda self: np.array(self._cpp_obj.getBinCenters()[0], dtype=np.float32))
    bounds = property(lambda self: tuple(self._cpp_obj.getBounds()[0]))
    nbins = property(lambda self: self._cpp_obj.getAxisSizes()[0])

    @_computed
    def box(self):
        from .box import Box

        b = self._cpp_obj.getBox()
        return Box(b.getLx(), b.getLy(), b.getLz(), b.getTiltFactorXY(), b.getTiltFactorXZ(), b.getTiltFactorYZ(),
                   b.is2D())


def _box_of(cpp_box):
    from .box import Box

    return Box(cpp_box.getLx(), cpp_box.getLy(), cpp_box.getLz(), cpp_box.getTiltFactorXY(), cpp_box.getTiltFactorXZ(),
               cpp_box.getTiltFactorYZ(), cpp_box.is2D())


class LocalDensity(_PairCompute):
    """``freud.density.LocalDensity``: fractional neighbour count inside ``r_max`` and the density it implies."""

    def __init__(self, r_max, diameter):
        self._cpp_obj = _ext()._density.LocalDensity(float(r_max), float(diameter))

    r_max = property(lambda self: self._cpp_obj.getRMax())
    diameter = property(lambda self: self._cpp_obj.getDiameter())

    @property
    def default_query_args(self):
        return dict(mode="ball", r_max=self.r_max + 0.5 * self.diameter)  # freud/density.py:510-514

    def compute(self, system, query_points=None, neighbors=None):
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        self._cpp_obj.compute(nq._cpp_obj, qp, nlist, qargs)
        self._called_compute = True
        return self

    box = _computed(lambda self: _box_of(self._cpp_obj.box))
    density = _computed(lambda self: self._cpp_obj.density)
    num_neighbors = _computed(lambda self: self._cpp_obj.num_neighbors)

    def __repr__(self):
        return f"freud.density.{type(self).__name__}(r_max={self.r_max}, diameter={self.diameter})"


class CorrelationFunction(_PairCompute):
    """``freud.density.CorrelationFunction``: C(r) = <conj(values_j) query_values_i> over the bonds of each distance bin."""

    def __init__(self, bins, r_max):
        self._cpp_obj = _ext()._density.CorrelationFunction(int(bins), float(r_max))
        self.r_max = float(r_max)
        self.is_complex = False

    @property
    def default_query_args(self):
        return dict(mode="ball", r_max=self.r_max)  # freud/locality.py:1013-1016

    def compute(self, system, values, query_points=None, query_values=None, neighbors=None, reset=True):
        if reset:
            self.is_complex = False
            self._cpp_obj.reset()
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        # freud/density.py:109-113: complex inputs in any accumulated frame make the result complex
        self.is_complex = bool(self.is_complex or np.any(np.iscomplex(values))
                               or (query_values is not None and np.any(np.iscomplex(query_values))))
        values = np.ascontiguousarray(values, dtype=np.complex128).ravel()
        # freud/density.py:131-135: the points correlate with themselves unless query points (and values) are given
        query_values = values if query_values is None else np.ascontiguousarray(query_values, dtype=np.complex128).ravel()
        self._cpp_obj.accumulateCF(nq._cpp_obj, values, qp, query_values, nlist, qargs)
        self._called_compute = True
        return self

    @_computed
    def correlation(self):
        c = self._cpp_obj.getCorrelation()
        return c if self.is_complex else np.real(c)  # freud/density.py:139-144

    bin_counts = _computed(lambda self: self._cpp_obj.getBinCounts())
    bin_edges = property(lambda self: np.array(self._cpp_obj.getBinEdges()[0], dtype=np.float32))
    bin_centers = property(lambda self: np.array(self._cpp_obj.getBinCenters()[0], dtype=np.float32))
    bounds = property(lambda self: tuple(self._cpp_obj.getBounds()[0]))
    nbins = property(lambda self: self._cpp_obj.getAxisSizes()[0])
    box = _computed(lambda self: _box_of(self._cpp_obj.getBox()))
