"""``freud.pmft.PMFTXY`` on the GPU path (reference ``freud/pmft.py:328-440`` + ``_PMFT`` :97-121 and the
``_SpatialHistogram`` properties of ``freud/locality.py:1019-1098``)."""

import numpy as np

from .density import _box_of
from .locality import _ext, _PairCompute


def _angles(orientations, n):
    """Angles in radians, one per query point; (N, 4) quaternions are reduced to their rotation about z
    (``freud/pmft.py:60-94``)."""
    a = np.asarray(orientations, dtype=np.float64).squeeze()
    if a.ndim == 2 and a.shape[1] == 4:
        a = 2.0 * np.arctan2(a[:, 3], a[:, 0])
    a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.float32)
    if a.shape != (n,):
        raise ValueError(f"orientations must have shape ({n},) or ({n}, 4)")
    return a


class PMFTXY(_PairCompute):
    def __init__(self, x_max, y_max, bins):
        try:
            n_x, n_y = bins
        except TypeError:
            n_x = n_y = bins
        self._cpp_obj = _ext()._pmft.PMFTXY(float(x_max), float(y_max), int(n_x), int(n_y))
        self.r_max = float(np.sqrt(x_max ** 2 + y_max ** 2))

    @property
    def default_query_args(self):
        return dict(mode="ball", r_max=self.r_max)  # freud/locality.py:1013-1016

    def compute(self, system, query_orientations, query_points=None, neighbors=None, reset=True):
        if reset:
            self._cpp_obj.reset()
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        self._cpp_obj.accumulate(nq._cpp_obj, _angles(query_orientations, len(qp)), qp, nlist, qargs)
        return self

    @property
    def _pcf(self):
        return self._cpp_obj.getPCF()

    @property
    def pmft(self):
        with np.errstate(divide="ignore"):
            return -np.log(np.copy(self._pcf))  # freud/pmft.py:111-116

    bin_counts = property(lambda self: self._cpp_obj.getBinCounts())
    bin_edges = property(lambda self: [np.array(e, dtype=np.float32) for e in self._cpp_obj.getBinEdges()])
    bin_centers = property(lambda self: [np.array(c, dtype=np.float32) for c in self._cpp_obj.getBinCenters()])
    bounds = property(lambda self: [tuple(b) for b in self._cpp_obj.getBounds()])
    nbins = property(lambda self: tuple(self._cpp_obj.getAxisSizes()))
    box = property(lambda self: _box_of(self._cpp_obj.getBox()))

    def __repr__(self):
        b = self.bounds
        return f"freud.pmft.PMFTXY(x_max={b[0][1]}, y_max={b[1][1]}, bins=({', '.join(str(n) for n in self.nbins)}))"
