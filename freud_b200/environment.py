"""``freud.environment.BondOrder`` on the GPU path (reference ``freud/environment.py:204-391`` and the
``_SpatialHistogram`` properties of ``freud/locality.py:1019-1098``)."""

import numpy as np

from .density import _box_of
from .locality import _computed, _ext, _PairCompute


def _quats(orientations, n, name):
    q = np.ascontiguousarray(orientations, dtype=np.float32)
    if q.shape != (n, 4):
        raise ValueError(f"{name} must have shape ({n}, 4)")
    return q


class BondOrder(_PairCompute):
    """Bond orientational order diagram: the histogram of bond directions on the sphere, (theta, phi) bins divided by
    their solid angle.  ``mode``: ``'bod'`` bond vectors as they are, ``'lbod'`` in the frame of the neighbouring point's
    orientation, ``'obcd'`` also rotated by the query particle's orientation, ``'oocd'`` the director of the query
    particle in that frame (freud/environment.py:213-256, BondOrder.cc:108-135)."""

    known_modes = ("bod", "lbod", "obcd", "oocd")

    def __init__(self, bins, mode="bod"):
        try:
            n_bins_theta, n_bins_phi = bins
        except TypeError:
            n_bins_theta = n_bins_phi = bins
        if mode not in self.known_modes:
            raise ValueError(f"Unknown BondOrder mode: {mode}")
        env = _ext()._environment
        self._cpp_obj = env.BondOrder(int(n_bins_theta), int(n_bins_phi), getattr(env, mode))

    @property
    def default_query_args(self):
        """No default query arguments (freud/environment.py:286-292)."""
        raise NotImplementedError("The BondOrder class does not provide default query arguments. You must either "
                                  "provide query arguments or a neighbor list to this compute method.")

    def compute(self, system, orientations=None, query_points=None, query_orientations=None, neighbors=None, reset=True):
        if reset:
            self._cpp_obj.reset()
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        n = len(nq.points)
        if orientations is None:
            orientations = np.tile(np.float32([1, 0, 0, 0]), (n, 1))  # freud/environment.py:341-342
        o = _quats(orientations, n, "orientations")
        qo = o if query_orientations is None else _quats(query_orientations, len(qp), "query_orientations")
        if len(qo) != len(qp):
            raise ValueError("query_orientations must hold one quaternion per query point")
        self._cpp_obj.accumulate(nq._cpp_obj, o, qp, qo, nlist, qargs)
        self._called_compute = True
        return self

    bond_order = _computed(lambda self: self._cpp_obj.getBondOrder())
    bin_counts = _computed(lambda self: self._cpp_obj.getBinCounts())
    box = _computed(lambda self: _box_of(self._cpp_obj.getBox()))
    bin_edges = property(lambda self: [np.array(e, dtype=np.float32) for e in self._cpp_obj.getBinEdges()])
    bin_centers = property(lambda self: [np.array(c, dtype=np.float32) for c in self._cpp_obj.getBinCenters()])
    bounds = property(lambda self: [tuple(b) for b in self._cpp_obj.getBounds()])
    nbins = property(lambda self: tuple(self._cpp_obj.getAxisSizes()))
    mode = property(lambda self: self.known_modes[int(self._cpp_obj.getMode())])
    #: bonds whose bin was decided by the host's libm rather than on the GPU (see csrc/pmft.cu)
    host_binned_bonds = property(lambda self: self._cpp_obj.getHostBinnedBonds())

    def __repr__(self):
        return f"freud.environment.BondOrder(bins=({', '.join(str(n) for n in self.nbins)}), mode='{self.mode}')"
