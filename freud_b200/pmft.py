"""``freud.pmft.PMFTXY``, ``PMFTXYZ``, ``PMFTXYT`` and ``PMFTR12`` on the GPU path (reference ``freud/pmft.py:124-590`` +
``_PMFT`` :97-121 and the ``_SpatialHistogram`` properties of ``freud/locality.py:1019-1098``)."""

import numpy as np

from .density import _box_of
from .locality import _computed, _ext, _PairCompute


def _quat_to_z_angle(orientations, num_points):
    """freud/pmft.py:58-84: a last dimension of length 4 means quaternions -- unless there are exactly four points and the
    array is 1-D, which is four angles.  Quaternions must be rotations about +z (or all be the identity); they become
    their rotation angle as ``rowan.to_axis_angle`` defines it: angle = 2 atan2(|v|, w), axis = v / sin(angle / 2)."""
    is_quat = (orientations.ndim == 1 and orientations.shape[0] == 4 and num_points != 4) or (
        orientations.ndim == 2 and orientations.shape[1] == 4)
    if not is_quat:
        return orientations
    q = np.asarray(orientations, dtype=np.float64)
    angles = 2.0 * np.atleast_1d(np.arctan2(np.linalg.norm(q[..., 1:], axis=-1), q[..., 0]))
    sines = np.sin(angles / 2.0)
    sines[sines == 0] = 1.0
    axes = np.where(angles[..., np.newaxis] != 0, np.atleast_2d(q)[..., 1:] / sines[..., np.newaxis], 0.0)
    axes, angles = axes.squeeze(), angles.squeeze()
    if not (np.allclose(angles, 0) or np.allclose(axes, [0, 0, 1])):
        raise ValueError("Orientations provided as quaternions must represent rotations about the z-axis.")
    return angles


def _angles(orientations, n):
    """Angles in radians, one per (query) point (``_gen_angle_array``, freud/pmft.py:87-94)."""
    a = _quat_to_z_angle(np.asarray(orientations).squeeze(), n)
    a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.float32)
    if a.shape != (n,):
        raise ValueError(f"orientations must have shape ({n},) or ({n}, 4)")
    return a


class _PMFT(_PairCompute):
    """freud/pmft.py:97-121 and the spatial-histogram properties every PMFT exposes."""

    @property
    def default_query_args(self):
        return dict(mode="ball", r_max=self.r_max)  # freud/locality.py:1013-1016

    _pcf = _computed(lambda self: self._cpp_obj.getPCF())
    bin_counts = _computed(lambda self: self._cpp_obj.getBinCounts())
    box = _computed(lambda self: _box_of(self._cpp_obj.getBox()))

    @property
    def pmft(self):
        with np.errstate(divide="ignore"):
            return -np.log(np.copy(self._pcf))  # freud/pmft.py:111-116

    bin_edges = property(lambda self: [np.array(e, dtype=np.float32) for e in self._cpp_obj.getBinEdges()])
    bin_centers = property(lambda self: [np.array(c, dtype=np.float32) for c in self._cpp_obj.getBinCenters()])
    bounds = property(lambda self: [tuple(b) for b in self._cpp_obj.getBounds()])
    nbins = property(lambda self: tuple(self._cpp_obj.getAxisSizes()))

    def _bins_repr(self):
        return ", ".join(str(n) for n in self.nbins)


def _three(bins):
    try:
        a, b, c = bins
    except TypeError:
        a = b = c = bins
    return int(a), int(b), int(c)


class PMFTXY(_PMFT):
    def __init__(self, x_max, y_max, bins):
        try:
            n_x, n_y = bins
        except TypeError:
            n_x = n_y = bins
        self._cpp_obj = _ext()._pmft.PMFTXY(float(x_max), float(y_max), int(n_x), int(n_y))
        self.r_max = float(np.sqrt(x_max ** 2 + y_max ** 2))

    def compute(self, system, query_orientations, query_points=None, neighbors=None, reset=True):
        if reset:
            self._cpp_obj.reset()
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        self._cpp_obj.accumulate(nq._cpp_obj, _angles(query_orientations, len(qp)), qp, nlist, qargs)
        self._called_compute = True
        return self

    def __repr__(self):
        b = self.bounds
        return f"freud.pmft.PMFTXY(x_max={b[0][1]}, y_max={b[1][1]}, bins=({self._bins_repr()}))"


class _PMFTAngles(_PMFT):
    """compute() of the two 2-D PMFTs that take the orientation of both particles (freud/pmft.py:150-214, 247-312)."""

    def compute(self, system, orientations, query_points=None, query_orientations=None, neighbors=None, reset=True):
        if reset:
            self._cpp_obj.reset()
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        o = _angles(orientations, len(nq.points))
        qo = o if query_orientations is None else _angles(query_orientations, len(qp))
        if len(qo) != len(qp):
            raise ValueError("query_orientations must hold one orientation per query point")
        self._cpp_obj.accumulate(nq._cpp_obj, o, qp, qo, nlist, qargs)
        self._called_compute = True
        return self

    #: bonds whose angle bin was decided by the host's libm rather than on the GPU (see csrc/pmft.cu)
    host_binned_bonds = property(lambda self: self._cpp_obj.getHostBinnedBonds())


class PMFTR12(_PMFTAngles):
    """freud/pmft.py:124-222: bins over (r, theta_1, theta_2)."""

    def __init__(self, r_max, bins):
        n_r, n_t1, n_t2 = _three(bins)
        self._cpp_obj = _ext()._pmft.PMFTR12(float(r_max), n_r, n_t1, n_t2)
        self.r_max = r_max

    def __repr__(self):
        return f"freud.pmft.PMFTR12(r_max={self.r_max}, bins=({self._bins_repr()}))"


class PMFTXYT(_PMFTAngles):
    """freud/pmft.py:225-325: bins over (x, y, theta)."""

    def __init__(self, x_max, y_max, bins):
        n_x, n_y, n_t = _three(bins)
        self._cpp_obj = _ext()._pmft.PMFTXYT(float(x_max), float(y_max), n_x, n_y, n_t)
        self.r_max = float(np.sqrt(x_max ** 2 + y_max ** 2))

    def __repr__(self):
        b = self.bounds
        return f"freud.pmft.PMFTXYT(x_max={b[0][1]}, y_max={b[1][1]}, bins=({self._bins_repr()}))"


class PMFTXYZ(_PMFT):
    """freud/pmft.py:443-590: bins over the bond vector in the frame of the query particle; ``shiftvec`` is subtracted
    from the query points, ``equiv_orientations`` (default: the identity) each add one count per bond."""

    def __init__(self, x_max, y_max, z_max, bins, shiftvec=None):
        n_x, n_y, n_z = _three(bins)
        self._cpp_obj = _ext()._pmft.PMFTXYZ(float(x_max), float(y_max), float(z_max), n_x, n_y, n_z)
        self.shiftvec = np.array([0, 0, 0] if shiftvec is None else shiftvec, dtype=np.float32)
        self.r_max = float(np.sqrt(x_max ** 2 + y_max ** 2 + z_max ** 2))

    def compute(self, system, query_orientations, query_points=None, equiv_orientations=None, neighbors=None,
                reset=True):
        if reset:
            self._cpp_obj.reset()
        nq, nlist, qargs, qp = self._preprocess_arguments(system, query_points, neighbors)
        qp = np.ascontiguousarray(qp - self.shiftvec.reshape(1, 3), dtype=np.float32)
        qo = np.ascontiguousarray(np.atleast_2d(query_orientations), dtype=np.float32)
        if qo.shape != (len(qp), 4):
            raise ValueError(f"query_orientations must have shape ({len(qp)}, 4)")
        if equiv_orientations is None:
            eq = np.array([[1, 0, 0, 0]], dtype=np.float32)
        else:
            eq = np.ascontiguousarray(equiv_orientations, dtype=np.float32)
            if eq.ndim != 2 or eq.shape[1] != 4:
                raise ValueError("equiv_orientations must have shape (N, 4)")
        self._cpp_obj.accumulate(nq._cpp_obj, qo, qp, eq, nlist, qargs)
        self._called_compute = True
        return self

    def __repr__(self):
        b = self.bounds
        return (f"freud.pmft.PMFTXYZ(x_max={b[0][1]}, y_max={b[1][1]}, z_max={b[2][1]}, bins=({self._bins_repr()}), "
                f"shiftvec={self.shiftvec.tolist()})")
