"""Periodic simulation box (host-side value type).

Mirrors the part of ``freud.box.Box`` the neighbour-query path needs (reference:
``freud/box.py:30-160, 199-330, 885-918`` for the Python surface and ``freud/box/Box.h:100-115,
212-255, 307-329, 489-518`` for the float32 arithmetic).  The array helpers here (``make_absolute``,
``make_fractional``, ``wrap``) are host utilities used to *prepare inputs* (``freud_b200.data``); the
per-pair minimum-image arithmetic of the hot path lives in the CUDA kernels
(``freud_b200/csrc/pair_math.cuh``) and is not routed through this module.

All arithmetic is float32 with one rounding per operation, in the reference's operation order, so
``make_absolute`` reproduces ``freud.data.make_random_system`` bit for bit.
"""

import numpy as np

_F = np.float32


class Box:
    def __init__(self, Lx, Ly, Lz=0.0, xy=0.0, xz=0.0, yz=0.0, is2D=None):
        if is2D is None:
            is2D = Lz == 0
        if is2D and Lz != 0:
            # freud warns and zeroes Lz (freud/box.py:93-100)
            Lz = 0.0
        self._is2D = bool(is2D)
        self._L = np.array([Lx, Ly, 0.0 if self._is2D else Lz], dtype=_F)
        self._tilt = np.array([xy, xz, yz], dtype=_F)
        if not (self._L[0] > 0 and self._L[1] > 0 and (self._is2D or self._L[2] > 0)):
            raise ValueError("Box lengths must be positive (Lz may be 0 only for 2D boxes).")

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def cube(cls, L):
        return cls(L, L, L, 0, 0, 0, is2D=False)

    @classmethod
    def square(cls, L):
        return cls(L, L, 0, 0, 0, 0, is2D=True)

    @classmethod
    def from_box(cls, box, dimensions=None):
        """Coerces a box-like object (reference: freud/box.py:751-843): a Box, an object with ``Lx, Ly[, Lz, xy, xz,
        yz, dimensions]`` attributes, a mapping with those keys, a 3x3 matrix of lattice vectors (columns), or a
        sequence ``[Lx, Ly]``, ``[Lx, Ly, Lz]`` or ``[Lx, Ly, Lz, xy, xz, yz]``.  ``dimensions`` overrides the
        detected dimensionality (2 if ``Lz == 0``) but may not contradict a ``dimensions`` the object carries."""
        if isinstance(box, cls) and dimensions in (None, box.dimensions):
            return box
        if not isinstance(box, dict) and not hasattr(box, "Lx") and np.shape(box) == (3, 3):
            return cls.from_matrix(box, dimensions=dimensions)

        def reconcile(own):
            if dimensions is not None and own is not None and own != dimensions:
                raise ValueError("The provided dimensions argument conflicts with the dimensions attribute of the "
                                 "provided box object.")
            return own if dimensions is None else dimensions

        if hasattr(box, "Lx") and hasattr(box, "Ly"):
            vals = [box.Lx, box.Ly] + [getattr(box, name, 0) for name in ("Lz", "xy", "xz", "yz")]
            dims = reconcile(getattr(box, "dimensions", None))
        elif isinstance(box, dict) or (hasattr(box, "keys") and hasattr(box, "get")):
            try:
                vals = [box["Lx"], box["Ly"]] + [box.get(name, 0) for name in ("Lz", "xy", "xz", "yz")]
            except KeyError as exc:
                raise ValueError("A box mapping needs at least the keys 'Lx' and 'Ly'.") from exc
            dims = reconcile(box.get("dimensions", None))
        else:
            try:
                n = len(box)
            except TypeError as exc:
                raise ValueError("Supplied box cannot be converted to a Box.") from exc
            if n not in (2, 3, 6):
                raise ValueError("List-like objects must have length 2, 3, or 6 to be converted to a Box.")
            vals = [box[0], box[1], box[2] if n > 2 else 0] + (list(box[3:6]) if n == 6 else [0, 0, 0])
            dims = dimensions
        if dims is None:
            dims = 2 if vals[2] == 0 else 3
        return cls(*vals, is2D=dims == 2)

    @classmethod
    def from_matrix(cls, box_matrix, dimensions=None):
        """Box from the 3x3 matrix whose columns are the lattice vectors, in float32 as upstream
        (freud/box.py:846-882; the HOOMD-blue box-matrix convention)."""
        m = np.asarray(box_matrix, dtype=_F)
        if m.shape != (3, 3):
            raise ValueError("A box matrix must be 3x3.")
        v0, v1, v2 = m[:, 0], m[:, 1], m[:, 2]
        Lx = np.sqrt(np.dot(v0, v0))
        a2x = np.dot(v0, v1) / Lx
        Ly = np.sqrt(np.dot(v1, v1) - a2x * a2x)
        xy = a2x / Ly
        n01 = np.cross(v0, v1)
        Lz = np.dot(v2, n01) / np.sqrt(np.dot(n01, n01))
        xz = yz = 0
        if Lz != 0:
            a3x = np.dot(v0, v2) / Lx
            xz = a3x / Lz
            yz = (np.dot(v1, v2) - a2x * a3x) / (Ly * Lz)
        if dimensions is None:
            dimensions = 2 if Lz == 0 else 3
        return cls(Lx, Ly, Lz, xy, xz, yz, is2D=dimensions == 2)

    @classmethod
    def from_box_lengths_and_angles(cls, L1, L2, L3, alpha, beta, gamma, dimensions=None):
        """Box from three lattice-vector lengths and the angles between them in radians (freud/box.py:921-983)."""
        for name, ang in (("alpha", alpha), ("beta", beta), ("gamma", gamma)):
            if not 0 < ang < np.pi:
                raise ValueError(f"{name} must be between 0 and pi.")
        a1 = np.array([L1, 0, 0])
        a2 = np.array([L2 * np.cos(gamma), L2 * np.sin(gamma), 0])
        a3x = np.cos(beta)
        a3y = (np.cos(alpha) - np.cos(beta) * np.cos(gamma)) / np.sin(gamma)
        under = 1 - a3x**2 - a3y**2
        if under < 0:
            raise ValueError("The provided angles can not form a valid box.")
        a3 = L3 * np.array([a3x, a3y, np.sqrt(under)])
        if dimensions is None:
            dimensions = 2 if L3 == 0 else 3
        return cls.from_matrix(np.array([a1, a2, a3]).T, dimensions=dimensions)

    def to_matrix(self):
        """Columns are the lattice vectors (freud/box.py:617-630)."""
        L, (xy, xz, yz) = self._L.astype(np.float64), self._tilt.astype(np.float64)
        return np.array([[L[0], xy * L[1], xz * L[2]], [0, L[1], yz * L[2]], [0, 0, L[2]]])

    # -- scalar properties ----------------------------------------------------------------------
    Lx = property(lambda self: float(self._L[0]))
    Ly = property(lambda self: float(self._L[1]))
    Lz = property(lambda self: float(self._L[2]))
    xy = property(lambda self: float(self._tilt[0]))
    xz = property(lambda self: float(self._tilt[1]))
    yz = property(lambda self: float(self._tilt[2]))
    is2D = property(lambda self: self._is2D)
    dimensions = property(lambda self: 2 if self._is2D else 3)
    L = property(lambda self: self._L.copy())

    @property
    def volume(self):
        L = self._L
        return float(L[0] * L[1]) if self._is2D else float(L[0] * L[1] * L[2])

    @property
    def periodic(self):
        return np.array([True, True, True])

    def as_array6(self):
        """(Lx, Ly, Lz, xy, xz, yz) float32 -- the layout the C ABI takes."""
        return np.concatenate([self._L, self._tilt]).astype(_F)

    def to_dict(self):
        return dict(Lx=self.Lx, Ly=self.Ly, Lz=self.Lz, xy=self.xy, xz=self.xz, yz=self.yz,
                    dimensions=self.dimensions)

    def nearest_plane_distance(self):
        """freud/box/Box.h:489-497."""
        xy, xz, yz = self._tilt
        L = self._L
        one = _F(1.0)
        t = xy * yz - xz
        return np.array([L[0] / np.sqrt(one + xy * xy + t * t), L[1] / np.sqrt(one + yz * yz), L[2]], dtype=_F)

    def __eq__(self, other):
        return (isinstance(other, Box) and np.array_equal(self._L, other._L)
                and np.array_equal(self._tilt, other._tilt) and self._is2D == other._is2D)

    def __mul__(self, scale):
        s = _F(scale)
        return Box(self._L[0] * s, self._L[1] * s, self._L[2] * s, *self._tilt, is2D=self._is2D)

    def __repr__(self):
        return (f"freud_b200.box.Box(Lx={self.Lx}, Ly={self.Ly}, Lz={self.Lz}, xy={self.xy}, xz={self.xz}, "
                f"yz={self.yz}, is2D={self.is2D})")

    # -- array helpers (input preparation only) -------------------------------------------------
    def _lo(self):
        return -(self._L * _F(0.5))

    def make_absolute(self, fractional_coordinates):
        """freud/box/Box.h:212-222."""
        f = np.atleast_2d(np.asarray(fractional_coordinates)).astype(_F)
        v = self._lo() + f * self._L
        xy, xz, yz = self._tilt
        v[:, 0] = v[:, 0] + (xy * v[:, 1] + xz * v[:, 2])
        v[:, 1] = v[:, 1] + yz * v[:, 2]
        if self._is2D:
            v[:, 2] = 0
        return v

    def make_fractional(self, absolute_coordinates):
        """freud/box/Box.h:243-255."""
        v = np.atleast_2d(np.asarray(absolute_coordinates)).astype(_F)
        xy, xz, yz = self._tilt
        d = v - self._lo()
        d[:, 0] = d[:, 0] - ((xz - yz * xy) * v[:, 2] + xy * v[:, 1])
        d[:, 1] = d[:, 1] - yz * v[:, 2]
        with np.errstate(divide="ignore", invalid="ignore"):
            d = d / self._L
        if self._is2D:
            d[:, 2] = 0
        return d

    def wrap(self, vecs):
        """freud/box/Box.h:307-329 with util::modulusPositive (freud/util/utils.h:29-32)."""
        f = self.make_fractional(vecs)
        one = _F(1.0)
        f = np.fmod(np.fmod(f, one) + one, one)
        return self.make_absolute(f)
