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
        """Accepts a Box, an object with Lx/Ly/Lz/xy/xz/yz attributes, a dict, or a 2/3/6 sequence
        (reference: freud/box.py:754-845)."""
        if isinstance(box, cls):
            return box
        if hasattr(box, "Lx"):
            vals = [box.Lx, box.Ly, getattr(box, "Lz", 0), getattr(box, "xy", 0), getattr(box, "xz", 0),
                    getattr(box, "yz", 0)]
            is2d = getattr(box, "is2D", getattr(box, "dimensions", 3) == 2)
            if callable(is2d):
                is2d = is2d()
            return cls(*vals, is2D=bool(is2d) if dimensions is None else dimensions == 2)
        if isinstance(box, dict):
            return cls(box["Lx"], box["Ly"], box.get("Lz", 0), box.get("xy", 0), box.get("xz", 0), box.get("yz", 0),
                       is2D=(box.get("dimensions", 3) == 2) if dimensions is None else dimensions == 2)
        seq = np.asarray(box, dtype=np.float64).ravel()
        if seq.size == 2:
            return cls(seq[0], seq[1], 0, 0, 0, 0, is2D=True)
        if seq.size == 3:
            return cls(seq[0], seq[1], seq[2], 0, 0, 0, is2D=(seq[2] == 0) if dimensions is None else dimensions == 2)
        if seq.size == 6:
            return cls(*seq, is2D=(seq[2] == 0) if dimensions is None else dimensions == 2)
        raise ValueError("Cannot interpret box: expected a Box, a dict, or a sequence of 2, 3 or 6 numbers.")

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
