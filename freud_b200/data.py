"""Synthetic inputs for the neighbour-query path (host-side, numpy only).

``make_random_system`` restates ``freud.data.make_random_system`` (reference ``freud/data.py:350-376``
+ ``freud/box/Box.h:212-222``) and yields the identical float32 bits for the same seed.
``UnitCell`` restates ``freud.data.UnitCell`` (``freud/data.py:14-330``; lattices, replication through
``locality.PeriodicBuffer``, noise) -- the generator behind BASELINE.json's FCC config.
``make_fcc_system`` (the bench's generator; no replication buffer) yields the same noisy FCC point *set* as
``freud.data.UnitCell.fcc().generate_system(n, scale, sigma_noise, seed)`` (``freud/data.py:44-153,
207-215``) up to a rigid lattice translation and the order of the particles; every quantity computed
on the path is per particle and translation invariant, so nothing downstream depends on either.
"""

import numpy as np

from .box import Box


def make_random_system(box_size, num_points, is2D=False, seed=None, tilt=None):
    """Uniform random points in a cubic (3-D) or square (2-D) periodic box.

    ``tilt=(xy, xz, yz)`` extends the reference recipe to the triclinic configs of BASELINE.json
    (config 4): the same fractional coordinates are pushed through ``Box.make_absolute``.
    """
    rs = np.random.RandomState(seed)
    frac = rs.random_sample((num_points, 3))
    if is2D:
        frac[:, 2] = 0
        box = Box.square(box_size) if tilt is None else Box(box_size, box_size, 0, tilt[0], 0, 0, is2D=True)
    else:
        box = Box.cube(box_size) if tilt is None else Box(box_size, box_size, box_size, *tilt, is2D=False)
    return box, box.make_absolute(frac)


def make_fcc_system(num_replicas, scale=1.0, sigma_noise=0.0, seed=None):
    """Face-centred cubic lattice, ``4 * num_replicas**3`` particles, Gaussian noise, wrapped."""
    n = int(num_replicas)
    basis = np.array([[0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5], [0.0, 0.0, 0.0]])
    g = np.arange(n, dtype=np.float64)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    pos = (basis[:, None, :] + cells[None, :, :]).reshape(-1, 3) * scale
    box = Box.cube(n * scale)
    pos = (pos - 0.5 * n * scale).astype(np.float32)
    pos = box.wrap(pos)
    if sigma_noise != 0:
        rs = np.random.RandomState(seed)
        var = sigma_noise * sigma_noise
        pos = pos + rs.multivariate_normal([0, 0, 0], np.diag([var, var, var]), size=len(pos)).astype(np.float32)
    return box, box.wrap(pos)


class UnitCell:
    """A crystal unit cell: a box of lattice vectors and fractional basis positions (freud/data.py:14-56).

    ``generate_system`` replicates it, optionally adds Gaussian noise and wraps -- point order as upstream: all
    replicas of the first basis position, then of the second, ... (``numpy.repeat`` order), each block running
    over the images with z fastest."""

    def __init__(self, box, basis_positions=None):
        self._box = Box.from_box(box)
        self._basis_positions = [[0, 0, 0]] if basis_positions is None else basis_positions

    box = property(lambda self: self._box)
    basis_positions = property(lambda self: self._basis_positions)
    lattice_vectors = property(lambda self: self._box.to_matrix())
    a1 = property(lambda self: self._box.to_matrix()[:, 0])
    a2 = property(lambda self: self._box.to_matrix()[:, 1])
    a3 = property(lambda self: self._box.to_matrix()[:, 2])
    dimensions = property(lambda self: self._box.dimensions)

    def generate_system(self, num_replicas=1, scale=1, sigma_noise=0, seed=None):
        """(box, positions) of ``num_replicas`` (an int or ``(nx, ny, nz)``) copies of the cell, scaled by ``scale``,
        with N(0, sigma_noise^2) displacements drawn from ``numpy.random.RandomState(seed)``
        (freud/data.py:58-150)."""
        from .locality import PeriodicBuffer

        try:
            nx, ny, nz = num_replicas
        except TypeError:
            nx = ny = num_replicas
            nz = 1 if self._box.is2D else num_replicas
        if not all(int(n) == n and n > 0 for n in (nx, ny, nz)):
            raise ValueError("The number of replicas must be a positive integer in each dimension.")
        if self._box.is2D and nz != 1:
            raise ValueError("The number of replicas in z must be 1 for a 2D unit cell.")
        basis = self._box.make_absolute(self._basis_positions)
        if max(nx, ny, nz) > 1:
            pb = PeriodicBuffer().compute((self._box, basis), buffer=(nx - 1, ny - 1, nz - 1), images=True,
                                          include_input_points=True)
            box, positions = pb.buffer_box * scale, pb.buffer_points.copy()
        else:
            box, positions = self._box * scale, basis
        # an even number of replicas puts a lattice plane where an odd number puts a cell centre: shift by L/2
        even = (np.array([nx, ny, nz]) + 1) % 2
        positions = (positions + even * self._box.make_absolute([1, 1, 1])).astype(np.float32)
        positions = box.wrap(positions * np.float32(scale))
        if sigma_noise != 0:
            var = sigma_noise * sigma_noise
            cov = np.diag([var, var, var if self.dimensions == 3 else 0])
            noise = np.random.RandomState(seed).multivariate_normal([0, 0, 0], cov, size=positions.shape[:-1])
            positions = (positions + noise).astype(np.float32)
        return box, box.wrap(positions)

    # -- the lattices of freud/data.py:157-268 ---------------------------------------------------------------
    @classmethod
    def fcc(cls):
        return cls([1, 1, 1], np.array([[0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5], [0, 0, 0]]))

    @classmethod
    def bcc(cls):
        return cls([1, 1, 1], np.array([[0.5, 0.5, 0.5], [0, 0, 0]]))

    @classmethod
    def sc(cls):
        return cls([1, 1, 1], np.array([[0, 0, 0]]))

    @classmethod
    def hcp(cls):
        box = Box.from_box_lengths_and_angles(1, 1, np.sqrt(8 / 3), *np.deg2rad([90, 90, 120.0]))
        return cls(box, np.array([[1 / 3, 2 / 3, 1 / 4], [2 / 3, 1 / 3, 3 / 4]]))

    @classmethod
    def square(cls):
        return cls([1, 1], np.array([[0, 0, 0]]))

    @classmethod
    def rectangular(cls, aspect=2.0, centered=False):
        return cls([1, aspect], np.array([[0, 0, 0]] + ([[0.5, 0.5, 0.0]] if centered else [])))

    @classmethod
    def oblique(cls, aspect=1.0, theta=45.0):
        box = Box.from_box_lengths_and_angles(1, aspect, 0, np.pi / 2, np.pi / 2, np.deg2rad(theta))
        return cls(box, np.array([[0.0, 0.0, 0.0]]))

    @classmethod
    def hex(cls):
        return cls([1, np.sqrt(3)], np.array([[0, 0, 0], [0.5, 0.5, 0]]))

    @classmethod
    def graphene(cls):
        return cls([1, np.sqrt(3)], np.array([[0, 0, 0], [0, 1 / 3, 0], [1 / 2, 5 / 6, 0], [1 / 2, 1 / 2, 0]]))
