"""Synthetic inputs for the neighbour-query path (host-side, numpy only).

``make_random_system`` restates ``freud.data.make_random_system`` (reference ``freud/data.py:350-376``
+ ``freud/box/Box.h:212-222``) and yields the identical float32 bits for the same seed.
``make_fcc_system`` yields the same noisy FCC point *set* as
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
